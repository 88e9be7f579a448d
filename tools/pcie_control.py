"""Control run for the end-to-end number: plain pinned-memory copies, no kernels.  Every rank moves the bytes one bench e2e step
moves (30.6 MB up, 30.6 MB down) on two streams, as bench.py does, and reports its own and the aggregate bandwidth.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_control.py      (or plain python for one GPU)
One JSON line from rank 0: GB/s per direction per rank, aggregate, and the evals/s ceiling they imply for the e2e loop."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 4096 * 933  # doubles per direction and step (modes + params up, gradient + params + cost down)
    streams = [torch.cuda.Stream() for _ in range(2)]
    host_up = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in streams]
    host_dn = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in streams]
    dev_up = [torch.empty(n, dtype=torch.float64, device=dev) for _ in streams]
    dev_dn = [torch.empty(n, dtype=torch.float64, device=dev) for _ in streams]

    def step(i):
        k = i % 2
        with torch.cuda.stream(streams[k]):
            dev_up[k].copy_(host_up[k], non_blocking=True)
            host_dn[k].copy_(dev_dn[k], non_blocking=True)

    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    streams[1].wait_event(e0)
    for i in range(steps):
        step(i)
    streams[0].wait_stream(streams[1])
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    every = [torch.zeros_like(ms) for _ in range(world)]
    if world > 1:
        dist.all_gather(every, ms)
    else:
        every = [ms]
    if rank == 0:
        per = [steps * n * 8 / (float(m) * 1e-3) / 1e9 for m in every]
        worst = max(float(m) for m in every)
        print(json.dumps({"n_gpus": world, "bytes_per_direction_per_step": n * 8, "GBps_per_direction_per_rank": per,
                          "aggregate_GBps_both_directions": 2 * sum(per),
                          "e2e_ceiling_evals_per_s": world * 4096 * steps / (worst * 1e-3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
