"""CPU: the multi-GPU path's host logic (seed sharding + final result gather) with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbithunter_b200.sweep import RECORD_FIELDS, gather_records, shard_bounds


def test_shard_bounds_cover_range_without_overlap():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, rank, world)
    seeds = torch.arange(lo, hi, dtype=torch.float64)
    rec = torch.stack([seeds, torch.full_like(seeds, rank), seeds * 2, seeds * 0.5, seeds + 44, seeds + 33, seeds * 0], dim=1)
    got = gather_records(rec)
    if rank == 0:
        torch.save(got, out)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 10, 1])
def test_gather_records_world2_gloo(tmp_path, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got.shape == (total, len(RECORD_FIELDS))
    assert torch.equal(got[:, 0], torch.arange(total, dtype=torch.float64))  # ordered by seed across ranks
    lo1, _ = shard_bounds(total, 1, 2)
    assert torch.equal(got[:, 1], (torch.arange(total) >= lo1).to(torch.float64))
    assert torch.equal(got[:, 3], got[:, 0] * 0.5)
