"""Multi-GPU orbit-hunting sweep: independent seeds sharded over ranks, one result gather at the end.

The reference has no distributed code; orbits are independent, so the only multi-GPU strategy is
a contiguous split of the seed range (one process per GPU, torchrun) and a final gather of the per-orbit result
records over torch.distributed (NCCL on GPUs; gloo in the CPU tests).  No collective runs during the solve.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD_FIELDS = ("seed", "status", "nit", "cost", "t", "x", "s")


def shard_bounds(total, rank, world):
    """Contiguous shard [lo, hi) of range(total) for `rank`; shard sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local: torch.Tensor, dst=0):
    """Gather (n_local, len(RECORD_FIELDS)) float64 record tensors of unequal length onto `dst` (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    nmax = int(max(int(c) for c in counts))
    padded = torch.zeros((nmax, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[: int(c)] for b, c in zip(bufs, counts)], dim=0)


def hunt_sweep(cls_name, n_seeds, T, L, discretization, methods="adj", chunk=4096, first_seed=0, device=None,
               seed_on_device=True, **hunt_kwargs):
    """Hunt seeds [first_seed, first_seed + n_seeds) of class `cls_name`, sharded over the initialised process group.

    Returns on rank 0 a (n_seeds, 7) float64 tensor with columns RECORD_FIELDS, ordered by seed; None on other ranks.
    Seeds come from the reference's generator (populate(attr='state', seed=i), ks/orbits.py:1738): reproduced on the
    device by default (BatchedOrbitKS.from_seeds; same MT19937 stream, normals equal to the last ulp or two), or on the
    host with ``seed_on_device=False`` (bit-identical to the reference, ~115 us per seed per core).
    """
    from .batch import BatchedOrbitKS
    from .ks import CLASSES
    from .optimize import hunt

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_bounds(n_seeds, rank, world)
    device = device or torch.device("cuda", torch.cuda.current_device())
    records = []
    cls = CLASSES[cls_name]
    for c0 in range(lo, hi, chunk):
        seeds = np.arange(first_seed + c0, first_seed + min(c0 + chunk, hi))
        if seed_on_device:
            batch = BatchedOrbitKS.from_seeds(seeds, (float(T), float(L), 0.0), cls_name, discretization, device=device)
        else:
            states = np.stack([cls(parameters=(T, L, 0.0)).populate(attr="state", seed=int(s), discretization=discretization).state
                               for s in seeds])
            params = np.tile([float(T), float(L), 0.0], (len(seeds), 1))
            batch = BatchedOrbitKS(torch.as_tensor(states).to(device), torch.as_tensor(params).to(device), cls_name, "modes")
        res = hunt(batch, methods, **{k: (v.copy() if hasattr(v, "copy") else v) for k, v in hunt_kwargs.items()})
        rec = torch.stack([torch.as_tensor(seeds, dtype=torch.float64, device=device), res.status.to(torch.float64),
                           (res.nit[-1] if isinstance(res.nit, list) else res.nit).to(torch.float64), res.costs[-1],
                           res.orbit.parameters[:, 0], res.orbit.parameters[:, 1], res.orbit.parameters[:, 2]], dim=1)
        records.append(rec)
    local = torch.cat(records, dim=0) if records else torch.zeros((0, len(RECORD_FIELDS)), dtype=torch.float64, device=device)
    return gather_records(local)


SWEEP_FIELDS = RECORD_FIELDS + ("code",)


def two_stage_sweep(n_seeds, refine, T=44.0, L=33.4, cls_name="OrbitKS", discretization=(32, 32), first_seed=0, adj_maxiter=10000,
                    adj_tol=1e-6, dense_maxiter=200, dense_tol=1e-10, dense_chunk=None, device=None):
    """BASELINE configs[4]: the orbit-hunting sweep, strong-scaled over the initialised process group.

    Every rank takes the contiguous shard of seeds [first_seed, first_seed + n_seeds) that ``shard_bounds`` gives it,
    generates them on the device, runs ``hunt('adj')`` (stage 1) on all of them and ``hunt('lstsq')`` (stage 2, the dense
    Newton step that actually converges orbits: configs[0]'s second method) on the first ``refine / world`` of its shard --
    the adjoint end cost does not predict which seeds converge (profiles/r02m_sweep_statistics.json), so the refined subset is
    a uniform sample -- classifies the refined orbits with ``preprocess`` (zero / equilibrium solutions) and gathers one
    record per seed on rank 0.  No collective runs before the gather.

    Returns (records, timings): records on rank 0 is an (n_seeds, 8) float64 tensor, columns SWEEP_FIELDS, ordered by seed
    (code = -1 for seeds that were not refined); timings is a dict of this rank's stage wall times in seconds.
    """
    import time

    from .batch import BatchedOrbitKS
    from .optimize import hunt

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_bounds(n_seeds, rank, world)
    rlo, rhi = shard_bounds(min(refine, n_seeds), rank, world)
    n_ref = min(rhi - rlo, hi - lo)
    device = device or torch.device("cuda", torch.cuda.current_device())
    tm = {}

    def lap(key, t0):
        torch.cuda.synchronize(device)
        tm[key] = time.perf_counter() - t0
        return time.perf_counter()

    t = time.perf_counter()
    seeds = np.arange(first_seed + lo, first_seed + hi)
    batch = BatchedOrbitKS.from_seeds(seeds, (float(T), float(L), 0.0), cls_name, discretization, device=device)
    t = lap("seed_generation_s", t)
    r1 = hunt(batch, "adj", maxiter=adj_maxiter, tol=adj_tol)
    t = lap("adjoint_s", t)
    status, nit, cost = r1.status.to(torch.float64), r1.nit.to(torch.float64), r1.costs[-1].clone()
    params = r1.orbit.parameters.clone()
    code = torch.full((hi - lo,), -1.0, dtype=torch.float64, device=device)
    tm["adjoint_steps"] = float(r1.nit.sum().item())
    tm["dense_solves"] = 0.0
    if n_ref > 0:
        sub = r1.orbit._like(r1.orbit.state[:n_ref].contiguous(), r1.orbit.parameters[:n_ref].contiguous())
        chunk_kw = {} if dense_chunk is None else {"chunk": dense_chunk}
        r2 = hunt(sub, "lstsq", maxiter=dense_maxiter, tol=dense_tol, **chunk_kw)
        t = lap("dense_newton_s", t)
        codes, _ = r2.orbit.preprocess_codes()
        status[:n_ref] = r2.status.to(torch.float64)
        nit[:n_ref] += r2.nit.to(torch.float64)
        cost[:n_ref] = r2.costs[-1]
        params[:n_ref] = r2.orbit.parameters
        code[:n_ref] = torch.as_tensor(codes, dtype=torch.float64, device=device)
        tm["dense_solves"] = float(r2.dense_solves)
        t = lap("preprocess_s", t)
    else:
        tm["dense_newton_s"] = tm["preprocess_s"] = 0.0
    local = torch.stack([torch.as_tensor(seeds, dtype=torch.float64, device=device), status, nit, cost, params[:, 0], params[:, 1],
                         params[:, 2], code], dim=1)
    if world > 1:  # wait for the slowest rank here, so that gather_s is the collective alone
        dist.barrier()
    t = lap("rank_imbalance_s", t)
    records = gather_records(local)
    lap("gather_s", t)
    tm["refined"] = float(n_ref)
    tm["seeds"] = float(hi - lo)
    return records, tm
