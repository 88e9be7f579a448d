"""Timing of the orbit-hunting sweep (BASELINE configs[4]) stage by stage on this rank's GPU.

    python tools/sweep_bench.py --seeds 4096 [--lstsq-maxiter 200] [--chunk 256]
Seeds 0..N-1 of OrbitKS 32x32, (T, L) = (44, 33.4): device seed generation, hunt('adj', 10000 steps), hunt('lstsq').
One JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=4096)
    ap.add_argument("--adj-maxiter", type=int, default=10000)
    ap.add_argument("--lstsq-maxiter", type=int, default=200)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--refine", type=int, default=2)
    args = ap.parse_args()
    warm = BatchedOrbitKS.from_seeds(np.arange(8), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    hunt(warm, ("adj", "lstsq"), maxiter=[8, 1], tol=[1e-6, 1e-10])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = BatchedOrbitKS.from_seeds(np.arange(args.seeds), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r1 = hunt(b, "adj", maxiter=args.adj_maxiter, tol=1e-6)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    r2 = hunt(r1.orbit, "lstsq", maxiter=args.lstsq_maxiter, tol=1e-10, chunk=args.chunk, refine=args.refine)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    st = r2.status.cpu().numpy()
    nit = r2.nit.cpu().numpy()
    print(json.dumps({"seeds": args.seeds, "seed_gen_s": t1 - t0, "adj_s": t2 - t1, "lstsq_s": t3 - t2, "total_s": t3 - t0,
                      "status_counts": {str(k): int((st == k).sum()) for k in (-1, 0, 1, 2, 3)},
                      "converged": int((st == 1).sum()), "converged_per_s": float((st == 1).sum()) / (t3 - t0),
                      "lstsq_outer_iterations": int(r2.outer_iterations), "dense_solves": int(r2.dense_solves),
                      "us_per_dense_solve": 1e6 * (t3 - t2) / max(int(r2.dense_solves), 1), "lstsq_nit_sum": int(nit.sum()),
                      "chunk": args.chunk, "refine": args.refine}), flush=True)


if __name__ == "__main__":
    main()
