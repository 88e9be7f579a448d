"""Per-seed outcome of the two-stage hunt (adj 10000 -> lstsq 200) on seeds 0..N-1 of the configs[1]/[4] workload, saved to
gpurun_out/ for offline analysis (which adjoint end costs go on to converge; iterations spent per outcome).

    python tools/sweep_stats.py --seeds 4096 --out gpurun_out/sweep_stats.npz
"""
import argparse
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
    ap.add_argument("--lstsq-maxiter", type=int, default=200)
    ap.add_argument("--out", default="gpurun_out/sweep_stats.npz")
    args = ap.parse_args()
    b = BatchedOrbitKS.from_seeds(np.arange(args.seeds), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    t0 = time.perf_counter()
    r1 = hunt(b, "adj", maxiter=10000, tol=1e-6)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r2 = hunt(r1.orbit, "lstsq", maxiter=args.lstsq_maxiter, tol=1e-10)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    code, _ = r2.orbit.preprocess_codes()
    np.savez_compressed(args.out, adj_cost=r1.costs[-1].cpu().numpy(), adj_status=r1.status.cpu().numpy(), adj_nit=r1.nit.cpu().numpy(),
                        status=r2.status.cpu().numpy(), nit=r2.nit.cpu().numpy(), cost=r2.costs[-1].cpu().numpy(),
                        params=r2.orbit.cpu_parameters(), adj_params=r1.orbit.cpu_parameters(), code=code, seconds=np.array([t1 - t0, t2 - t1]))
    st = r2.status.cpu().numpy()
    print({"adj_s": t1 - t0, "lstsq_s": t2 - t1, "converged": int((st == 1).sum()), "equilibria_among_converged": int(((st == 1) & (code != 0)).sum())})


if __name__ == "__main__":
    main()
