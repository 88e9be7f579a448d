"""Profiling helper: a short batched adjoint-descent run (4096 OrbitKS 32x32 seeds).  python tools/profile_adj.py [maxiter]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS, OrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402

maxiter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = 4096
base = np.stack([OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=s % 64, discretization=(32, 32)).state
                 for s in range(B)])
orb = BatchedOrbitKS(base, np.tile([44.0, 33.4, 0.0], (B, 1)), "OrbitKS")
hunt(orb, "adj", maxiter=20)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = hunt(orb, "adj", maxiter=maxiter)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"maxiter={maxiter} wall={dt:.4f}s per_pass_us={1e6 * dt / (maxiter + 8):.1f} steps/s={float(r.nit.sum()) / dt:.3e}")
