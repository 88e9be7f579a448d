"""Profiling helper: Newton-LSQR hunt on 4096 OrbitKS 32x32 seeds after a short adjoint warm-up.
python tools/profile_lsqr.py [newton_iters] [iter_lim]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 5
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 300
B = 4096
u = BatchedOrbitKS.from_seeds(np.arange(B), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
r = hunt(u, "adj", maxiter=2000)
torch.cuda.synchronize()
t0 = time.perf_counter()
r2 = hunt(r.orbit, "lsqr", maxiter=nk, tol=1e-10, scipy_kwargs={"iter_lim": lim})
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"lsqr newton_iters={nk} iter_lim={lim}: wall {dt:.3f} s, {dt / nk * 1e3:.1f} ms per Newton iteration, "
      f"median cost {float(r2.costs[-1].median()):.3e}")
