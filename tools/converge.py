"""Experiment: how many random 32x32 OrbitKS seeds converge under adj -> Newton hybrids, and how fast.
python tools/converge.py B"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS, OrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = np.stack([OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=s, discretization=(32, 32)).state for s in range(B)])
u = BatchedOrbitKS(base, np.tile([44.0, 33.4, 0.0], (B, 1)), "OrbitKS")
hunt(u, "adj", maxiter=10)
for tag, methods, kw in (
    ("adj10k->gmres", ("adj", "gmres"), dict(maxiter=[10000, 30], tol=[1e-6, 1e-10], scipy_kwargs=[{}, {"restart": 50, "maxiter": 2, "rtol": 1e-6}])),
    ("adj10k->lsqr", ("adj", "lsqr"), dict(maxiter=[10000, 30], tol=[1e-6, 1e-10], scipy_kwargs=[{}, {"iter_lim": 300}])),
    ("adj10k->newton_descent", ("adj", "newton_descent"), dict(maxiter=[10000, 30], tol=[1e-6, 1e-10])),
):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = hunt(u, methods, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = r.status.cpu().numpy()
    print(json.dumps({"tag": tag, "B": B, "wall_s": dt, "status": {int(k): int((st == k).sum()) for k in np.unique(st)},
                      "converged_per_s": float((st == 1).sum()) / dt, "median_cost": float(r.costs[-1].median())}), flush=True)
