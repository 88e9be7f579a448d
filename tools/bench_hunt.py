"""Side measurements of the Newton-Krylov hunts at BASELINE configs 3 and 4 (bounded): wall time, outcome counts.
python tools/bench_hunt.py [cfg3|cfg4|all] [B3] [B4] [gmres_cluster_mode]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402
from bench_shapes import seeds  # noqa: E402


def run(tag, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = r.status.cpu().numpy()
    rec = {"tag": tag, "wall_s": dt, "status": {int(k): int((st == k).sum()) for k in np.unique(st)},
           "median_cost": float(r.costs[-1].median()), "nit": [int(x.sum()) if torch.is_tensor(x) else x for x in (r.nit if isinstance(r.nit, list) else [r.nit])]}
    print(json.dumps(rec), flush=True)
    return r


which = sys.argv[1] if len(sys.argv) > 1 else "all"
B3 = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B4 = int(sys.argv[3]) if len(sys.argv) > 3 else 64
if len(sys.argv) > 4:  # GMRES step kernel selection (ks_debug_gmres_cluster)
    from orbithunter_b200 import _lib
    _lib.lib().ks_debug_gmres_cluster(int(sys.argv[4]))
if which in ("cfg3", "all"):
    for cls in ("ShiftReflectionOrbitKS", "AntisymmetricOrbitKS", "RelativeOrbitKS"):
        u = BatchedOrbitKS(seeds(cls, 64, 64, 60.0, 44.0, B3), np.tile([60.0, 44.0, 0.0], (B3, 1)), cls)
        hunt(u, "adj", maxiter=10)
        r = run(f"cfg3 {cls} 64x64 B={B3} adj 2000", lambda: hunt(u, "adj", maxiter=2000))
        run(f"cfg3 {cls} 64x64 B={B3} gmres(50 x2) maxiter=5", lambda: hunt(r.orbit, "gmres", tol=1e-10, maxiter=5,
                                                                      scipy_kwargs={"restart": 50, "rtol": 1e-6, "maxiter": 2}))
if which in ("cfg4", "all"):
    u = BatchedOrbitKS(seeds("OrbitKS", 256, 256, 200.0, 200.0, B4), np.tile([200.0, 200.0, 0.0], (B4, 1)), "OrbitKS")
    hunt(u, "gmres", maxiter=1, scipy_kwargs={"restart": 5, "rtol": 1e-2, "maxiter": 1})
    run(f"cfg4 OrbitKS 256x256 B={B4} gmres(50) maxiter=2", lambda: hunt(u, "gmres", tol=1e-10, maxiter=2,
                                                                     scipy_kwargs={"restart": 50, "rtol": 1e-6, "maxiter": 2}))
