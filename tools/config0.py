"""BASELINE configs[0] on the device: single random OrbitKS 32x32 seeds, hunt('adj') then 'lstsq' to cost 1e-10.
python tools/config0.py [n_seeds]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import OrbitKS  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
o = OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=0, discretization=(32, 32))
hunt(o, ("adj", "lstsq"), maxiter=[10, 1])  # warm-up
for seed in range(n):
    o = OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=seed, discretization=(32, 32))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = hunt(o, ("adj", "lstsq"), maxiter=[10000, 100], tol=[1e-6, 1e-10])
    torch.cuda.synchronize()
    print(json.dumps({"seed": seed, "wall_s": round(time.perf_counter() - t0, 3), "status": r.status, "nit": list(r.nit),
                      "costs": [float(c) for c in r.costs], "T_L": [round(r.orbit.t, 4), round(r.orbit.x, 4)]}), flush=True)
