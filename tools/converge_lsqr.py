import json, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from orbithunter_b200 import BatchedOrbitKS
from orbithunter_b200.optimize import hunt
B = 1024
u = BatchedOrbitKS.from_seeds(np.arange(B), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
r0 = hunt(u, "adj", maxiter=10000)
for nk, lim in ((60, 300), (200, 300), (60, 1000), (200, 1000)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = hunt(r0.orbit, "lsqr", maxiter=nk, tol=1e-10, scipy_kwargs={"iter_lim": lim})
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = r.status.cpu().numpy()
    print(json.dumps({"newton": nk, "iter_lim": lim, "wall_s": round(dt, 3), "status": {int(k): int((st == k).sum()) for k in np.unique(st)}, "conv_per_s": round(float((st == 1).sum()) / dt, 2)}), flush=True)
