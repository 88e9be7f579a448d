"""ncu target: a few 32x32 matvec / rmatvec launches at B = 65 536 (python tools/profile_x1_vops.py under ncu -k regex:ks_warp32x1)."""
import numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS
u = BatchedOrbitKS.from_seeds(np.arange(65536), (44.0, 44.0, 0.0), "OrbitKS", (32, 32))
v = BatchedOrbitKS.from_seeds(np.arange(65536) + 7, (44.0, 44.0, 0.0), "OrbitKS", (32, 32))
for _ in range(3):
    u.matvec(v); u.rmatvec(v)
torch.cuda.synchronize()
