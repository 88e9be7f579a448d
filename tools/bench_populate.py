"""Side measurement: seeds/s of the device seed generator (BatchedOrbitKS.from_seeds) vs the host generator (the product's
NumPy populate = the reference's algorithm) on one core.  python tools/bench_populate.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS, OrbitKS  # noqa: E402

for (n, m, B) in ((32, 32, 65536), (64, 64, 16384), (256, 256, 512)):
    seeds = np.arange(B)
    BatchedOrbitKS.from_seeds(seeds[:64], (44.0, 33.4, 0.0), "OrbitKS", (n, m))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = BatchedOrbitKS.from_seeds(seeds, (44.0, 33.4, 0.0), "OrbitKS", (n, m))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    b = BatchedOrbitKS.from_seeds(seeds, (44.0, 33.4, 0.0), "OrbitKS", (n, m))
    e1.record()
    torch.cuda.synchronize()
    nh = 200 if n <= 64 else 8
    t0 = time.perf_counter()
    for s in range(nh):
        OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=s, discretization=(n, m))
    th = (time.perf_counter() - t0) / nh
    print(json.dumps({"grid": [n, m], "seeds": B, "device_wall_s": dt, "device_event_ms": e0.elapsed_time(e1),
                      "device_seeds_per_s": B / dt, "host_us_per_seed_one_core": th * 1e6, "host_seeds_per_s_one_core": 1 / th}), flush=True)
