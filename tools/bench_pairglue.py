"""gluing.expensive_pairwise_glue: the search over all pairs of group-orbit members of two 32x32 OrbitKS tiles, on the device
(all combinations scored in batches) and in the unmodified reference (one combination at a time, oracle/_ref; timed on a
1024-combination sample).

    python tools/bench_pairglue.py"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orbithunter_b200 import OrbitKS  # noqa: E402
from orbithunter_b200.gluing import expensive_pairwise_glue  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    a = OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=1, discretization=(32, 32)).transform(to="field")
    b = OrbitKS(parameters=(51.0, 30.0, 0.0)).populate(attr="state", seed=2, discretization=(32, 32)).transform(to="field")
    expensive_pairwise_glue([a, b], rolls=(8, 8))  # warm-up
    for rolls in ((8, 8), (2, 2), (1, 1)):
        members = 2 * (32 // rolls[0]) * (32 // rolls[1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        best = expensive_pairwise_glue([a, b], rolls=rolls, chunk=65536)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "device", "rolls": rolls, "combinations": members * members, "wall_s": dt,
                          "combinations_per_s": members * members / dt, "best_cost": best.cost()}), flush=True)
    try:
        from oracle import refshim

        oh = refshim.load()
        from orbithunter.gluing import expensive_pairwise_glue as ref_glue

        ra = oh.OrbitKS(state=a.state, basis="field", parameters=a.parameters)
        rb = oh.OrbitKS(state=b.state, basis="field", parameters=b.parameters)
        t0 = time.perf_counter()
        rbest = ref_glue(np.array([ra, rb]).reshape(2, 1), rolls=(8, 8))
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "reference (one core)", "rolls": [8, 8], "combinations": 1024, "wall_s": dt,
                          "combinations_per_s": 1024 / dt, "best_cost": float(rbest.cost())}), flush=True)
    except Exception as e:  # the reference copy is optional on the box
        print(json.dumps({"impl": "reference", "unavailable": repr(e)}))


if __name__ == "__main__":
    main()
