"""Newton-GMRES(50) at BASELINE configs[2] (4096 orbits of a 64x64 class, adjoint warm-up first) with and without the operator
cache of the Krylov driver (ks_debug_operator_cache): wall time of the GMRES stage, identical results required.

    python tools/bench_opcache.py [--batch 4096]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS, _lib  # noqa: E402
from orbithunter_b200.optimize import hunt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    args = ap.parse_args()
    L = _lib.lib()
    for cls in ("RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS"):
        u = BatchedOrbitKS.from_seeds(np.arange(args.batch), (60.0, 44.0, 0.0), cls, (64, 64))
        r1 = hunt(u, "adj", maxiter=2000)
        ref = None
        for on in (1, 0, 1):
            L.ks_debug_operator_cache(on)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r2 = hunt(r1.orbit, "gmres", maxiter=3, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-6})
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            chk = (r2.orbit.state.clone(), r2.costs[-1].clone())
            same = True if ref is None else bool(torch.equal(chk[0], ref[0]) and torch.equal(chk[1], ref[1]))
            ref = ref or chk
            print(json.dumps({"class": cls, "batch": args.batch, "operator_cache": on, "newton_gmres50_s": dt,
                              "operator_applications": int(r2.operator_applications), "identical_to_first_run": same}), flush=True)
    L.ks_debug_operator_cache(0)


if __name__ == "__main__":
    main()
