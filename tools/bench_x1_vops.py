"""32x32 matvec / rmatvec on the two-orbits-per-warp kernel (variant 3, the default server of these ops) vs the
one-orbit-per-warp kernel (variant 7), and the dense Newton stage with the Jacobian built by that kernel's Jacobian mode vs a
batched matvec over a stored identity.

    python tools/bench_x1_vops.py [--batch 65536] [--dense-batch 2048]"""
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


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--dense-batch", type=int, default=2048)
    args = ap.parse_args()
    L = _lib.lib()
    for cls in ("OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS"):
        u = BatchedOrbitKS.from_seeds(np.arange(args.batch), (44.0, 44.0, 0.0), cls, (32, 32))
        v = BatchedOrbitKS.from_seeds(np.arange(args.batch) + 7, (44.0, 44.0, 0.0), cls, (32, 32))
        for variant in (3, 7, 3, 7):
            L.ks_debug_warp32_variant(variant)
            mv = timed(lambda: u.matvec(v), 20)
            rmv = timed(lambda: u.rmatvec(v), 20)
            print(json.dumps({"class": cls, "batch": args.batch, "variant": variant, "matvec_M_per_s": args.batch / mv / 1e3,
                              "rmatvec_M_per_s": args.batch / rmv / 1e3}), flush=True)
        L.ks_debug_warp32_variant(4)
    u = BatchedOrbitKS.from_seeds(np.arange(args.dense_batch), (44.0, 44.0, 0.0), "OrbitKS", (32, 32))
    w = hunt(u, "adj", maxiter=10000).orbit
    for on in (2, 1, 0, 2, 1, 0):
        L.ks_debug_dense_jacobian_mode(on)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = hunt(w, "lstsq", maxiter=200, tol=1e-10)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"dense_batch": args.dense_batch, "jacobian_mode": on, "dense_stage_s": dt,
                          "converged": int((r.status == 1).sum()), "dense_solves": int(r.dense_solves)}), flush=True)
    L.ks_debug_dense_jacobian_mode(2)
    # Jacobian alone
    for on in (1,):
        t = timed(lambda: w.jacobian() if w.batch <= 256 else BatchedOrbitKS(w.state[:256].contiguous(), w.parameters[:256].contiguous(),
                                                                           w.cls_id, "modes", (32, 32), w.constraints).jacobian(), 5)
        print(json.dumps({"jacobian_256_orbits_ms": t}), flush=True)


if __name__ == "__main__":
    main()
