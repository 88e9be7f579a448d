"""Tile path with and without TMA staging: results must agree bit for bit (same arithmetic), timing per evaluation.

    python tools/bench_tile_tma.py [--shapes 256x256x64,128x128x256] [--modes 0,1,2,3] [--cls 0]
One JSON line per (shape, op, mode)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS, _lib  # noqa: E402
from orbithunter_b200.batch import CLASS_NAMES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="256x256x64,128x128x256,64x256x128,256x128x96")
    ap.add_argument("--modes", default="0,1,2,3")
    ap.add_argument("--cls", default="0")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--chunk-mb", type=int, default=256)
    args = ap.parse_args()
    L = _lib.lib()
    L.ks_debug_tile(2, args.chunk_mb)  # multi-pass tiles for every shape (also 64x64)
    for cid in [int(c) for c in args.cls.split(",")]:
        for shp in args.shapes.split(","):
            n, m, B = (int(v) for v in shp.split("x"))
            T, Lx = (200.0, 200.0) if max(n, m) >= 256 else (60.0, 44.0)
            u = BatchedOrbitKS.from_seeds(np.arange(B), (T, Lx, 0.3 if cid == 1 else 0.0), CLASS_NAMES[cid], (n, m))
            v = BatchedOrbitKS.from_seeds(np.arange(B) + 500, (T, Lx, 0.0), CLASS_NAMES[cid], (n, m))
            v.parameters[:, :] = 0.1
            ops = {"costgrad": lambda: torch.cat([u.costgrad()[1].state.ravel(), u.costgrad()[0], u.costgrad()[1].parameters.ravel()]),
                   "eqn": lambda: u.eqn().state.ravel(), "matvec": lambda: u.matvec(v).state.ravel(),
                   "rmatvec": lambda: torch.cat([u.rmatvec(v).state.ravel(), u.rmatvec(v).parameters.ravel()])}
            timed = {"costgrad": lambda: u.costgrad(), "eqn": lambda: u.eqn(), "matvec": lambda: u.matvec(v), "rmatvec": lambda: u.rmatvec(v)}
            ref = {}
            for mode in [int(x) for x in args.modes.split(",")]:
                L.ks_debug_tile_tma(mode)
                for op, fn in ops.items():
                    out = fn()
                    torch.cuda.synchronize()
                    if mode == 0 or op not in ref:
                        ref.setdefault(op, out.clone())
                    diff = float((out - ref[op]).abs().max() / ref[op].abs().max())
                    for _ in range(3):
                        timed[op]()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        timed[op]()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.iters
                    print(json.dumps({"cls": CLASS_NAMES[cid], "grid": [n, m], "batch": B, "op": op, "tma_mode": mode, "ms": ms,
                                      "per_s": B / ms * 1e3, "max_rel_diff_vs_mode0": diff, "finite": bool(torch.isfinite(out).all())}), flush=True)
    L.ks_debug_tile_tma(0)
    L.ks_debug_tile(1, 256)


if __name__ == "__main__":
    main()
