"""Tuning tool: 32x32 fused costgrad through each kernel variant at several batch sizes, plain launches and CUDA-graph replay.

    python tools/bench_variants.py [--variants 4,5] [--batches 4096,65536] [--iters 200]
One JSON line per (variant, batch).  Inputs are device-generated seeds; sets are rotated so the working set exceeds L2.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbithunter_b200 import BatchedOrbitKS, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="4,5")
    ap.add_argument("--batches", default="4096,8192,16384,65536")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cls", type=int, default=0)
    ap.add_argument("--staggers", default="0")
    ap.add_argument("--pingpong", default="1", help="variant 6: comma list of 0/1")
    args = ap.parse_args()
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    cls_name = ["OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS"][args.cls]
    for B in [int(x) for x in args.batches.split(",")]:
        nsets = max(2, int(np.ceil(260e6 / (B * 15000.0))))
        sets = []
        for s in range(nsets):
            orb = BatchedOrbitKS.from_seeds(np.arange(B) + s * B, (44.0, 33.4, 0.0), cls_name, (32, 32), device=dev)
            sets.append((orb, torch.empty(B, dtype=torch.float64, device=dev), torch.empty_like(orb.state),
                         torch.empty_like(orb.parameters)))
        nm = sets[0][0].state.shape[1] * sets[0][0].state.shape[2]
        wsb = L.ks_workspace_bytes(args.cls, 32, 32, B)
        ws = torch.empty(max(wsb, 1024), dtype=torch.uint8, device=dev)

        def step(i):
            o, c, g, gp = sets[i % nsets]
            _lib.check(L.ks_costgrad(args.cls, 32, 32, B, o.state.data_ptr(), o.parameters.data_ptr(), o.cmask, 0, c.data_ptr(),
                                     g.data_ptr(), gp.data_ptr(), None, ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream))

        ref = None
        for v, stg, pp in [(int(x), int(y), int(z)) for x in args.variants.split(",") for y in args.staggers.split(",")
                           for z in (args.pingpong.split(",") if int(x) >= 5 else ["0"])]:
            L.ks_debug_warp32_variant(v)
            L.ks_debug_warp32_pingpong(pp)
            L.ks_debug_warp32_stagger(stg)
            for i in range(2 * nsets):
                step(i)
            torch.cuda.synchronize()
            chk = float(sets[0][1].sum()), float(sets[0][2].abs().sum())
            if ref is None:
                ref = chk
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            ms_plain = e0.elapsed_time(e1) / args.iters
            # graph replay of nsets * 8 launches
            n_g = nsets * 8
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                for i in range(nsets):
                    step(i)
                side.synchronize()
                with torch.cuda.graph(g, stream=side):
                    for i in range(n_g):
                        step(i)
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
            reps = max(1, args.iters // n_g)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms_graph = e0.elapsed_time(e1) / (reps * n_g)
            print(json.dumps({"variant": v, "stagger_ns": stg, "pingpong": pp, "cls": cls_name, "batch": B, "sets": nsets, "us_plain": 1e3 * ms_plain,
                              "us_graph": 1e3 * ms_graph, "Mevals_plain": B / ms_plain / 1e3, "Mevals_graph": B / ms_graph / 1e3,
                              "hbm_frac_graph": B * 8 * (2 * nm + 7) / (ms_graph * 1e-3) / 6451.2e9,
                              "rel_diff_vs_first": [abs(chk[0] - ref[0]) / abs(ref[0]), abs(chk[1] - ref[1]) / abs(ref[1])]}),
                  flush=True)
        L.ks_debug_warp32_variant(4)
        L.ks_debug_warp32_stagger(0)
        L.ks_debug_warp32_pingpong(0)
        del sets


if __name__ == "__main__":
    main()
