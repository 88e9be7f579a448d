"""Profiling helper for the large-grid tile path: repeated fused costgrad on one shape.
python tools/profile_tile.py CLASS N M BATCH [iters] [chunk_mb]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS, _lib  # noqa: E402
from bench_shapes import seeds, timeit  # noqa: E402

if __name__ == "__main__":
    cls, n, m, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    if len(sys.argv) > 6:
        _lib.lib().ks_debug_tile(1, int(sys.argv[6]))
    T, L = (200.0, 200.0) if n >= 256 else (60.0, 44.0)
    u = BatchedOrbitKS(seeds(cls, n, m, T, L, B), np.tile([T, L, 0.0], (B, 1)), cls)
    ms = timeit(lambda: u.costgrad(), iters)
    nm = u.state.shape[1] * u.state.shape[2]
    print(f"{cls} {n}x{m} B={B}: {ms * 1e3:.1f} us/eval-batch, {B / ms * 1e3:.4g} evals/s, alg {B * 8 * (2 * nm + 7) / ms / 1e6:.1f} GB/s")
