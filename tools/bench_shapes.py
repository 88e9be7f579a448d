"""Side measurements (not the bench.py contract): evals/s of the fused costgrad and of matvec/rmatvec at the other
BASELINE shapes, through BatchedOrbitKS, CUDA-event timed.  python tools/bench_shapes.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from orbithunter_b200 import BatchedOrbitKS  # noqa: E402
from orbithunter_b200.ks import CLASSES  # noqa: E402


def seeds(cls, n, m, T, L, B):
    base = np.stack([CLASSES[cls](parameters=(T, L, 0.0)).populate(attr="state", seed=s, discretization=(n, m)).state
                     for s in range(min(B, 16))])
    reps = (B + base.shape[0] - 1) // base.shape[0]
    return np.tile(base, (reps, 1, 1))[:B] * (1.0 + 1e-3 * np.arange(B)[:, None, None] / B)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    from orbithunter_b200 import _lib  # noqa: E402

    chunks = [int(a) for a in sys.argv[1:]] or [64]
    out = []
    for chunk_mb in chunks:
      _lib.lib().ks_debug_tile(1, chunk_mb)
      for cls, n, m, T, L, B in (("OrbitKS", 32, 32, 44.0, 33.4, 4096), ("RelativeOrbitKS", 64, 64, 60.0, 44.0, 4096),
                               ("ShiftReflectionOrbitKS", 64, 64, 60.0, 44.0, 4096), ("AntisymmetricOrbitKS", 64, 64, 60.0, 44.0, 4096),
                               ("OrbitKS", 128, 128, 100.0, 100.0, 1024), ("OrbitKS", 256, 256, 200.0, 200.0, 64),
                               ("OrbitKS", 256, 256, 200.0, 200.0, 512)):
          if chunk_mb != chunks[0] and n == 32:
              continue
          M = seeds(cls, n, m, T, L, B)
          p = np.tile([T, L, 0.0], (B, 1))
          u = BatchedOrbitKS(M, p, cls)
          v = BatchedOrbitKS(np.random.RandomState(0).randn(*M.shape), np.random.RandomState(1).randn(B, 3), cls)
          nm = M.shape[1] * M.shape[2]
          rec = {"class": cls, "grid": [n, m], "batch": B, "chunk_mb": chunk_mb}
          ms = timeit(lambda: u.costgrad(), 20)
          rec["costgrad_evals_per_s"] = B / ms * 1e3
          rec["costgrad_alg_GBps"] = B * 8 * (2 * nm + 7) / ms / 1e6
          ms = timeit(lambda: u.matvec(v), 20)
          rec["matvec_per_s"] = B / ms * 1e3
          ms = timeit(lambda: u.rmatvec(v), 20)
          rec["rmatvec_per_s"] = B / ms * 1e3
          rec["matvec_alg_GBps"] = B * 8 * (3 * nm + 9) / ms / 1e6
          out.append(rec)
          print(json.dumps(rec), flush=True)
