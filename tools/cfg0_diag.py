"""Diagnostic behind tests/test_gpu_r2.py::test_cfg0_status_table_matches_reference: BASELINE configs[0] (adj 10000 -> lstsq 200)
on the seeds of tests/golden/ks_cfg0_table.npz (the unmodified reference's outcomes), for several chunk / refinement settings of
the dense step; prints the converged sets, their differences and the exit-code confusion matrix.

    python tools/cfg0_diag.py      (on a GPU box, from the repository root)
"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
from orbithunter_b200 import BatchedOrbitKS
from orbithunter_b200.optimize import hunt
tab = np.load("tests/golden/ks_cfg0_table.npz")["table"]
seeds = tab[:, 0].astype(int)
b = BatchedOrbitKS.from_seeds(seeds, (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
r1 = hunt(b, "adj", maxiter=10000, tol=1e-6)
print("adj status equal", np.array_equal(r1.status.cpu().numpy(), tab[:, 1].astype(int)), "nit equal", np.array_equal(r1.nit.cpu().numpy(), tab[:, 2].astype(int)),
      "cost maxrel", float(np.max(np.abs(r1.costs[-1].cpu().numpy() - tab[:, 3]) / tab[:, 3])))
for chunk, refine in ((256, 2), (64, 2), (256, 4)):
    r2 = hunt(r1.orbit, "lstsq", maxiter=200, tol=1e-10, chunk=chunk, refine=refine)
    got, ref = r2.status.cpu().numpy(), tab[:, 4].astype(int)
    conf = {f"{a}{c}": int(np.sum((ref == a) & (got == c))) for a in (0, 1, 2, 3) for c in (0, 1, 2, 3)}
    conv = ref == 1
    P = r2.orbit.cpu_parameters()
    both = conv & (got == 1)
    print(json.dumps({"chunk": chunk, "refine": refine, "ref_conv": int(conv.sum()), "got_conv": int((got == 1).sum()), "both": int(both.sum()),
                      "ref_only": seeds[conv & (got != 1)].tolist(), "got_only": seeds[(~conv) & (got == 1)].tolist(), "agree": float(np.mean(got == ref)),
                      "T_maxrel_both": float(np.max(np.abs(P[both, 0] - tab[both, 7]) / tab[both, 7])) if both.any() else None, "confusion": conf}))
