"""Generate tests/golden/ks_family.npz by running the UNMODIFIED reference's span_family -- TEST INFRASTRUCTURE (container only).

    python oracle/make_golden_family.py

Root = the polished stored rpo of tests/golden/ks_continuation.npz ('start/*', itself an output of the reference's hunt).
span_family (continuation.py:280-384) walks t and x in steps of 0.01 inside tight bounds, each step re-converged by the
reference's hunt('gmres'); Orbit.to_h5 (h5py is absent here) is replaced by a recorder, so the golden also holds which
members the reference saves and under which names."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

warnings.filterwarnings("ignore")
oh = refshim.load()
from orbithunter.continuation import span_family  # noqa: E402

KW = dict(methods="gmres", maxiter=40, tol=1e-6, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ks_continuation.npz"))
    root = oh.RelativeOrbitKS(state=g["start/modes"], basis="modes", parameters=tuple(g["start/params"]), discretization=(32, 32))
    saved = []

    def record(self, *a, **kw):
        saved.append((kw.get("dataname"), kw.get("filename"), self.parameters))

    oh.RelativeOrbitKS.to_h5 = record
    T, L = root.t, root.x
    bounds = {"t": (T - 0.025, T + 0.025), "x": (L - 0.015, L + 0.015)}
    fam = span_family(root, bounds=bounds, step_sizes={"t": 0.01, "x": 0.01}, leafnames=True,
                      **{k: (dict(v) if isinstance(v, dict) else v) for k, v in KW.items()})
    out = {"bounds_t": np.array(bounds["t"]), "bounds_x": np.array(bounds["x"]), "n_branches": np.array(len(fam)),
           "default_filename": np.array("family_" + root.filename()), "root_filename": np.array(root.filename())}
    for b, branch in enumerate(fam):
        out[f"branch{b}/size"] = np.array(len(branch))
        for i, o in enumerate(branch):
            out[f"branch{b}/{i}/modes"] = o.state
            out[f"branch{b}/{i}/params"] = np.array(o.parameters, dtype=float)
            out[f"branch{b}/{i}/leafname"] = np.array(o.filename(cls_name=False, extension="").lstrip("_"))
        print("branch", b, [tuple(np.round(o.parameters, 4)) for o in branch])
    out["saved_names"] = np.array([s[0] for s in saved])
    out["saved_filename"] = np.array(saved[0][1])
    print(saved)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ks_family.npz"), **out)


if __name__ == "__main__":
    main()
