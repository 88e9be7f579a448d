"""Generate tests/golden/ks_gluing.npz by running the UNMODIFIED reference's glue / tile -- TEST INFRASTRUCTURE
(container only; needs /root/reference).      python oracle/make_golden_gluing.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

warnings.filterwarnings("ignore")
oh = refshim.load()
from orbithunter.gluing import glue, tile  # noqa: E402


def main():
    a = oh.OrbitKS(parameters=(40.0, 22.0, 0.0)).populate(attr="state", seed=1, discretization=(32, 32)).transform(to="field")
    b = oh.OrbitKS(parameters=(60.0, 26.0, 0.0)).populate(attr="state", seed=2, discretization=(32, 32)).transform(to="field")
    c = oh.OrbitKS(parameters=(30.0, 30.0, 0.0)).populate(attr="state", seed=3, discretization=(16, 32)).transform(to="field")
    out = {}
    for tag, o in (("a", a), ("b", b), ("c", c)):
        out[f"in/{tag}/field"] = o.state
        out[f"in/{tag}/params"] = np.array(o.parameters, dtype=float)

    def arr(rows):
        x = np.empty((len(rows), len(rows[0])), dtype=object)
        for i, r in enumerate(rows):
            for j, o in enumerate(r):
                x[i, j] = o
        return x

    cases = {
        "time": (arr([[a], [b]]), dict()),
        "space": (arr([[a, b]]), dict()),
        "time_strip": (arr([[a], [b]]), dict(strip_wise=True)),
        "space_strip": (arr([[a, b]]), dict(strip_wise=True)),
        "grid_strip": (arr([[a, b], [b, a]]), dict(strip_wise=True)),
        "mixed_strip": (arr([[a], [c]]), dict(strip_wise=True)),
    }
    for tag, (oa, kw) in cases.items():
        g = glue(oa, oh.OrbitKS, **kw)
        out[f"{tag}/field"] = g.transform(to="field").state
        out[f"{tag}/params"] = np.array(g.parameters, dtype=float)
        out[f"{tag}/shape"] = np.array(g.discretization)
        print(tag, g.discretization, g.parameters)
    t = tile(np.array([[0, 1], [1, 1]]), {0: a, 1: b}, oh.OrbitKS, strip_wise=True)
    out["tile/field"] = t.transform(to="field").state
    out["tile/params"] = np.array(t.parameters, dtype=float)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ks_gluing.npz"), **out)


if __name__ == "__main__":
    main()
