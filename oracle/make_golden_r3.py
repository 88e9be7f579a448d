"""Round-2 (second batch) golden vectors from the UNMODIFIED reference -- TEST INFRASTRUCTURE (container only).

    python oracle/make_golden_r3.py

Writes tests/golden/ks_aux_r3.npz: seed generation with every modulation profile / xshift (ks/orbits.py:1738-1849), the
tutorial's catalog seed (notebooks/tutorial.ipynb: gaussian x truncate, rescale(5); cost 8807.937... as in
notebooks/data/catalog/numerical_method_benchmarks.csv:2), Orbit.rescale (core.py:1831-1853), resize (core.py:1167-1226),
dimension_based_discretization (ks/orbits.py:1638-1694) and the norms preprocess thresholds (:1566-1596, :2969-3008).
Everything stored is an output of chaotic-systems/orbithunter v1.3.4 itself (oracle/refshim.py).
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
GOLD = os.path.join(ROOT, "tests", "golden")
CLASSES = {0: oh.OrbitKS, 1: oh.RelativeOrbitKS, 2: oh.ShiftReflectionOrbitKS, 3: oh.AntisymmetricOrbitKS}

SPATIAL = ("gaussian", "laplace", "laplace_sqrt", "plateau_linear", "exponential_linear", "flat_top", "none")
TEMPORAL = ("gaussian", "laplace", "flat_top", "laplace_plateau", "truncate", "none")
XSHIFT = ("sqrt", "log", "lin", "none")


def main():
    out = {}
    # --- modulations: every spatial profile with the default temporal one and vice versa, every xshift; two grids, 4 classes
    combos = [(s, "gaussian", "sqrt") for s in SPATIAL] + [("laplace", t, "sqrt") for t in TEMPORAL] + [("gaussian", "truncate", x) for x in XSHIFT]
    for cid, C in CLASSES.items():
        for (n, m) in ((32, 32), (16, 24)):
            for (sm, tm, xs) in combos:
                o = C(parameters=(52.0, 41.0, 0.0)).populate(attr="state", seed=5 + cid, discretization=(n, m), spatial_modulation=sm,
                                                            temporal_modulation=tm, xshift=xs)
                out[f"pop/{cid}_{n}x{m}/{sm}/{tm}/{xs}"] = o.state
    o = oh.OrbitKS(parameters=(52.0, 41.0, 0.0)).populate(attr="state", seed=9, discretization=(32, 32), tscale=3, xscale=4, xvar=2.5,
                                                        tvar=1.5, spatial_modulation="gaussian", temporal_modulation="laplace")
    out["pop/override"] = o.state
    # --- the catalog seed of the reference's only published timing
    np.random.seed(0)
    o = oh.OrbitKS().populate(seed=0, attr="all", discretization=(64, 64), spatial_modulation="gaussian",
                              temporal_modulation="truncate").rescale(5)
    out["catalog/state"] = o.state
    out["catalog/params"] = np.array(o.parameters, dtype=float)
    out["catalog/cost"] = np.array(o.cost())
    out["catalog/basis"] = np.array(o.basis)
    print("catalog seed", o.parameters, o.cost(), o.basis)
    # --- rescale, all methods, from modes and from field
    for cid, C in CLASSES.items():
        o = C(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=21 + cid, discretization=(32, 32))
        out[f"rescale/{cid}/in"] = o.state
        for method, mag in (("inf", 5.0), ("L1", 7.0), ("L2", 3.0), ("LP", 1.5)):
            out[f"rescale/{cid}/{method}"] = o.rescale(mag, method=method).state
            out[f"rescale/{cid}/{method}_field"] = o.transform(to="field").rescale(mag, method=method).state
    # --- resize: pad / truncate in both axes at once, and resize() without a shape
    for cid, C in CLASSES.items():
        o = C(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=31 + cid, discretization=(32, 32))
        out[f"resize/{cid}/in"] = o.state
        for (n2, m2) in ((48, 40), (16, 24), (32, 64), (24, 32), (64, 16)):
            out[f"resize/{cid}/{n2}x{m2}"] = o.resize(n2, m2).state
        big = C(parameters=(150.0, 70.0, 0.0)).populate(attr="state", seed=41 + cid, discretization=(32, 32))
        r = big.resize()
        out[f"resize/{cid}/auto_shape"] = np.array(r.discretization)
        out[f"resize/{cid}/auto_in"] = big.state
        out[f"resize/{cid}/auto"] = r.state
    # --- dimension_based_discretization table
    rows = []
    for cid, C in CLASSES.items():
        for res in ("default", "coarse", "fine", "power"):
            for T in (0.0, 7.0, 20.0, 44.0, 100.0, 333.3, 1000.0):
                for L in (0.0, 10.0, 22.0, 33.4, 64.0, 200.0, 512.5):
                    n, m = C.dimension_based_discretization((T, L), resolution=res)
                    rows.append([cid, ("default", "coarse", "fine", "power").index(res), T, L, n, m])
    out["dbd/table"] = np.array(rows, dtype=float)
    out["dbd/pinned"] = np.array(oh.OrbitKS.dimension_based_discretization((44.0, 33.4), discretization=(20, None)), dtype=float)
    # --- preprocess: the norms it thresholds and its decision (class name of the result), incl. engineered cases
    for cid, C in CLASSES.items():
        base = C(parameters=(44.0, 33.4, 0.7 if cid == 1 else 0.0)).populate(attr="state", seed=51 + cid, discretization=(32, 32))
        base = C(**{**vars(base), "parameters": (44.0, 33.4, 0.7 if cid == 1 else 0.0)})
        steady = base.state.copy()
        steady[1:, :] = 0.0  # only the j = 0 row: no temporal variation
        cases = {"generic": base.state, "tiny": 1e-12 * base.state, "steady": steady, "nearly_steady": steady + 1e-12 * base.state}
        for tag, st in cases.items():
            o = C(**{**vars(base), "state": st})
            f = o.transform(to="field")
            out[f"prep/{cid}/{tag}/state"] = st
            out[f"prep/{cid}/{tag}/params"] = np.array(o.parameters, dtype=float)
            out[f"prep/{cid}/{tag}/norms"] = np.array([f.norm(), f.dt().transform(to="field").norm()])
            res = o.preprocess()
            out[f"prep/{cid}/{tag}/result_class"] = np.array(res.__class__.__name__)
            out[f"prep/{cid}/{tag}/result_params"] = np.array(res.parameters, dtype=float)
            print("prep", cid, tag, res.__class__.__name__, out[f"prep/{cid}/{tag}/norms"])
    path = os.path.join(GOLD, "ks_aux_r3.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), len(out))


if __name__ == "__main__":
    main()
