"""Golden vectors for the symmetry operations and the pairwise group-orbit gluing search, from the UNMODIFIED reference --
TEST INFRASTRUCTURE (container only).

    python oracle/make_golden_pairglue.py

Writes tests/golden/ks_pairglue.npz: OrbitKS.roll / cell_shift / reflection (ks/orbits.py:1143-1232), Orbit.concat
(core.py:1282-1380; ShiftReflection :3640-3676), group_orbit (ks/orbits.py:1355-1401: default, continuous and discrete
branches) and gluing.expensive_pairwise_glue (gluing.py:424-496, objective "cost") on field-basis orbits.  Everything stored is
an output of chaotic-systems/orbithunter v1.3.4 itself (oracle/refshim.py).
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
from orbithunter.gluing import expensive_pairwise_glue  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CLASSES = {0: oh.OrbitKS, 1: oh.RelativeOrbitKS, 2: oh.ShiftReflectionOrbitKS, 3: oh.AntisymmetricOrbitKS}


def seed_orbit(C, seed, shape, params):
    o = C(parameters=params).populate(attr="state", seed=seed, discretization=shape).transform(to="field")
    return C(**{**vars(o), "parameters": params})  # RelativeOrbitKS.populate resets S to 0: put the requested shift back


def main():
    out = {}
    for cid, C in CLASSES.items():
        a = seed_orbit(C, 3 + cid, (16, 16), (44.0, 33.4, 1.7 if cid == 1 else 0.0))
        b = seed_orbit(C, 13 + cid, (16, 16), (51.0, 29.0, -0.6 if cid == 1 else 0.0))
        out[f"{cid}/a"], out[f"{cid}/b"] = a.state, b.state
        out[f"{cid}/a_params"], out[f"{cid}/b_params"] = np.array(a.parameters, float), np.array(b.parameters, float)
        out[f"{cid}/roll"] = a.roll(3, axis=0).roll(-5, axis=1).state
        out[f"{cid}/cell_shift_t"] = a.cell_shift(2, axis=0).state
        out[f"{cid}/cell_shift_x"] = a.cell_shift(-4, axis=1).state
        r = a.reflection()
        out[f"{cid}/reflection"], out[f"{cid}/reflection_params"] = r.state, np.array(r.parameters, float)
        am = a.transform(to="modes")
        out[f"{cid}/reflection_modes"] = am.reflection().state
        for axis in (0, 1):
            c = a.concat(b, axis=axis)
            out[f"{cid}/concat{axis}"], out[f"{cid}/concat{axis}_params"] = c.state, np.array(c.parameters, float)
            out[f"{cid}/concat{axis}_disc"] = np.array(c.discretization)
        for tag, kw in (("default", dict(rolls=(8, 4))), ("continuous", dict(continuous=True, rolls=(4, 8))), ("discrete", dict(discrete=True))):
            members = list(a.group_orbit(**kw))
            out[f"{cid}/group/{tag}"] = np.stack([g.state for g in members])
            out[f"{cid}/group/{tag}_params"] = np.array([g.parameters for g in members], float)
        for axis in (0, 1):
            for tag, kw in (("default", dict(rolls=(4, 4))), ("discrete", dict(discrete=True))):
                best = expensive_pairwise_glue(np.array([a, b]).reshape((2, 1) if axis == 0 else (1, 2)), objective="cost", axis=axis, **kw)
                out[f"{cid}/pair/{tag}{axis}"] = best.state
                out[f"{cid}/pair/{tag}{axis}_params"] = np.array(best.parameters, float)
                out[f"{cid}/pair/{tag}{axis}_cost"] = np.array(best.cost())
                print(cid, axis, tag, best.state.shape, best.parameters, float(best.cost()), flush=True)
    # --- slicing and fundamental domains (core.py:553-668; ks/orbits.py:2770-2833, :3010-3014, :3229-3252, :3688-3711)
    for cid, C in CLASSES.items():
        a = C(state=out[f"{cid}/a"], basis="field", parameters=tuple(out[f"{cid}/a_params"]))
        cut = a[2:10, 4:16]
        out[f"{cid}/getitem"], out[f"{cid}/getitem_params"] = cut.state, np.array(cut.parameters, float)
        row = a[3:4, :]
        out[f"{cid}/getitem_row_params"] = np.array(row.parameters, float)
        if cid == 0:
            continue
        for half in ((0,) if cid == 1 else (0, 1)):
            kw = {} if cid == 1 else {"half": half}
            fd = a.to_fundamental_domain(**kw)
            out[f"{cid}/to_fd{half}"], out[f"{cid}/to_fd{half}_params"] = fd.state, np.array(fd.parameters, float)
            back = fd.from_fundamental_domain(**kw)
            out[f"{cid}/from_fd{half}"], out[f"{cid}/from_fd{half}_params"] = back.state, np.array(back.parameters, float)
            fdm = a.transform(to="modes").to_fundamental_domain(**kw)
            out[f"{cid}/to_fd{half}_modes"] = fdm.state
        members = list(a.group_orbit(discrete=True, fundamental_domain=True))
        out[f"{cid}/group/discrete_fd"] = np.stack([g.state for g in members])
        out[f"{cid}/group/discrete_fd_params"] = np.array([g.parameters for g in members], float)
    # --- the pairwise search over fundamental domains (gluing.py:424-496 with fundamental_domain=True)
    for cid in (2, 3):
        C = CLASSES[cid]
        a = C(state=out[f"{cid}/a"], basis="field", parameters=tuple(out[f"{cid}/a_params"]))
        b = C(state=out[f"{cid}/b"], basis="field", parameters=tuple(out[f"{cid}/b_params"]))
        for axis in (0, 1):
            for tag, kw in (("default", dict(rolls=(4, 4))), ("discrete", dict(discrete=True))):
                best = expensive_pairwise_glue(np.array([a, b]).reshape((2, 1) if axis == 0 else (1, 2)), objective="cost", axis=axis,
                                               fundamental_domain=True, **kw)
                out[f"{cid}/pairfd/{tag}{axis}"] = best.state
                out[f"{cid}/pairfd/{tag}{axis}_params"] = np.array(best.parameters, float)
                out[f"{cid}/pairfd/{tag}{axis}_cost"] = np.array(best.cost())
                print("fd", cid, axis, tag, best.state.shape, best.parameters, float(best.cost()), flush=True)
    np.savez_compressed(os.path.join(GOLD, "ks_pairglue.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
