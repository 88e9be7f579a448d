"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE (container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

Everything written here is an output of chaotic-systems/orbithunter v1.3.4 itself (imported via oracle/refshim.py),
never of this repository's code, so the golden files pin both the NumPy oracle (tests/test_oracle_golden.py, CPU)
and the CUDA path (tests/test_gpu_parity.py, -m gpu).  Inputs come from the reference's own generator
(``populate(attr='state', seed=..)``, ks/orbits.py:1738), its test fixture (tests/test_basic.py:328-343) and its
stored orbits (tests/test_data.h5, read with oracle/h5mini.py).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from oracle.h5mini import H5Mini  # noqa: E402

warnings.filterwarnings("ignore")
oh = refshim.load()
GOLD = os.path.join(ROOT, "tests", "golden")
CLASSES = {
    0: oh.OrbitKS,
    1: oh.RelativeOrbitKS,
    2: oh.ShiftReflectionOrbitKS,
    3: oh.AntisymmetricOrbitKS,
}
# reference tests/test_basic.py:328-343
FIXED6 = np.array(
    [
        [1.76405235, 0.40015721, 0.97873798, 2.2408932, 1.86755799, -0.97727788],
        [0.95008842, -0.15135721, -0.10321885, 0.4105985, 0.14404357, 1.45427351],
        [0.76103773, 0.12167502, 0.44386323, 0.33367433, 1.49407907, -0.20515826],
        [0.3130677, -0.85409574, -2.55298982, 0.6536186, 0.8644362, -0.74216502],
        [2.26975462, -1.45436567, 0.04575852, -0.18718385, 1.53277921, 1.46935877],
        [0.15494743, 0.37816252, -0.88778575, -1.98079647, -0.34791215, 0.15634897],
    ]
)


def cons_tuple(o):
    return np.array([bool(o.constraints[k]) for k in "txs"])


def evaluate(o, v, vparams, out, key, sub=None):
    """All hot-path outputs of the reference for orbit ``o`` (modes basis) and vector ``v``."""
    C = o.__class__
    take = (lambda a: a) if sub is None else (lambda a: a[sub])
    vo = C(**{**vars(o), "state": v, "parameters": tuple(float(x) for x in vparams)})
    F = o.eqn()
    out[key + "/params"] = np.array(o.parameters, dtype=float)
    out[key + "/constraints"] = cons_tuple(o)
    out[key + "/vparams"] = np.array(vparams, dtype=float)
    if sub is None:
        out[key + "/modes"] = o.state
        out[key + "/v"] = v
        out[key + "/field"] = o.transform(to="field").state
        out[key + "/smodes"] = o.transform(to="spatial_modes").state
    out[key + "/modes_norm"] = np.array(o.norm())
    out[key + "/eqn"] = take(F.state)
    out[key + "/eqn_norm"] = np.array(F.norm())
    out[key + "/cost"] = np.array(o.cost())
    mv = o.matvec(vo)
    out[key + "/matvec"] = take(mv.state)
    out[key + "/matvec_norm"] = np.array(mv.norm())
    rm = o.rmatvec(vo)
    out[key + "/rmatvec"] = take(rm.state)
    out[key + "/rmatvec_norm"] = np.array(rm.norm())
    out[key + "/rmatvec_params"] = np.array(rm.parameters, dtype=float)
    for pc in (False, True):
        g = o.costgrad(F, preconditioning=pc)
        tag = "/costgrad_pc" if pc else "/costgrad"
        out[key + tag] = take(g.state)
        out[key + tag + "_norm"] = np.array(g.norm())
        out[key + tag + "_params"] = np.array(g.parameters, dtype=float)


def main():
    os.makedirs(GOLD, exist_ok=True)
    out = {}
    # ---- 1. the reference's 6x6 fixture, params (44, 44, 5): tests/test_basic.py:525-528, :622-702 -------------
    for cid, C in CLASSES.items():
        f = C(state=FIXED6, basis="field", parameters=(44, 44, 5))
        s = f.transform(to="spatial_modes")
        mo = s.transform(to="modes")
        sb = mo.transform(to="spatial_modes")
        fb = sb.transform(to="field")
        key = f"fixed6/{cid}"
        out[key + "/transform_norms"] = np.array([f.norm(), s.norm(), mo.norm(), sb.norm(), fb.norm()])
        out[key + "/field_in"] = FIXED6
        out[key + "/field_back"] = fb.state
        evaluate(mo, mo.state, np.array(mo.parameters), out, key)
        # the reference's own known answers use v = self (rmatvec/matvec) and v = eqn (costgrad)
        out[key + "/rmatvec_self_norm"] = np.array(mo.rmatvec(mo).norm())
        out[key + "/matvec_self_norm"] = np.array(mo.matvec(mo).norm())
        out[key + "/costgrad_eqn_norm"] = np.array(mo.costgrad(mo.eqn()).norm())

    # ---- 2. stored converged orbits of the in-scope classes (tests/test_data.h5) ----------------------------------
    h5 = H5Mini(os.path.join(refshim.REFERENCE_ROOT, "tests", "test_data.h5")).datasets()
    for name in ("rpo", "defect", "large_defect", "wiggle"):
        arr, attrs = h5[name + "/0"]
        C = getattr(oh, attrs["class"])
        cid = [k for k, v in CLASSES.items() if v is C][0]
        kw = {"frame": attrs["frame"]} if "frame" in attrs else {}
        o = C(
            state=arr,
            basis=attrs["basis"],
            parameters=tuple(float(x) for x in attrs["parameters"]),
            discretization=tuple(int(x) for x in attrs["discretization"]),
            **kw,
        ).transform(to="modes")
        key = f"stored/{name}"
        out[key + "/cls"] = np.array(cid)
        out[key + "/stored_basis_is_field"] = np.array(attrs["basis"] == "field")
        out[key + "/stored_state"] = arr
        evaluate(o, o.state, np.array(o.parameters), out, key)
        out[key + "/rmatvec_self_norm"] = np.array(o.rmatvec(o).norm())
        out[key + "/matvec_self_norm"] = np.array(o.matvec(o).norm())

    # ---- 3. random seeds from the reference's generator at the BASELINE shapes (+ one non-2^k, one rectangular) ----
    rng = np.random.RandomState(12345)
    shapes = {0: [(32, 32), (64, 64), (26, 34), (16, 128)], 1: [(32, 32), (64, 64)], 2: [(32, 32), (64, 64), (34, 26)],
              3: [(32, 32), (64, 64)]}
    for cid, C in CLASSES.items():
        for (n, m) in shapes[cid]:
            for variant, cons in (("default", None), ("tfixed", {"t": True}), ("xfixed", {"x": True})):
                if variant != "default" and (n, m) != (32, 32):
                    continue
                T, L = (44.0, 33.4) if (n, m) == (32, 32) else (60.0, 44.0)
                S = 1.7 if cid == 1 else 0.0
                o = C(parameters=(T, L, 0.0)).populate(attr="state", seed=n + 3 * m + cid, discretization=(n, m))
                constraints = dict(o.constraints)
                if cons:
                    constraints.update(cons)
                o = C(**{**vars(o), "parameters": (T, L, S), "constraints": constraints})
                v = rng.randn(*o.state.shape)
                vparams = rng.randn(3)
                key = f"random/{cid}_{n}x{m}_{variant}"
                out[key + "/seed"] = np.array(n + 3 * m + cid)
                evaluate(o, v, vparams, out, key)
    np.savez_compressed(os.path.join(GOLD, "ks_eval.npz"), **out)

    # ---- 4. large domain OrbitKS 256x256 (BASELINE config 4): sub-sampled outputs, inputs regenerated from seed ----
    big = {}
    n = m = 256
    T, L = 200.0, 200.0
    o = oh.OrbitKS(parameters=(T, L, 0.0)).populate(attr="state", seed=7, discretization=(n, m))
    rs = np.random.RandomState(99)
    v = rs.randn(*o.state.shape)
    vparams = rs.randn(3)
    sub = (slice(None, None, 7), slice(None, None, 5))
    big["big/seed"] = np.array(7)
    big["big/vseed"] = np.array(99)
    big["big/shape"] = np.array([n, m])
    big["big/sub"] = np.array([7, 5])
    evaluate(o, v, vparams, big, "big", sub=sub)
    big["big/modes_sub"] = o.state[sub]
    np.savez_compressed(os.path.join(GOLD, "ks_eval_256.npz"), **big)

    # ---- 5. hunt() outcomes: adjoint descent and Newton-Krylov on reference seeds -------------------------------
    hunts = {}
    for cid, C in CLASSES.items():
        o = C(parameters=(44.0, 44.0, 0.0)).populate(attr="state", seed=3, discretization=(32, 32))
        key = f"hunt/{cid}"
        hunts[key + "/modes"] = o.state
        hunts[key + "/params"] = np.array(o.parameters, dtype=float)
        for pc in (False, True):
            r = oh.hunt(o, "adj", maxiter=300, preconditioning=pc)
            k2 = key + ("/adj300_pc" if pc else "/adj300")
            hunts[k2 + "/state"] = r.orbit.state
            hunts[k2 + "/params"] = np.array(r.orbit.parameters, dtype=float)
            hunts[k2 + "/status_nit"] = np.array([r.status, r.nit])
            hunts[k2 + "/costs"] = np.array(r.costs, dtype=float)
        for method, kw in (("gmres", {"restart": 20, "maxiter": 2}), ("lsqr", {"iter_lim": 50})):
            r = oh.hunt(o, method, maxiter=3, scipy_kwargs=dict(kw))
            k2 = key + "/" + method
            hunts[k2 + "/state"] = r.orbit.state
            hunts[k2 + "/params"] = np.array(r.orbit.parameters, dtype=float)
            hunts[k2 + "/status_nit"] = np.array([r.status, r.nit])
            hunts[k2 + "/costs"] = np.array(r.costs, dtype=float)
    # a converging run: the stored rpo perturbed (BASELINE config 1 "also run the stored rpo perturbed by 1e-3 randn")
    arr, attrs = h5["rpo/0"]
    rs = np.random.RandomState(5)
    o = oh.RelativeOrbitKS(state=arr + 1e-3 * rs.randn(*arr.shape), basis="modes",
                           parameters=tuple(float(x) for x in attrs["parameters"]), discretization=(32, 32))
    r = oh.hunt(o, "adj", maxiter=20000, tol=1e-5)
    hunts["hunt/rpo_adj/modes"] = o.state
    hunts["hunt/rpo_adj/params"] = np.array(o.parameters, dtype=float)
    hunts["hunt/rpo_adj/state"] = r.orbit.state
    hunts["hunt/rpo_adj/out_params"] = np.array(r.orbit.parameters, dtype=float)
    hunts["hunt/rpo_adj/status_nit"] = np.array([r.status, r.nit])
    hunts["hunt/rpo_adj/costs"] = np.array(r.costs, dtype=float)
    r2 = oh.hunt(o, "gmres", maxiter=20, tol=1e-8, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})
    hunts["hunt/rpo_gmres/state"] = r2.orbit.state
    hunts["hunt/rpo_gmres/out_params"] = np.array(r2.orbit.parameters, dtype=float)
    hunts["hunt/rpo_gmres/status_nit"] = np.array([r2.status, r2.nit])
    hunts["hunt/rpo_gmres/costs"] = np.array(r2.costs, dtype=float)
    # OrbitKS seeds 0..7 at BASELINE config 2 settings but 1500 steps: per-seed status / cost / nit table
    tab = []
    for seed in range(8):
        o = oh.OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=seed, discretization=(32, 32))
        r = oh.hunt(o, "adj", maxiter=1500)
        tab.append([seed, r.status, r.nit, r.orbit.cost(), r.orbit.t, r.orbit.x])
    hunts["hunt/seeds_adj1500"] = np.array(tab, dtype=float)
    np.savez_compressed(os.path.join(GOLD, "ks_hunt.npz"), **hunts)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
