"""CPU: pin the NumPy oracle (oracle/ks_oracle.py) to outputs of the unmodified reference (tests/golden/*.npz,
made by oracle/make_golden.py) and to the reference's own known-answer values (reference tests/test_basic.py)."""
import numpy as np
import pytest

from conftest import golden_cases, relerr
from oracle import ks_oracle as ko

TOL = 1e-12  # north_star: residual / matvec relative error <= 1e-12


def _case_cls(gold, key):
    if key.startswith("fixed6/"):
        return int(key.split("/")[1])
    if key.startswith("stored/"):
        return int(gold[key + "/cls"])
    return int(key.split("/")[1].split("_")[0])


def _all_cases(gold):
    return golden_cases(gold, "fixed6") + golden_cases(gold, "stored") + golden_cases(gold, "random")


def test_reference_known_answers(gold_eval):
    """Numbers asserted by the reference's own tests (tests/test_basic.py:390-436, :634-651)."""
    g = gold_eval
    norms = {
        0: [6.783718210674288, 5.2245770410096055, 5.050390311315469, 5.050390311315469, 5.050390311315469],
        1: [6.783718210674288, 5.2245770410096055, 5.050390311315469, 5.050390311315469, 5.050390311315469],
        2: [6.783718210674288, 5.2245770410096055, 3.635581670379944, 3.6355816703799446, 3.6355816703799437],
        3: [6.783718210674288, 5.2245770410096055, 3.821906904088631, 3.821906904088631, 3.821906904088631],
    }
    for cid, expect in norms.items():
        f = g[f"fixed6/{cid}/field_in"]
        s = ko.space_transform(f)
        mo = ko.time_transform(s, cid)
        sb = ko.inv_time_transform(mo, cid)
        fb = ko.inv_space_transform(sb)
        got = [np.linalg.norm(x) for x in (f, s, mo, sb, fb)]
        assert np.allclose(got, expect, rtol=1e-6)  # the reference's tolerance
        assert np.allclose(got, g[f"fixed6/{cid}/transform_norms"], rtol=1e-13)
    # OrbitKS fixture: rmatvec(self), costgrad(eqn), matvec(self)
    M, p = g["fixed6/0/modes"], g["fixed6/0/params"]
    rs, _ = ko.rmatvec(M, p, M, ko.ORBIT)
    assert np.linalg.norm(rs) == pytest.approx(1.295386, rel=1e-6)
    _, gs, _, _ = ko.costgrad(M, p, ko.ORBIT)
    assert np.linalg.norm(gs) == pytest.approx(0.468832, rel=1e-6)
    assert np.linalg.norm(ko.matvec(M, p, M, p, ko.ORBIT)) == pytest.approx(0.3641827310989932, rel=1e-9)
    # stored rpo (RelativeOrbitKS 32x32)
    M, p = g["stored/rpo/modes"], g["stored/rpo/params"]
    rs, _ = ko.rmatvec(M, p, M, ko.RELATIVE)
    assert np.linalg.norm(rs) == pytest.approx(60.18805016, rel=1e-8)
    assert np.linalg.norm(ko.matvec(M, p, M, p, ko.RELATIVE)) == pytest.approx(155.86156902189782, rel=1e-9)
    # the reference's unused-but-useful eqn norms (tests/test_basic.py:439-450) for the classes that still reproduce
    for cid, val in ((0, 1.258363954936136), (2, 0.8876335300535965), (3, 0.9747965276558949)):
        F = ko.eqn(g[f"fixed6/{cid}/modes"], g[f"fixed6/{cid}/params"], cid)
        assert np.linalg.norm(F) == pytest.approx(val, rel=1e-9)


def test_stored_orbits_are_solutions(gold_eval):
    """reference tests/test_basic.py:383-387: every stored orbit has cost < 1e-7."""
    g = gold_eval
    for key in golden_cases(g, "stored"):
        cid = int(g[key + "/cls"])
        stored = g[key + "/stored_state"]
        modes = ko.to_modes(stored, cid) if bool(g[key + "/stored_basis_is_field"]) else stored
        assert relerr(modes, g[key + "/modes"]) < TOL
        c = float(ko.cost(modes, g[key + "/params"], cid))
        assert c < 1e-7
        assert abs(c - float(g[key + "/cost"])) < 1e-10


def test_eval_parity_all_cases(gold_eval):
    g = gold_eval
    cases = _all_cases(g)
    assert len(cases) >= 20
    for key in cases:
        cid = _case_cls(g, key)
        M, p, v, vp = g[key + "/modes"], g[key + "/params"], g[key + "/v"], g[key + "/vparams"]
        cons = tuple(bool(x) for x in g[key + "/constraints"])
        assert relerr(ko.to_field(M, cid), g[key + "/field"]) < TOL, key
        assert relerr(ko.inv_time_transform(M, cid), g[key + "/smodes"]) < TOL, key
        assert relerr(ko.to_modes(g[key + "/field"], cid), M) < TOL, key
        # converged stored orbits: F ~ 0 by cancellation, so the error scale is that of the cancelling terms
        fscale = float(np.max(np.abs(ko.eqn_linear(M, p, cid)))) if key.startswith("stored/") else None
        assert relerr(ko.eqn(M, p, cid), g[key + "/eqn"], fscale) < TOL, key
        assert abs(float(ko.cost(M, p, cid)) - float(g[key + "/cost"])) <= 1e-12 * max(1.0, float(g[key + "/cost"])), key
        assert relerr(ko.matvec(M, p, v, vp, cid, cons), g[key + "/matvec"]) < TOL, key
        rs, rp = ko.rmatvec(M, p, v, cid, cons)
        assert relerr(rs, g[key + "/rmatvec"]) < TOL, key
        assert np.allclose(rp, g[key + "/rmatvec_params"], rtol=1e-11, atol=1e-11), key
        for pc, tag in ((False, "/costgrad"), (True, "/costgrad_pc")):
            _, gs, gp, _ = ko.costgrad(M, p, cid, cons, pc)
            gscale = None if fscale is None else fscale * float(np.max(np.abs(g[key + "/rmatvec"]))) / float(np.max(np.abs(M)))
            assert relerr(gs, g[key + tag], gscale) < TOL, key
            assert np.allclose(gp, g[key + tag + "_params"], rtol=1e-11, atol=1e-11 if gscale is None else 1e-11 * gscale), key


def test_populate_matches_reference_generator(gold_eval):
    g = gold_eval
    for key in golden_cases(g, "random"):
        cid = _case_cls(g, key)
        n, m = ko.field_shape(g[key + "/modes"], cid)
        T, L = (44.0, 33.4) if (n, m) == (32, 32) else (60.0, 44.0)
        got = ko.populate_state(T, L, n, m, int(g[key + "/seed"]), cid)
        assert relerr(got, g[key + "/modes"]) < 1e-13, key


def test_large_domain_256(gold_big):
    g = gold_big
    n, m = (int(x) for x in g["big/shape"])
    sub = (slice(None, None, int(g["big/sub"][0])), slice(None, None, int(g["big/sub"][1])))
    M = ko.populate_state(200.0, 200.0, n, m, int(g["big/seed"]), ko.ORBIT)
    assert relerr(M[sub], g["big/modes_sub"]) < 1e-13
    p, vp = g["big/params"], g["big/vparams"]
    v = np.random.RandomState(int(g["big/vseed"])).randn(*M.shape)
    F = ko.eqn(M, p, ko.ORBIT)
    assert relerr(F[sub], g["big/eqn"]) < TOL
    assert np.linalg.norm(F) == pytest.approx(float(g["big/eqn_norm"]), rel=1e-12)
    assert relerr(ko.matvec(M, p, v, vp, ko.ORBIT)[sub], g["big/matvec"]) < TOL
    rs, rp = ko.rmatvec(M, p, v, ko.ORBIT)
    assert relerr(rs[sub], g["big/rmatvec"]) < TOL
    assert np.allclose(rp, g["big/rmatvec_params"], rtol=1e-11)
    _, gs, gp, _ = ko.costgrad(M, p, ko.ORBIT, None, True)
    assert relerr(gs[sub], g["big/costgrad_pc"]) < TOL
    assert np.allclose(gp, g["big/costgrad_pc_params"], rtol=1e-10)


def test_adjoint_identity():
    """<Jv, w> = <v, J^T w> (state + free parameters), SURVEY.md section 4 (ii)."""
    rng = np.random.RandomState(0)
    for cid in range(4):
        n, m = 16, 24
        M = ko.populate_state(50.0, 40.0, n, m, 11, cid)
        p = np.array([50.0, 40.0, 1.3 if cid == ko.RELATIVE else 0.0])
        v, w, vp = rng.randn(*M.shape), rng.randn(*M.shape), rng.randn(3)
        cons = ko.default_constraints(cid)
        vp = np.where(cons, 0.0, vp)
        lhs = np.sum(ko.matvec(M, p, v, vp, cid) * w)
        rs, rp = ko.rmatvec(M, p, w, cid)
        rhs = np.sum(rs * v) + np.sum(rp * vp)
        assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)


def test_batched_equals_looped():
    rng = np.random.RandomState(3)
    for cid in range(4):
        Ms = np.stack([ko.populate_state(44.0, 33.4, 32, 32, s, cid) for s in range(3)])
        ps = np.stack([[44.0 + s, 33.4 - s, 0.5 * s if cid == 1 else 0.0] for s in range(3)])
        vs = rng.randn(*Ms.shape)
        c, gs, gp, F = ko.costgrad(Ms, ps, cid)
        mv = ko.matvec(Ms, ps, vs, ps, cid)
        for i in range(3):
            ci, gsi, gpi, Fi = ko.costgrad(Ms[i], ps[i], cid)
            assert np.allclose(F[i], Fi, rtol=0, atol=1e-13 * np.max(np.abs(Fi)))
            assert np.allclose(gs[i], gsi, rtol=0, atol=1e-13 * np.max(np.abs(gsi)))
            assert np.allclose(gp[i], gpi, rtol=1e-12) and np.isclose(c[i], ci, rtol=1e-13)
            assert np.allclose(mv[i], ko.matvec(Ms[i], ps[i], vs[i], ps[i], cid), rtol=0, atol=1e-12 * np.max(np.abs(mv[i])))


def test_hunt_outcomes(gold_hunt):
    """End-state parity of the oracle's solver loops with reference hunt(): status, nit, cost (1e-10), parameters."""
    g = gold_hunt
    for cid in range(4):
        key = f"hunt/{cid}"
        M, p = g[key + "/modes"], g[key + "/params"]
        for pc, tag in ((False, "/adj300"), (True, "/adj300_pc")):
            Mo, po, st = ko.adjoint_descent(M, p, cid, maxiter=300, preconditioning=pc)
            status, nit = (int(x) for x in g[key + tag + "/status_nit"])
            assert (ko.hunt_status(st, st["cost"], 1e-6), st["nit"]) == (status, nit)
            assert abs(st["cost"] - g[key + tag + "/costs"][-1]) < 1e-10
            assert relerr(Mo, g[key + tag + "/state"]) < 1e-10 and np.allclose(po, g[key + tag + "/params"], rtol=1e-11)
        for method, kw in (("gmres", {"restart": 20, "maxiter": 2}), ("lsqr", {"iter_lim": 50})):
            Mo, po, st = ko.newton_krylov(M, p, cid, method=method, maxiter=3, scipy_kwargs=kw)
            status, nit = (int(x) for x in g[key + "/" + method + "/status_nit"])
            assert (ko.hunt_status(st, st["cost"], 1e-6), st["nit"]) == (status, nit)
            assert abs(st["cost"] - g[key + "/" + method + "/costs"][-1]) < 1e-9
    tab = g["hunt/seeds_adj1500"]
    for row in tab[:3]:
        seed = int(row[0])
        M = ko.populate_state(44.0, 33.4, 32, 32, seed, ko.ORBIT)
        Mo, po, st = ko.adjoint_descent(M, np.array([44.0, 33.4, 0.0]), ko.ORBIT, maxiter=1500)
        assert (ko.hunt_status(st, st["cost"], 1e-6), st["nit"]) == (int(row[1]), int(row[2]))
        assert abs(st["cost"] - row[3]) < 1e-9 and np.allclose(po[:2], row[4:6], rtol=1e-11)


def test_product_host_populate_matches_reference_generator(gold_eval):
    """The product's own host-side populate() (orbithunter_b200/ks.py, no GPU involved) against the reference's generator
    outputs in tests/golden and against the oracle at sizes where the reference's integer truncation of its frequency
    tables matters (m = 24, 34: ks/orbits.py:1788-1791)."""
    from orbithunter_b200.ks import CLASSES
    from orbithunter_b200.batch import CLASS_NAMES

    g = gold_eval
    n_checked = 0
    for key in sorted({k.rsplit("/", 1)[0] for k in g.files if k.startswith("random/")}):
        cid = int(key.split("/")[1].split("_")[0])
        if key + "/seed" not in g.files:
            continue
        M, p = g[key + "/modes"], g[key + "/params"]
        n, m = ko.field_shape(M, cid)
        o = CLASSES[CLASS_NAMES[cid]](parameters=(float(p[0]), float(p[1]), 0.0)).populate(attr="state", seed=int(g[key + "/seed"]),
                                                                                           discretization=(n, m))
        assert relerr(o.state, M) < 1e-14, key
        n_checked += 1
    for cid in range(4):
        for (n, m) in ((16, 24), (26, 34), (34, 26), (64, 64)):
            o = CLASSES[CLASS_NAMES[cid]](parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=11, discretization=(n, m))
            assert relerr(o.state, ko.populate_state(44.0, 33.4, n, m, 11, cid)) < 1e-14


def test_product_host_resize_matches_reference():
    """OrbitKS.resize / _pad / _truncate of the product's host classes in the modes basis (pure NumPy, no GPU) against the
    unmodified reference's outputs (tests/golden/ks_continuation.npz, oracle/make_golden_continuation.py)."""
    import os

    from orbithunter_b200.batch import CLASS_NAMES
    from orbithunter_b200.ks import CLASSES

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_continuation.npz"))
    for cid in range(4):
        o = CLASSES[CLASS_NAMES[cid]](state=g[f"resize/{cid}/modes"], basis="modes", parameters=(44.0, 44.0, 0.0))
        assert o.discretization == (16, 24)
        for shape in ((16, 24), (20, 24), (16, 30), (12, 20), (24, 16), (32, 32)):
            r = o.resize(*shape)
            want = g[f"resize/{cid}/{shape[0]}x{shape[1]}"]
            assert r.discretization == shape and r.basis == "modes" and r.state.shape == want.shape
            assert np.array_equal(r.state, want), (cid, shape)
        with pytest.raises(ValueError):
            o.resize(16, 25)
