"""GPU (-m gpu): hunt() end-state parity with the reference (tests/golden/ks_hunt.npz, produced by the unmodified
reference) and with the oracle: identical status and nit, cost to 1e-10, parameters to 1e-9."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ks_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _cls(cid):
    from orbithunter_b200.ks import CLASSES
    from orbithunter_b200.batch import CLASS_NAMES

    return CLASSES[CLASS_NAMES[cid]]


@pytest.mark.parametrize("cid", range(4))
def test_adj_golden_single_orbit_api(gold_hunt, cid):
    from orbithunter_b200.optimize import hunt

    g = gold_hunt
    key = f"hunt/{cid}"
    o = _cls(cid)(state=g[key + "/modes"], basis="modes", parameters=tuple(g[key + "/params"]))
    for pc, tag in ((False, "/adj300"), (True, "/adj300_pc")):
        r = hunt(o, "adj", maxiter=300, preconditioning=pc)
        status, nit = (int(x) for x in g[key + tag + "/status_nit"])
        assert (r.status, r.nit) == (status, nit)
        assert abs(r.costs[-1] - g[key + tag + "/costs"][-1]) < 1e-10
        assert abs(r.costs[0] - g[key + tag + "/costs"][0]) < 1e-10 * g[key + tag + "/costs"][0]
        assert relerr(r.orbit.state, g[key + tag + "/state"]) < 1e-9
        assert np.allclose(r.orbit.parameters, g[key + tag + "/params"], rtol=1e-10)
        assert r.method == "adj" and r.maxiter == 300 and isinstance(r.message, str)


def test_adj_seeds_batched(gold_hunt):
    """BASELINE config 2 settings on reference seeds 0..7 (1500 steps): per-seed status / nit / cost / (T, L)."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    tab = gold_hunt["hunt/seeds_adj1500"]
    seeds = tab[:, 0].astype(int)
    M = np.stack([ko.populate_state(44.0, 33.4, 32, 32, int(s), ko.ORBIT) for s in seeds])
    p = np.tile([44.0, 33.4, 0.0], (len(seeds), 1))
    r = hunt(BatchedOrbitKS(M, p, "OrbitKS"), "adj", maxiter=1500)
    assert np.array_equal(r.status.cpu().numpy(), tab[:, 1].astype(int))
    assert np.array_equal(r.nit.cpu().numpy(), tab[:, 2].astype(int))
    assert np.allclose(r.costs[-1].cpu().numpy(), tab[:, 3], rtol=0, atol=1e-8)
    assert np.allclose(r.orbit.cpu_parameters()[:, :2], tab[:, 4:6], rtol=1e-9)


def test_adj_converges_status_1(gold_hunt):
    """The stored rpo perturbed by 1e-3 randn converges under 'adj' with tol 1e-5 (reference: status 1)."""
    from orbithunter_b200 import RelativeOrbitKS
    from orbithunter_b200.optimize import hunt

    g = gold_hunt
    o = RelativeOrbitKS(state=g["hunt/rpo_adj/modes"], basis="modes", parameters=tuple(g["hunt/rpo_adj/params"]))
    r = hunt(o, "adj", maxiter=20000, tol=1e-5)
    status, nit = (int(x) for x in g["hunt/rpo_adj/status_nit"])
    assert r.status == status == 1
    assert abs(r.nit - nit) <= max(2, nit // 100)  # thousands of data-dependent steps: allow a 1 % drift in the count
    assert abs(r.costs[-1] - g["hunt/rpo_adj/costs"][-1]) < 1e-7 and r.costs[-1] <= 1e-5
    assert np.allclose(r.orbit.parameters, g["hunt/rpo_adj/out_params"], rtol=1e-6)


def test_mixed_batch_exit_codes_and_lists():
    """Per-orbit exit codes inside one batch; per-method list kwargs as in optimize.py:414-424."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    M = np.stack([ko.populate_state(44.0, 44.0, 32, 32, 3, 0), 1e-9 * ko.populate_state(44.0, 44.0, 32, 32, 4, 0),
                  ko.populate_state(44.0, 44.0, 32, 32, 5, 0)])
    p = np.tile([44.0, 44.0, 0.0], (3, 1))
    r = hunt(BatchedOrbitKS(M, p, "OrbitKS"), ("adj", "adj"), maxiter=[20, 10], tol=[1e-6, 1e-7], preconditioning=[False, True])
    st, nit = r.status.cpu().numpy(), [x.cpu().numpy() for x in r.nit]
    assert st[1] == 1 and nit[0][1] == 0 and st[0] == 2 and list(nit[0][[0, 2]]) == [20, 20] and list(nit[1][[0, 2]]) == [10, 10]
    assert r.method == ["adj", "adj"] and len(r.costs) == 4
    for b in (0, 2):
        m1, p1, s1 = ko.adjoint_descent(M[b], p[b], 0, maxiter=20)
        m2, p2, s2 = ko.adjoint_descent(m1, p1, 0, maxiter=10, tol=1e-7, preconditioning=True)
        assert abs(float(r.costs[-1][b]) - s2["cost"]) < 1e-9 * max(1.0, s2["cost"])
    with pytest.raises(NotImplementedError):
        hunt(BatchedOrbitKS(M, p, "OrbitKS"), "nelder-mead")
    with pytest.raises(TypeError):
        hunt(BatchedOrbitKS(M, p, "OrbitKS"), "adj", tol="tight")


@pytest.mark.parametrize("cid", range(4))
def test_newton_krylov_golden(gold_hunt, cid):
    """hunt('gmres') / hunt('lsqr') end state vs the unmodified reference (which runs SciPy's solvers): identical
    status and nit; cost to 1e-9 for GMRES.  Krylov iteration histories themselves are unpinned by the reference."""
    from orbithunter_b200.optimize import hunt

    g = gold_hunt
    key = f"hunt/{cid}"
    o = _cls(cid)(state=g[key + "/modes"], basis="modes", parameters=tuple(g[key + "/params"]))
    r = hunt(o, "gmres", maxiter=3, scipy_kwargs={"restart": 20, "maxiter": 2})
    status, nit = (int(x) for x in g[key + "/gmres/status_nit"])
    assert (r.status, r.nit) == (status, nit)
    assert abs(r.costs[-1] - g[key + "/gmres/costs"][-1]) < 1e-9 * max(1.0, g[key + "/gmres/costs"][-1])
    assert relerr(r.orbit.state, g[key + "/gmres/state"]) < 1e-8
    r = hunt(o, "lsqr", maxiter=3, scipy_kwargs={"iter_lim": 50})
    status, nit = (int(x) for x in g[key + "/lsqr/status_nit"])
    assert (r.status, r.nit) == (status, nit)
    assert abs(r.costs[-1] - g[key + "/lsqr/costs"][-1]) < 1e-3 * max(1.0, g[key + "/lsqr/costs"][-1])


def test_gmres_converges_status_1(gold_hunt):
    """Perturbed rpo: hunt('gmres', restart 50, 4 cycles, rtol 1e-8, tol 1e-8) converges in the reference (status 1, 7 its)."""
    from orbithunter_b200 import RelativeOrbitKS
    from orbithunter_b200.optimize import hunt

    g = gold_hunt
    o = RelativeOrbitKS(state=g["hunt/rpo_adj/modes"], basis="modes", parameters=tuple(g["hunt/rpo_adj/params"]))
    r = hunt(o, "gmres", maxiter=20, tol=1e-8, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})
    status, nit = (int(x) for x in g["hunt/rpo_gmres/status_nit"])
    assert r.status == status == 1 and r.nit == nit
    assert abs(r.costs[-1] - g["hunt/rpo_gmres/costs"][-1]) < 1e-10
    assert np.allclose(r.orbit.parameters, g["hunt/rpo_gmres/out_params"], rtol=1e-8)


@pytest.mark.parametrize("shape,B,mode", [((32, 32), 9, 4), ((64, 64), 5, 2), ((128, 128), 3, 1), ((256, 256), 2, 1)])
def test_gmres_cluster_step_matches_single_cta_step(shape, B, mode):
    """Two implementations of the GMRES step (thread-block cluster with register-resident slices vs one CTA per orbit):
    same outer status / nit, cost and state to round-off (only the association of the dot-product sums differs)."""
    from orbithunter_b200 import BatchedOrbitKS, _lib
    from orbithunter_b200.optimize import hunt
    from oracle import ks_oracle as ko

    n, m = shape
    T, L = (200.0, 200.0) if n == 256 else (60.0, 44.0)
    M = np.stack([ko.populate_state(T, L, n, m, s, 0) for s in range(B)])
    u = BatchedOrbitKS(M, np.tile([T, L, 0.0], (B, 1)), "OrbitKS")
    kw = dict(maxiter=2, tol=1e-10, scipy_kwargs={"restart": 12, "maxiter": 2, "rtol": 1e-6})
    L_ = _lib.lib()
    try:
        L_.ks_debug_gmres_cluster(mode)
        a = hunt(u, "gmres", **kw)
        L_.ks_debug_gmres_cluster(0)
        b = hunt(u, "gmres", **kw)
    finally:
        L_.ks_debug_gmres_cluster(1)
    assert torch.equal(a.status, b.status) and torch.equal(a.nit, b.nit)
    assert torch.allclose(a.costs[-1], b.costs[-1], rtol=1e-9)
    assert relerr(a.orbit.cpu_state(), b.orbit.cpu_state()) < 1e-9


def test_hybrid_adj_then_gmres_batched_64():
    """BASELINE config 3 shape in miniature: 64x64 symmetry classes, adjoint warm-up then Newton-GMRES, vs the oracle."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.batch import CLASS_NAMES
    from orbithunter_b200.optimize import hunt

    for cid in (1, 2, 3):
        M = np.stack([ko.populate_state(60.0, 44.0, 64, 64, 10 + s, cid) for s in range(3)])
        p = np.tile([60.0, 44.0, 0.0], (3, 1))
        r = hunt(BatchedOrbitKS(M, p, CLASS_NAMES[cid]), ("adj", "gmres"), maxiter=[60, 2], tol=[1e-6, 1e-10],
                 scipy_kwargs=[{}, {"restart": 10, "maxiter": 1, "rtol": 1e-6}])
        for b in range(3):
            m1, p1, s1 = ko.adjoint_descent(M[b], p[b], cid, maxiter=60)
            m2, p2, s2 = ko.newton_krylov(m1, p1, cid, method="gmres", maxiter=2, tol=1e-10,
                                          scipy_kwargs={"restart": 10, "maxiter": 1, "rtol": 1e-6})
            assert int(r.nit[0][b]) == s1["nit"] and int(r.nit[1][b]) == s2["nit"]
            assert int(r.status[b]) == ko.hunt_status(s2, s2["cost"], 1e-10)
            assert abs(float(r.costs[-1][b]) - s2["cost"]) < 1e-8 * max(1.0, s2["cost"])


def test_newton_descent_matrix_free():
    """'newton_descent' restated matrix-free: the step equals pinv(J)(-F) (dense reference path, optimize.py:785-787)."""
    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.optimize import hunt

    M = ko.populate_state(44.0, 44.0, 16, 16, 3, 0)
    o = OrbitKS(state=M, basis="modes", parameters=(44.0, 44.0, 0.0))
    J = o.jacobian()
    F = o.eqn().state.ravel()
    dx = np.linalg.pinv(J) @ (-F)
    expect = o.increment(o.from_numpy_array(dx))
    c0, c1 = o.cost(), expect.cost()
    r = hunt(o, "newton_descent", maxiter=1)
    assert r.nit == 1
    if c1 <= c0:  # full step accepted by the reference logic
        assert relerr(r.orbit.state, expect.state) < 1e-7 and abs(r.costs[-1] - c1) < 1e-7 * max(1.0, c1)


def test_continuation_matches_reference_family():
    """continuation() (reference continuation.py:59-136) on the device hunt, against the family members the
    unmodified reference reaches from its stored rpo (tests/golden/ks_continuation.npz, oracle/make_golden_continuation.py):
    single-orbit drop-in and the lock-step batch form."""
    import os

    from orbithunter_b200 import BatchedOrbitKS, RelativeOrbitKS
    from orbithunter_b200.continuation import continuation, continuation_batch

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_continuation.npz"))
    kw = dict(methods="gmres", maxiter=40, tol=1e-6, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})
    for tag in ("up", "down"):
        o = RelativeOrbitKS(state=g["start/modes"], basis="modes", parameters=tuple(g["start/params"]))
        r = continuation(o, ("t", float(g[tag + "/target"])), step_size=float(g[tag + "/step"]),
                         **{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
        assert r.status == int(g[tag + "/status"]) == 1
        assert np.allclose(r.orbit.parameters, g[tag + "/params"], rtol=1e-7)
        assert relerr(r.orbit.state, g[tag + "/modes"]) < 1e-6
        assert r.orbit.cost() < 1e-6
    B = 3
    T0 = float(g["start/params"][0])
    batch = BatchedOrbitKS(np.repeat(g["start/modes"][None], B, 0), np.repeat(g["start/params"][None], B, 0), "RelativeOrbitKS")
    targets = np.array([float(g["up/target"]), float(g["down/target"]), T0])
    rb = continuation_batch(batch, "t", targets, step_size=0.02, **kw)
    assert rb.status.tolist() == [1, 1, 1] and rb.reached.tolist() == [True, True, True] and rb.steps.tolist() == [3, 2, 0]
    P = rb.orbit.cpu_parameters()
    assert np.allclose(P[0], g["up/params"], rtol=1e-7) and np.allclose(P[1], g["down/params"], rtol=1e-7)
    assert np.array_equal(P[2], g["start/params"])
    S = rb.orbit.cpu_state()
    assert relerr(S[0], g["up/modes"]) < 1e-6 and relerr(S[1], g["down/modes"]) < 1e-6 and np.array_equal(S[2], g["start/modes"])
    # a family whose first re-convergence fails keeps its last converged member and reports the failing hunt's status
    rf = continuation_batch(batch, "t", targets, step_size=0.02, methods="gmres", maxiter=1, tol=1e-14,
                            scipy_kwargs={"restart": 5, "maxiter": 1, "rtol": 1e-2})
    assert rf.status.tolist() == [2, 2, 1] and rf.reached.tolist() == [False, False, True] and rf.steps.tolist() == [0, 0, 0]
    assert np.array_equal(rf.orbit.cpu_state(), batch.cpu_state()) and np.array_equal(rf.orbit.cpu_parameters(), batch.cpu_parameters())


def test_discretization_continuation_matches_reference():
    """discretization_continuation (reference continuation.py:170-279) with resize on the host classes and hunt on the
    device: the stored rpo taken from 32x32 to 32x36, as the unmodified reference does it."""
    import os

    from orbithunter_b200 import RelativeOrbitKS
    from orbithunter_b200.continuation import discretization_continuation

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_continuation.npz"))
    o = RelativeOrbitKS(state=g["start/modes"], basis="modes", parameters=tuple(g["start/params"]))
    r = discretization_continuation(o, (32, 36), methods="gmres", maxiter=40, tol=1e-6,
                                    scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})
    assert r.status == int(g["disc/status"]) == 1 and tuple(r.orbit.discretization) == tuple(g["disc/shape"]) == (32, 36)
    assert np.allclose(r.orbit.parameters, g["disc/params"], rtol=1e-7) and relerr(r.orbit.state, g["disc/modes"]) < 1e-6
    assert r.orbit.cost() < 1e-6


def test_device_seed_generator(gold_eval):
    """BatchedOrbitKS.from_seeds (ks_populate_state, MT19937 + legacy randn on the device) against the
    reference generator's outputs in tests/golden and the oracle, incl. sizes where the reference's integer truncation of
    its frequency tables matters, and a large batch against the host generator."""
    from orbithunter_b200 import BatchedOrbitKS, OrbitKS
    from orbithunter_b200.batch import CLASS_NAMES

    g = gold_eval
    for key in sorted({k.rsplit("/", 1)[0] for k in g.files if k.startswith("random/") and k.endswith("/seed")}):
        cid = int(key.split("/")[1].split("_")[0])
        M, p = g[key + "/modes"], g[key + "/params"]
        n, m = ko.field_shape(M, cid)
        b = BatchedOrbitKS.from_seeds([int(g[key + "/seed"])], (float(p[0]), float(p[1]), 0.0), CLASS_NAMES[cid], (n, m))
        assert relerr(b.cpu_state()[0], M) < 1e-13, key
    for cid in range(4):
        for (n, m) in ((16, 24), (26, 34), (64, 64), (256, 256)):
            seeds = [0, 5, 2**32 - 1]
            b = BatchedOrbitKS.from_seeds(seeds, (60.0, 44.0, 0.0), CLASS_NAMES[cid], (n, m))
            for i, s in enumerate(seeds):
                assert relerr(b.cpu_state()[i], ko.populate_state(60.0, 44.0, n, m, s, cid)) < 1e-13
    B = 1000
    b = BatchedOrbitKS.from_seeds(np.arange(B), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    host = np.stack([OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=s, discretization=(32, 32)).state
                     for s in range(B)])
    assert relerr(b.cpu_state(), host) < 1e-13
    with pytest.raises(ValueError):
        BatchedOrbitKS.from_seeds([-1], (44.0, 33.4, 0.0))


def test_gluing_matches_reference():
    """glue / tile / aspect_ratio_correction (reference gluing.py:16-302) against the unmodified reference's
    outputs (tests/golden/ks_gluing.npz, oracle/make_golden_gluing.py): plain and strip-wise gluing in time, space and on a
    2x2 grid, tiles of different sizes, and a symbol-array tiling."""
    import os

    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.gluing import glue, tile

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_gluing.npz"))
    o = {t: OrbitKS(state=g[f"in/{t}/field"], basis="field", parameters=tuple(g[f"in/{t}/params"])) for t in "abc"}

    def arr(rows):
        x = np.empty((len(rows), len(rows[0])), dtype=object)
        for i, r in enumerate(rows):
            for j, v in enumerate(r):
                x[i, j] = v
        return x

    a, b, c = o["a"], o["b"], o["c"]
    cases = {"time": (arr([[a], [b]]), {}), "space": (arr([[a, b]]), {}), "time_strip": (arr([[a], [b]]), {"strip_wise": True}),
             "space_strip": (arr([[a, b]]), {"strip_wise": True}), "grid_strip": (arr([[a, b], [b, a]]), {"strip_wise": True}),
             "mixed_strip": (arr([[a], [c]]), {"strip_wise": True})}
    for tag, (oa, kw) in cases.items():
        r = glue(oa, OrbitKS, **kw)
        assert tuple(r.discretization) == tuple(g[tag + "/shape"]), tag
        assert np.allclose(r.parameters, g[tag + "/params"], rtol=1e-14), tag
        assert relerr(r.transform(to="field").state, g[tag + "/field"]) < 1e-13, tag
    t = tile(np.array([[0, 1], [1, 1]]), {0: a, 1: b}, OrbitKS, strip_wise=True)
    assert np.allclose(t.parameters, g["tile/params"], rtol=1e-14) and relerr(t.transform(to="field").state, g["tile/field"]) < 1e-13
    # the glued orbit is a usable initial condition on the device path
    assert t.transform(to="modes").cost() > 0.0


def test_lstsq_matches_reference():
    """hunt('lstsq') (reference optimize.py:855-1022, the second stage of BASELINE configs[0]): dense Jacobian from the
    device matvec, minimum-norm step from the SVD pseudo-inverse on the device, reference outer loop.  Goldens from the
    unmodified reference (tests/golden/ks_lstsq.npz): a 16x16 seed and configs[0] in miniature (adj then lstsq at 32x32)."""
    import os

    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.optimize import hunt

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_lstsq.npz"))
    o = OrbitKS(state=g["lstsq/modes"], basis="modes", parameters=tuple(g["lstsq/params"]))
    r = hunt(o, "lstsq", maxiter=3)
    assert (r.status, r.nit) == tuple(int(x) for x in g["lstsq/status_nit"])
    assert abs(r.costs[-1] - g["lstsq/costs"][-1]) < 1e-6 * g["lstsq/costs"][-1]
    assert relerr(r.orbit.state, g["lstsq/state"]) < 1e-6 and np.allclose(r.orbit.parameters, g["lstsq/out_params"], rtol=1e-7)
    # configs[0] in miniature: adjoint descent, then lstsq
    o2 = OrbitKS(state=g["cfg0/modes"], basis="modes", parameters=tuple(g["cfg0/params"]))
    r2 = hunt(o2, ("adj", "lstsq"), maxiter=[300, 4], tol=[1e-6, 1e-10])
    assert r2.status == int(g["cfg0/status"]) and list(r2.nit) == [int(x) for x in g["cfg0/nit"]]
    assert np.allclose(r2.costs, g["cfg0/costs"], rtol=1e-6)
    assert relerr(r2.orbit.state, g["cfg0/state"]) < 1e-6
    with pytest.raises(NotImplementedError):
        hunt(o, "lstsq", maxiter=1, preconditioning=True)
    with pytest.raises(NotImplementedError):
        hunt(o, "solve", maxiter=1)
