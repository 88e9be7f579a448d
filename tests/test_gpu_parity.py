"""GPU (-m gpu): the nvcc-built sm_100a kernels, called through the C ABI via BatchedOrbitKS / the drop-in classes,
against (a) the reference's own outputs in tests/golden, (b) the NumPy oracle on seeded inputs, (c) size-independent
properties at the BASELINE sizes.  Tolerance: north_star's 1e-12 relative for residual / matvec / rmatvec."""
import numpy as np
import pytest

from conftest import golden_cases, relerr
from oracle import ks_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-12


def _b(cid, M, p, cons=None, basis="modes"):
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.batch import CLASS_NAMES

    c = None if cons is None else dict(zip("txs", (bool(x) for x in cons)))
    return BatchedOrbitKS(np.ascontiguousarray(M), np.ascontiguousarray(p), CLASS_NAMES[cid], basis, constraints=c)


def _case_cls(g, key):
    if key.startswith("fixed6/"):
        return int(key.split("/")[1])
    if key.startswith("stored/"):
        return int(g[key + "/cls"])
    return int(key.split("/")[1].split("_")[0])


def _tol(M, cid, L):
    n, m = ko.field_shape(M, cid)
    return max(TOL, 64 * 2.2e-16 * (2 * np.pi * (m // 2 - 1) / L) ** 4)


def test_native_library_is_loaded():
    from orbithunter_b200 import _lib

    assert _lib.lib().ks_b200_abi_version() == 1
    maps = open("/proc/self/maps").read()
    assert "libks_b200.so" in maps


def test_golden_all_cases(gold_eval):
    """Every golden case (6x6 fixture, stored orbits incl. 26x26 / 34x34, random seeds at 32^2, 64^2, 26x34, 16x128)."""
    g = gold_eval
    cases = golden_cases(g, "fixed6") + golden_cases(g, "stored") + golden_cases(g, "random")
    for key in cases:
        cid = _case_cls(g, key)
        M, p, v, vp = g[key + "/modes"], g[key + "/params"], g[key + "/v"], g[key + "/vparams"]
        cons = tuple(bool(x) for x in g[key + "/constraints"])
        tol = _tol(M, cid, p[1])
        u = _b(cid, M[None], p[None], cons)
        fscale = float(np.max(np.abs(ko.eqn_linear(M, p, cid)))) if key.startswith("stored/") else None
        assert relerr(u.transform("field").cpu_state()[0], g[key + "/field"]) < TOL, key
        assert relerr(u.transform("spatial_modes").cpu_state()[0], g[key + "/smodes"]) < TOL, key
        assert relerr(_b(cid, g[key + "/field"][None], p[None], cons, "field").transform("modes").cpu_state()[0], M) < TOL, key
        F, c = u.eqn(return_cost=True)
        assert relerr(F.cpu_state()[0], g[key + "/eqn"], fscale) < tol, key
        assert abs(float(c[0]) - float(g[key + "/cost"])) <= max(1e-10, 1e-12 * float(g[key + "/cost"])), key
        vb = _b(cid, v[None], vp[None], cons)
        assert relerr(u.matvec(vb).cpu_state()[0], g[key + "/matvec"]) < tol, key
        r = u.rmatvec(vb)
        assert relerr(r.cpu_state()[0], g[key + "/rmatvec"]) < tol, key
        assert np.allclose(r.cpu_parameters()[0], g[key + "/rmatvec_params"], rtol=1e-10, atol=1e-10), key
        for pc, tag in ((False, "/costgrad"), (True, "/costgrad_pc")):
            c2, gr = u.costgrad(preconditioning=pc)
            gscale = None if fscale is None else fscale * float(np.max(np.abs(g[key + "/rmatvec"]))) / float(np.max(np.abs(M)))
            assert relerr(gr.cpu_state()[0], g[key + tag], gscale) < tol, key
            assert np.allclose(gr.cpu_parameters()[0], g[key + tag + "_params"], rtol=1e-10,
                               atol=1e-10 if gscale is None else 1e-10 * gscale), key


def test_reference_known_answers_single_orbit_api(gold_eval):
    """The reference's asserted numbers (tests/test_basic.py:634-651) through the drop-in classes."""
    from orbithunter_b200 import OrbitKS, RelativeOrbitKS

    g = gold_eval
    o = OrbitKS(state=g["fixed6/0/field_in"], basis="field", parameters=(44, 44, 5)).transform(to="modes")
    assert o.rmatvec(o).norm() == pytest.approx(1.295386, rel=1e-6)
    assert o.costgrad(o.eqn()).norm() == pytest.approx(0.468832, rel=1e-6)
    assert o.costgrad().norm() == pytest.approx(0.468832, rel=1e-6)
    assert o.matvec(o).norm() == pytest.approx(0.3641827310989932, rel=1e-9)
    r = RelativeOrbitKS(state=g["stored/rpo/modes"], basis="modes", parameters=tuple(g["stored/rpo/params"]))
    assert r.rmatvec(r).norm() == pytest.approx(60.18805016, rel=1e-8)
    assert r.matvec(r).norm() == pytest.approx(155.86156902189782, rel=1e-9)
    assert r.cost() < 1e-7
    J = o.jacobian()
    assert J.shape == (o.size, o.size + 2)
    v = np.random.RandomState(0).randn(J.shape[1])
    vo = o.from_numpy_array(v)
    assert relerr(J @ v, o.matvec(vo).state.ravel()) < 1e-12


def test_large_domain_256(gold_big):
    g = gold_big
    n, m = (int(x) for x in g["big/shape"])
    sub = (slice(None, None, int(g["big/sub"][0])), slice(None, None, int(g["big/sub"][1])))
    M = ko.populate_state(200.0, 200.0, n, m, int(g["big/seed"]), ko.ORBIT)
    p, vp = g["big/params"], g["big/vparams"]
    v = np.random.RandomState(int(g["big/vseed"])).randn(*M.shape)
    u, vb = _b(0, M[None], p[None]), _b(0, v[None], vp[None])
    F, c = u.eqn(return_cost=True)
    assert relerr(F.cpu_state()[0][sub], g["big/eqn"]) < TOL
    assert float(c[0]) == pytest.approx(float(g["big/cost"]), rel=1e-12)
    assert relerr(u.matvec(vb).cpu_state()[0][sub], g["big/matvec"]) < TOL
    r = u.rmatvec(vb)
    assert relerr(r.cpu_state()[0][sub], g["big/rmatvec"]) < TOL
    assert np.allclose(r.cpu_parameters()[0], g["big/rmatvec_params"], rtol=1e-10)
    _, gr = u.costgrad(preconditioning=True)
    assert relerr(gr.cpu_state()[0][sub], g["big/costgrad_pc"]) < TOL


@pytest.mark.parametrize("cid", range(4))
@pytest.mark.parametrize("shape,B", [((32, 32), 301), ((64, 64), 33), ((26, 34), 5), ((128, 128), 3)])
def test_oracle_parity_batched(cid, shape, B):
    n, m = shape
    M = np.stack([ko.populate_state(60.0, 44.0, n, m, 1000 + s, cid) for s in range(B)])
    rng = np.random.RandomState(n + cid)
    p = np.stack([60.0 + 5 * rng.rand(B), 44.0 + 3 * rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    v, vp = rng.randn(*M.shape), rng.randn(B, 3)
    for cons in (ko.default_constraints(cid), (True, False, True), (True, True, True)):
        u, vb = _b(cid, M, p, cons), _b(cid, v, vp, cons)
        tol = _tol(M[0], cid, 44.0)
        F, c = u.eqn(return_cost=True)
        assert relerr(F.cpu_state(), ko.eqn(M, p, cid)) < tol
        assert np.allclose(c.cpu().numpy(), ko.cost(M, p, cid), rtol=1e-12)
        assert relerr(u.matvec(vb).cpu_state(), ko.matvec(M, p, v, vp, cid, cons)) < tol
        r = u.rmatvec(vb)
        ors, orp = ko.rmatvec(M, p, v, cid, cons)
        assert relerr(r.cpu_state(), ors) < tol
        assert np.allclose(r.cpu_parameters(), orp, rtol=1e-10, atol=1e-10 * np.max(np.abs(orp)) + 1e-300)
        for pc in (False, True):
            c2, gr, Fo = u.costgrad(preconditioning=pc, return_eqn=True)
            oc, ogs, ogp, oF = ko.costgrad(M, p, cid, cons, pc)
            assert relerr(gr.cpu_state(), ogs) < tol and relerr(Fo.cpu_state(), oF) < tol
            assert np.allclose(c2.cpu().numpy(), oc, rtol=1e-12)
            assert np.allclose(gr.cpu_parameters(), ogp, rtol=1e-10, atol=1e-10 * np.max(np.abs(ogp)) + 1e-300)


@pytest.mark.parametrize("cid", range(4))
def test_warp32_matches_generic_kernel(cid):
    """Two independent CUDA implementations (register-resident vs generic-size) agree at 32x32."""
    from orbithunter_b200 import _lib

    B = 64
    M = np.stack([ko.populate_state(44.0, 33.4, 32, 32, s, cid) for s in range(B)])
    p = np.tile([44.0, 33.4, 0.4 if cid == 1 else 0.0], (B, 1))
    u = _b(cid, M, p)
    c1, g1 = u.costgrad()
    _lib.lib().ks_debug_force_generic(1)
    try:
        c2, g2 = u.costgrad()
    finally:
        _lib.lib().ks_debug_force_generic(0)
    assert relerr(g1.cpu_state(), g2.cpu_state()) < TOL and np.allclose(c1.cpu().numpy(), c2.cpu().numpy(), rtol=1e-13)


@pytest.mark.parametrize("variant", [3, 4, 5, 6])
@pytest.mark.parametrize("cid", range(4))
def test_warp32_variants_vs_oracle(cid, variant):
    """Every build of the 32x32 eqn / costgrad kernel (two orbits per warp; one orbit per warp with 8 and with 12 warps per
    SM) against the oracle, on a batch that spans several rounds of the persistent warps and with mixed constraints."""
    from orbithunter_b200 import _lib

    B = 3 * 148 * 12 + 77
    M = np.stack([ko.populate_state(44.0, 33.4, 32, 32, s % 97, cid) for s in range(B)]) * (1.0 + 1e-3 * np.arange(B)[:, None, None])
    rng = np.random.RandomState(cid)
    p = np.stack([44.0 + rng.rand(B), 33.4 + rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    L = _lib.lib()
    L.ks_debug_warp32_variant(variant)
    try:
        for cons in (ko.default_constraints(cid), (True, False, False)):
            u = _b(cid, M, p, cons)
            F, c = u.eqn(return_cost=True)
            assert relerr(F.cpu_state(), ko.eqn(M, p, cid)) < TOL
            assert np.allclose(c.cpu().numpy(), ko.cost(M, p, cid), rtol=1e-12)
            for pc in (False, True):
                c2, gr, Fo = u.costgrad(preconditioning=pc, return_eqn=True)
                oc, ogs, ogp, oF = ko.costgrad(M, p, cid, cons, pc)
                assert relerr(gr.cpu_state(), ogs) < TOL and relerr(Fo.cpu_state(), oF) < TOL
                assert np.allclose(c2.cpu().numpy(), oc, rtol=1e-12)
                assert np.allclose(gr.cpu_parameters(), ogp, rtol=1e-10, atol=1e-10 * np.max(np.abs(ogp)) + 1e-300)
    finally:
        L.ks_debug_warp32_variant(4)


@pytest.mark.parametrize("cid", range(4))
@pytest.mark.parametrize("shape,B", [((64, 64), 97), ((128, 64), 9), ((64, 256), 5), ((256, 256), 3)])
def test_tile_matches_generic_kernel(cid, shape, B):
    """The multi-pass large-grid path (ks_tile.cuh) against the generic-size kernel, incl. a batch split into chunks."""
    from orbithunter_b200 import _lib

    n, m = shape
    rng = np.random.RandomState(n + m + cid)
    M = np.stack([ko.populate_state(60.0, 44.0, n, m, s, cid) for s in range(B)])
    p = np.stack([60.0 + rng.rand(B), 44.0 + rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    u, vb = _b(cid, M, p), _b(cid, rng.randn(*M.shape), rng.randn(B, 3))
    L = _lib.lib()

    def run():
        c, g, F = u.costgrad(preconditioning=True, return_eqn=True)
        r = u.rmatvec(vb)
        return [c.cpu().numpy(), g.cpu_state(), g.cpu_parameters(), F.cpu_state(), u.matvec(vb).cpu_state(), r.cpu_state(),
                r.cpu_parameters(), u.eqn().cpu_state()]

    try:
        L.ks_debug_tile(2, 2)  # multi-pass tiles, 2 MB workspace budget: 1 .. 20 orbits per chunk
        small = run()
        L.ks_debug_tile(2, 256)
        big = run()
        L.ks_debug_tile(1, 256)  # default: at 64x64 the CTA-resident fused kernel (ks_tile64.cuh), else the same tiles
        dflt = run()
        L.ks_debug_tile(0, 0)
        ref = run()
    finally:
        L.ks_debug_tile(1, 256)
    for x, y, d, z in zip(small, big, dflt, ref):
        assert np.array_equal(x, y)  # chunking does not change a single bit
        assert relerr(x, z) < _tol(M[0], cid, 44.0) and relerr(d, z) < _tol(M[0], cid, 44.0)


@pytest.mark.parametrize("cid", range(4))
def test_properties_at_full_batch(cid):
    """BASELINE config 2/3 sizes: adjoint identity <Jv,w> = <v,J^T w> + params, transform round trip, determinism."""
    n, m, B = (32, 32, 4096) if cid == 0 else (64, 64, 512)
    rng = np.random.RandomState(cid)
    M0 = ko.populate_state(60.0, 44.0, n, m, 3, cid)
    M = M0[None] * (1.0 + 0.1 * rng.rand(B, 1, 1))
    p = np.stack([60.0 + rng.rand(B), 44.0 + rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    u = _b(cid, M, p)
    v, w = _b(cid, rng.randn(*M.shape), rng.randn(B, 3)), _b(cid, rng.randn(*M.shape), np.zeros((B, 3)))
    cons = np.array(ko.default_constraints(cid))
    v.parameters[:, torch.as_tensor(cons)] = 0.0
    lhs = u.matvec(v).dot(w)
    r = u.rmatvec(w)
    rhs = r.dot(v) + (r.parameters * v.parameters).sum(dim=1)
    assert torch.all((lhs - rhs).abs() <= 1e-11 * lhs.abs().clamp_min(1.0))
    back = u.transform("field").transform("modes")
    assert relerr(back.cpu_state(), M) < 1e-13
    c1, g1 = u.costgrad()
    c2, g2 = u.costgrad()
    assert torch.equal(g1.state, g2.state) and torch.equal(c1, c2)


def test_edge_cases():
    from orbithunter_b200 import BatchedOrbitKS, KsB200Error, OrbitKS

    e = BatchedOrbitKS(torch.zeros(0, 31, 30, dtype=torch.float64, device="cuda"),
                       torch.zeros(0, 3, dtype=torch.float64, device="cuda"))
    c, g = e.costgrad()
    assert c.numel() == 0 and g.state.numel() == 0
    one = BatchedOrbitKS(np.zeros((1, 31, 30)), np.array([[44.0, 33.4, 0.0]]))
    c, g = one.costgrad()
    assert float(c[0]) == 0.0 and float(g.state.abs().max()) == 0.0
    with pytest.raises(ValueError):
        OrbitKS(state=np.zeros((31, 30)), basis="modes", parameters=(44.0, 33.4, 0.0)).transform(to="nonsense")
    with pytest.raises(AssertionError):
        OrbitKS(state=np.zeros((32, 32)), basis="field", parameters=(44.0, 33.4, 0.0)).eqn()
    with pytest.raises(KsB200Error):
        BatchedOrbitKS(np.zeros((1, 2, 1)), np.array([[44.0, 33.4, 0.0]]), basis="modes").eqn()  # n = 3: unsupported
