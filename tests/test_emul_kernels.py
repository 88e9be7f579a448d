"""CPU: the real kernel sources, compiled by g++ against the CPU-thread shim (tests/emul), match the oracle.

This is a logic check of the CUDA code in a container without a GPU (indexing, scalings, projections, reductions,
synchronisation structure); the parity tests proper are the -m gpu ones, which call the nvcc-built library.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_lib as E  # noqa: E402
from conftest import relerr  # noqa: E402
from oracle import ks_oracle as ko  # noqa: E402


def _inputs(cid, n, m, B, seed0=0):
    M = np.stack([ko.populate_state(44.0, 33.4, n, m, seed0 + s, cid) for s in range(B)])
    p = np.stack([[44.0 + s, 33.4 - 0.5 * s, 0.7 * s + 0.3 if cid == 1 else 0.0] for s in range(B)])
    rng = np.random.RandomState(7 + n + m)
    return M, p, rng.randn(*M.shape), rng.randn(B, 3)


def _tol(n, m, L=30.0):
    qmax4 = (2 * np.pi * (m // 2 - 1) / L) ** 4  # adjoint amplifies FFT round-off by q^4 (ill-resolved grids only)
    return max(1e-12, 64 * 2.2e-16 * qmax4)


def _check_all_ops(cid, n, m, B, cons):
    M, p, V, vp = _inputs(cid, n, m, B)
    cmask = sum(1 << i for i in range(3) if cons[i])
    tol = _tol(n, m)
    F, c = E.eqn(cid, n, m, M, p)
    assert relerr(F, ko.eqn(M, p, cid)) < tol
    assert relerr(c, ko.cost(M, p, cid)) < 1e-13
    for pc in (False, True):
        c2, g, gp, F2 = E.costgrad(cid, n, m, M, p, cmask, pc, want_f=pc)
        oc, ogs, ogp, oF = ko.costgrad(M, p, cid, cons, pc)
        assert relerr(g, ogs) < tol and relerr(c2, oc) < 1e-13
        assert np.allclose(gp, ogp, rtol=1e-11, atol=1e-11 * np.max(np.abs(ogp)))
        if pc:
            assert relerr(F2, oF) < tol
    assert relerr(E.matvec(cid, n, m, M, p, V, vp, cmask), ko.matvec(M, p, V, vp, cid, cons)) < tol
    rs, rp = E.rmatvec(cid, n, m, M, p, V, cmask)
    ors, orp = ko.rmatvec(M, p, V, cid, cons)
    assert relerr(rs, ors) < tol
    assert np.allclose(rp, orp, rtol=1e-11, atol=1e-11 * max(np.max(np.abs(orp)), 1e-300))


@pytest.mark.parametrize("variant", [3, 4, 5, 6, 7])
@pytest.mark.parametrize("cid", range(4))
def test_warp32_kernel_all_ops(cid, variant):
    E.lib().ks_debug_force_generic(0)
    E.lib().ks_debug_warp32_variant(variant)
    try:
        for cons in (ko.default_constraints(cid), (True, False, True), (False, True, False)):
            _check_all_ops(cid, 32, 32, 5, cons)  # odd batch: exercises the half-empty warp and the persistent loop
        _check_all_ops(cid, 32, 32, 37, ko.default_constraints(cid))  # more pairs than resident warps
    finally:
        E.lib().ks_debug_warp32_variant(4)


@pytest.mark.parametrize("cid", range(4))
@pytest.mark.parametrize("shape", [(32, 32), (6, 6), (26, 34), (16, 64)])
def test_generic_kernel_all_ops(cid, shape):
    E.lib().ks_debug_force_generic(1)
    try:
        for cons in (ko.default_constraints(cid), (True, True, True)):
            _check_all_ops(cid, shape[0], shape[1], 3, cons)
    finally:
        E.lib().ks_debug_force_generic(0)


@pytest.mark.parametrize("cid,shape", [(0, (64, 64)), (1, (64, 64)), (2, (64, 64)), (3, (64, 64)), (1, (128, 64)), (2, (64, 128)),
                                       (0, (64, 256)), (3, (256, 64))])
def test_tile_kernels_all_ops(cid, shape):
    """Large-grid multi-pass path (ks_tile.cuh): every pass kernel, every radix plan (64 = 8x8, 128 = 16x8, 256 = 16x16)."""
    L = E.lib()
    L.ks_debug_force_generic(0)
    L.ks_debug_tile(2, 1 if shape == (64, 64) else 64)  # multi-pass tiles; 1 MB budget: the 3-orbit batch runs as several chunks
    try:
        cons_list = (ko.default_constraints(cid), (True, True, True)) if shape == (64, 64) else (ko.default_constraints(cid),)
        for cons in cons_list:
            _check_all_ops(cid, shape[0], shape[1], 3, cons)
    finally:
        L.ks_debug_tile(1, 64)


@pytest.mark.parametrize("cid", range(4))
def test_fused64_kernel_all_ops(cid):
    """CTA-resident 64x64 kernel (ks_tile64.cuh): all ops, more orbits than resident CTAs (persistent loop)."""
    L = E.lib()
    L.ks_debug_force_generic(0)
    L.ks_debug_tile(1, 64)
    for cons in (ko.default_constraints(cid), (True, True, True)):
        _check_all_ops(cid, 64, 64, 3 if cons[1] else 11, cons)


@pytest.mark.parametrize("cid", range(4))
def test_transforms_all_bases(cid):
    for (n, m) in ((32, 32), (10, 12)):
        M, _, _, _ = _inputs(cid, n, m, 3)
        f, sm = ko.to_field(M, cid), ko.inv_time_transform(M, cid)
        assert relerr(E.transform(cid, n, m, M, 2, 0), f) < 1e-13
        assert relerr(E.transform(cid, n, m, M, 2, 1), sm) < 1e-13
        assert relerr(E.transform(cid, n, m, sm, 1, 0), f) < 1e-13
        rf = np.random.RandomState(5).randn(3, n, m)
        s2 = ko.space_transform(rf)
        assert relerr(E.transform(cid, n, m, rf, 0, 1), s2) < 1e-13
        assert relerr(E.transform(cid, n, m, rf, 0, 2), ko.to_modes(rf, cid)) < 1e-13
        assert relerr(E.transform(cid, n, m, s2, 1, 2), ko.time_transform(s2, cid)) < 1e-13


def test_golden_vectors_through_emulated_kernels(gold_eval):
    """The reference's own outputs (tests/golden) against the kernel sources, 32x32 cases of every class."""
    g = gold_eval
    E.lib().ks_debug_force_generic(0)
    for key in sorted({k.rsplit("/", 1)[0] for k in g.files if k.startswith("random/") and "_32x32_" in k}):
        cid = int(key.split("/")[1].split("_")[0])
        M, p, v, vp = (g[key + s][None] for s in ("/modes", "/params", "/v", "/vparams"))
        cons = tuple(bool(x) for x in g[key + "/constraints"])
        cmask = sum(1 << i for i in range(3) if cons[i])
        F, c = E.eqn(cid, 32, 32, M, p)
        assert relerr(F[0], g[key + "/eqn"]) < 1e-12 and abs(c[0] - g[key + "/cost"]) < 1e-12 * g[key + "/cost"]
        assert relerr(E.matvec(cid, 32, 32, M, p, v, vp, cmask)[0], g[key + "/matvec"]) < 1e-12
        rs, rp = E.rmatvec(cid, 32, 32, M, p, v, cmask)
        assert relerr(rs[0], g[key + "/rmatvec"]) < 1e-12
        assert np.allclose(rp[0], g[key + "/rmatvec_params"], rtol=1e-11, atol=1e-11)
        for pc, tag in ((False, "/costgrad"), (True, "/costgrad_pc")):
            _, gs, gp, _ = E.costgrad(cid, 32, 32, M, p, cmask, pc)
            assert relerr(gs[0], g[key + tag]) < 1e-12
            assert np.allclose(gp[0], g[key + tag + "_params"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("cid", range(4))
def test_derivatives_and_precondition(cid):
    n, m, B = 12, 16, 3
    M, p, V, vp = _inputs(cid, n, m, B)
    for order in (1, 2, 3, 4):
        assert relerr(E.deriv(cid, n, m, M, p, 0, order), ko.dt(M, p[:, 0], order)) < 1e-13
        sm = ko.inv_time_transform(M, cid)
        assert relerr(E.deriv(cid, n, m, sm, p, 1, order), ko.dx_spatial(sm, p[:, 1], order)) < 1e-13
        if cid in (0, 1) or order % 2 == 0:
            assert relerr(E.deriv(cid, n, m, M, p, 1, order), ko.dx(M, p[:, 1], cid, order)) < 1e-13
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    out, gp = E.precondition(cid, n, m, V, vp, p, cmask)
    eo, egp = ko.precondition(V, vp, p, n, m, cid, cons)
    assert relerr(out, eo) < 1e-13 and np.allclose(gp, egp, rtol=1e-13)


@pytest.mark.parametrize("cid,shape", [(0, (32, 32)), (1, (16, 24)), (2, (64, 64)), (3, (26, 34))])
def test_device_seed_generator_matches_reference_stream(cid, shape):
    """ks_populate_state: NumPy's legacy seed -> MT19937 -> randn stream and the reference's modulation, per seed."""
    n, m = shape
    seeds = np.array([0, 1, 7, 12345, 2**32 - 1], dtype=np.uint64)
    params = np.stack([[44.0 + 3 * i, 33.4 + 5 * i, 0.0] for i in range(len(seeds))])
    out = E.populate(cid, n, m, seeds, params)
    for i, s in enumerate(seeds):
        ref = ko.populate_state(params[i, 0], params[i, 1], n, m, int(s), cid)
        assert relerr(out[i], ref) < 1e-14


@pytest.mark.parametrize("cid,shape", [(1, (32, 32)), (2, (32, 32)), (1, (8, 10)), (2, (6, 8))])
def test_jacobian_columns_match_oracle_matvec(cid, shape):
    """ks_jacobian (OrbitKS.jacobian, ks/orbits.py:557-657): column j = J(u) e_j.  At 32x32 the one-orbit-per-warp kernel builds
    the unit vectors itself (Jacobian mode) and reads orbit e / ncol for evaluation e; other sizes go through a stored identity.
    Checked against the oracle's matvec on unit vectors -- a sample of columns at 32x32, all of them on the small grids."""
    n, m = shape
    B = 2
    M, p, _, _ = _inputs(cid, n, m, B)
    for cons in (ko.default_constraints(cid), (True, False, True)):
        cmask = sum(1 << i for i in range(3) if cons[i]) | (0 if cid == 1 else 4)
        Jt = E.jacobian(cid, n, m, M, p, cmask)
        nm = M.shape[1] * M.shape[2]
        free = [i for i in range(3) if not (cmask >> i) & 1]
        assert Jt.shape == (B, nm + len(free), nm)
        cols = range(nm + len(free)) if nm < 200 else list(range(0, nm, 41)) + list(range(nm - 3, nm + len(free)))
        scale = max(np.max(np.abs(Jt)), 1.0)
        for col in cols:
            V = np.zeros((B, nm))
            vp = np.zeros((B, 3))
            if col < nm:
                V[:, col] = 1.0
            else:
                vp[:, free[col - nm]] = 1.0
            ref = ko.matvec(M, p, V.reshape(M.shape), vp, cid, tuple(bool((cmask >> i) & 1) for i in range(3)))
            assert np.max(np.abs(Jt[:, col, :] - ref.reshape(B, nm))) < _tol(n, m) * scale, col
