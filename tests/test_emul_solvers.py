"""CPU: the batched solver drivers (C++ host loops + decision kernels) under the CPU-thread emulation vs the oracle's
restatement of optimize.py -- same status / nit, cost to 1e-10, for every class, with and without preconditioning."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_lib as E  # noqa: E402
from conftest import relerr  # noqa: E402
from oracle import ks_oracle as ko  # noqa: E402


@pytest.fixture(params=[128, 7, 0], ids=["multi128", "multi7", "perpass"])
def adj_inner(request):
    """32x32 adjoint descent kernels: multi-pass (state on chip, 128 or 7 passes per launch) and one pass per launch."""
    E.lib().ks_debug_adjoint_inner(request.param)
    yield request.param
    E.lib().ks_debug_adjoint_inner(128)


@pytest.mark.parametrize("cid", range(4))
def test_adjoint_descent_matches_oracle(cid, adj_inner):
    n = m = 32
    B = 3
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    for pc, maxiter, tol in ((False, 40, 1e-6), (True, 25, 1e-6), (False, 30, 600.0)):
        Mo, po, status, nit, cost, step, passes = E.adjoint_descent(cid, n, m, M, p, cmask, tol=tol, maxiter=maxiter, precond=pc)
        for b in range(B):
            rm, rp, st = ko.adjoint_descent(M[b], p[b], cid, maxiter=maxiter, tol=tol, preconditioning=pc)
            assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"]), (cid, pc, b)
            assert abs(cost[b] - st["cost"]) <= 1e-10 * max(1.0, st["cost"])
            assert step[b] == st["step_size"]
            assert relerr(Mo[b], rm) < 1e-10 and np.allclose(po[b], rp, rtol=1e-11)
        assert passes >= int(nit.max())


def test_adjoint_descent_exit_codes(adj_inner):
    """min-step (0), maxiter (2), insufficient decrease (3), already-converged (-1 with nit 0) per orbit in one batch."""
    n = m = 32
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3, 0), 1e-9 * ko.populate_state(44.0, 44.0, n, m, 4, 0)])
    p = np.tile([44.0, 44.0, 0.0], (2, 1))
    # orbit 1 starts below tol: loop never entered
    Mo, po, status, nit, cost, step, _ = E.adjoint_descent(0, n, m, M, p, 4, maxiter=5)
    assert status[1] == -1 and nit[1] == 0 and np.array_equal(Mo[1], M[1]) and status[0] == 2 and nit[0] == 5
    # huge ftol -> status 3 on the first counted step, orbit unchanged
    Mo, po, status, nit, cost, step, _ = E.adjoint_descent(0, n, m, M[:1], p[:1], 4, maxiter=50, ftol=0.9)
    rm, rp, st = ko.adjoint_descent(M[0], p[0], 0, maxiter=50, ftol=0.9)
    assert status[0] == 3 == st["status"] and nit[0] == st["nit"] and relerr(Mo[0], rm) < 1e-12
    # min_step just under the first accepted step size -> status 0
    rm, rp, st = ko.adjoint_descent(M[0], p[0], 0, maxiter=50, min_step=0.2)
    Mo, po, status, nit, cost, step, _ = E.adjoint_descent(0, n, m, M[:1], p[:1], 4, maxiter=50, min_step=0.2)
    assert (int(status[0]), int(nit[0])) == (st["status"], st["nit"]) and st["status"] == 0
    assert relerr(Mo[0], rm) < 1e-12


@pytest.mark.parametrize("cid,method,precond,gm_cluster", [(1, 0, False, 1), (2, 0, True, 4), (0, 0, True, 0), (0, 0, False, 2),
                                                          (0, 1, False, 1)])
def test_newton_krylov_matches_scipy_based_oracle(cid, method, precond, gm_cluster):
    """Device GMRES (normal equations, optional diagonal M) / LSQR vs the oracle, which calls SciPy's own solvers as
    the reference does: same outer status / nit; GMRES trajectories agree to round-off, LSQR to its own sensitivity."""
    n = m = 16
    B = 2
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    if method == 0:
        kw, okw = dict(restart=10, inner=2, rtol=1e-5), dict(restart=10, maxiter=2)
    else:
        kw, okw = dict(inner=30, atol=1e-6), dict(iter_lim=30)
    E.lib().ks_debug_gmres_cluster(gm_cluster)  # GMRES step: 0 one CTA per orbit, 1 auto, 2/4 CTAs per orbit (cluster)
    try:
        Mo, po, status, nit, cost, cnt = E.newton_krylov(cid, n, m, M, p, cmask, method, maxiter=2, precond=precond, **kw)
    finally:
        E.lib().ks_debug_gmres_cluster(1)
    for b in range(B):
        rm, rp, st = ko.newton_krylov(M[b], p[b], cid, method="gmres" if method == 0 else "lsqr", maxiter=2,
                                      preconditioning=precond, scipy_kwargs=okw)
        assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"])
        tol = 1e-10 if method == 0 else 1e-4
        assert abs(cost[b] - st["cost"]) <= tol * max(1.0, st["cost"])
        assert relerr(Mo[b], rm) < tol * 10
    assert cnt[3] == 2 and cnt[0] > 0


@pytest.mark.parametrize("cid,precond", [(0, False), (2, True), (1, False)])
def test_gmres_square_system_when_every_parameter_is_constrained(cid, precond):
    """cdof().size == eqn().size (t, x and s all constrained, continuation's case): the reference hands GMRES the square
    system J dx = -F itself, not the normal equations (optimize.py:1764, :1803-1850)."""
    n = m = 16
    B = 2
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = (True, True, True)
    Mo, po, status, nit, cost, cnt = E.newton_krylov(cid, n, m, M, p, 7, 0, maxiter=2, precond=precond, restart=10, inner=2, rtol=1e-5)
    for b in range(B):
        rm, rp, st = ko.newton_krylov(M[b], p[b], cid, constraints=cons, method="gmres", maxiter=2, preconditioning=precond,
                                      scipy_kwargs=dict(restart=10, maxiter=2))
        assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"])
        assert abs(cost[b] - st["cost"]) <= 1e-10 * max(1.0, st["cost"])
        assert relerr(Mo[b], rm) < 1e-9 and np.array_equal(po[b], p[b])


@pytest.mark.parametrize("cid", [0, 1, 3])
def test_lsmr_matches_scipy_based_oracle(cid):
    """hunt('lsmr') (optimize.py:1319-1321): the device LSMR recurrences (SciPy _isolve/lsmr.py restated) vs the oracle,
    which calls scipy.sparse.linalg.lsmr as the reference does."""
    n = m = 16
    B = 2
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    # 20 Golub-Kahan steps: the trajectories agree to round-off (at 40 steps and N ~ 200 orthogonality is lost and the same
    # comparison drifts to 3e-6, as it does between SciPy and itself under 1e-15 input noise)
    Mo, po, status, nit, cost, cnt = E.newton_krylov(cid, n, m, M, p, cmask, 2, maxiter=2, inner=20, atol=1e-6, btol=1e-6)
    for b in range(B):
        rm, rp, st = ko.newton_krylov(M[b], p[b], cid, method="lsmr", maxiter=2, scipy_kwargs=dict(maxiter=20))
        assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"])
        assert abs(cost[b] - st["cost"]) <= 1e-10 * max(1.0, st["cost"])
        assert relerr(Mo[b], rm) < 1e-9


@pytest.mark.parametrize("cid", [1])
def test_dense_newton_matches_lstsq_oracle(cid):
    """hunt('lstsq') on the device (ks_newton_dense: Jacobian by batched matvec, blocked Cholesky of J J^T, refinement) vs the
    oracle's restatement of optimize.py:855-1022 with scipy.linalg.lstsq: same status / nit, cost to 1e-8 -- two orbits in two
    chunks (chunk = 1), two Cholesky blocks (110 modes).  (The thread emulation of the 256-thread factorisation kernels is slow;
    all classes and 16x16 pass the same check in 11 min, and the -m gpu tests cover 32x32.)"""
    n = m = 12
    B = 2
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    try:
        for outer in (4, 1):  # two-level factorisation: inner (rank 64, panel columns only) and whole-matrix updates
            E.lib().ks_debug_dense_outer_blocks(outer)
            # (the emulated factorisation kernels are slow: two orbits in two chunks for the default, the second orbit alone
            # for the other variant)
            sel = slice(0, B) if outer == 4 else slice(1, 2)
            nb = sel.stop - sel.start
            Mo, po, status, nit, cost, cnt = E.newton_dense(cid, n, m, M[sel], p[sel], cmask, maxiter=1, tol=1e-10, chunk=1)
            for b in range(nb):
                rm, rp, st = ko.newton_krylov(M[sel][b], p[sel][b], cid, method="lstsq", maxiter=1, tol=1e-10)
                assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"]), (b, status, nit, st)
                assert abs(cost[b] - st["cost"]) <= 1e-8 * max(1.0, st["cost"]), (b, cost[b], st["cost"])
                assert relerr(Mo[b], rm) < 1e-6 and np.allclose(po[b], rp, rtol=1e-7)
            assert cnt[3] >= 1 and cnt[1] >= nb
    finally:
        E.lib().ks_debug_dense_outer_blocks(4)


@pytest.mark.parametrize("cid,method", [(2, 0), (1, 1)])
def test_operator_cache_of_the_64x64_krylov_driver_changes_no_bit(cid, method):
    """Newton-GMRES / LSQR at 64x64 (CTA-resident kernel): with the operator cache (N(u, u) computed once per outer iteration,
    the two forward transforms of u u skipped in every matvec / rmatvec; opt-in, ks_debug_operator_cache) and without it the
    driver must return identical bits -- the cached tuples are the very values the uncached kernel recomputes."""
    n = m = 64
    B = 2
    M = np.stack([ko.populate_state(60.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([60.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    kw = dict(restart=3, inner=1, rtol=1e-5) if method == 0 else dict(inner=4, atol=1e-6)
    out = {}
    try:
        for on in (1, 0):
            E.lib().ks_debug_operator_cache(on)
            out[on] = E.newton_krylov(cid, n, m, M, p, cmask, method, maxiter=1, **kw)
    finally:
        E.lib().ks_debug_operator_cache(0)
    for a, b in zip(out[1][:5], out[0][:5]):
        assert np.array_equal(a, b)
    assert out[1][5][0] == out[0][5][0] > 0  # the same number of operator applications
    rm, rp, st = ko.newton_krylov(M[0], p[0], cid, method="gmres" if method == 0 else "lsqr", maxiter=1,
                                  scipy_kwargs=dict(restart=3, maxiter=1) if method == 0 else dict(iter_lim=4))
    assert abs(out[1][4][0] - st["cost"]) <= 1e-9 * st["cost"] and relerr(out[1][0][0], rm) < 1e-9


@pytest.mark.parametrize("cid", [0, 1])
def test_lsqr_at_32x32_runs_matvec_and_rmatvec_on_the_one_orbit_per_warp_kernel(cid):
    """32x32 Newton-LSQR: the operator applications go through ks_warp32x1_kernel<MATVEC / RMATVEC> with the cdof layouts of the
    Krylov driver (v and the free parameters in one vector of stride N, results likewise) -- against the SciPy-based oracle."""
    n = m = 32
    B = 3  # more orbits than one round of the emulated grid's first warps; odd
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 3 + s, cid) for s in range(B)])
    p = np.tile([44.0, 44.0, 0.3 if cid == 1 else 0.0], (B, 1))
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    Mo, po, status, nit, cost, cnt = E.newton_krylov(cid, n, m, M, p, cmask, 1, maxiter=1, inner=4, atol=1e-6, btol=1e-6)
    for b in range(B):
        rm, rp, st = ko.newton_krylov(M[b], p[b], cid, method="lsqr", maxiter=1, scipy_kwargs=dict(iter_lim=4))
        assert (int(status[b]), int(nit[b])) == (st["status"], st["nit"])
        assert abs(cost[b] - st["cost"]) <= 1e-9 * max(1.0, st["cost"])
        assert relerr(Mo[b], rm) < 1e-8 and np.allclose(po[b], rp, rtol=1e-8)
    assert cnt[0] > 0


@pytest.mark.parametrize("cid,modes", [(0, (2, 1)), (1, (2,)), (3, (2,))])
def test_dense_newton_32x32_matrix_free_gram_matches_lstsq_oracle(cid, modes):
    """hunt('lstsq') at 32x32: no Jacobian is stored -- J J^T is built column by column as J (J^T e_j) by the one-orbit-per-warp
    kernel (rmatvec over unit vectors made in the kernel, then matvec over the result), and the solve's J^T y / J dx are single
    rmatvec / matvec launches.  One Newton step of one orbit against the oracle's scipy.linalg.lstsq step (OrbitKS, Relative
    with its three free parameters, Antisymmetric with its half-width modes), and for OrbitKS also against the Jacobian + GEMM
    route (ks_debug_dense_jacobian_mode(1))."""
    n = m = 32
    M = np.stack([ko.populate_state(44.0, 33.4, n, m, 3, cid)])
    p = np.array([[44.0, 33.4, 0.3 if cid == 1 else 0.0]])
    cons = ko.default_constraints(cid)
    cmask = sum(1 << i for i in range(3) if cons[i])
    out = {}
    try:
        for mode in modes:
            E.lib().ks_debug_dense_jacobian_mode(mode)
            out[mode] = E.newton_dense(cid, n, m, M, p, cmask, maxiter=1, tol=1e-10, chunk=1)
    finally:
        E.lib().ks_debug_dense_jacobian_mode(2)
    rm, rp, st = ko.newton_krylov(M[0], p[0], cid, method="lstsq", maxiter=1, tol=1e-10)
    for mode in modes:
        Mo, po, status, nit, cost, cnt = out[mode]
        assert (int(status[0]), int(nit[0])) == (st["status"], st["nit"])
        assert abs(cost[0] - st["cost"]) <= 1e-8 * max(1.0, st["cost"]), (mode, cost[0], st["cost"])
        assert relerr(Mo[0], rm) < 1e-6 and np.allclose(po[0], rp, rtol=1e-7)
    if len(modes) > 1:
        assert relerr(out[2][0][0], out[1][0][0]) < 1e-9
