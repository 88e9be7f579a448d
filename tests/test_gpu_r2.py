"""GPU (-m gpu): round-2 parity cases against outputs of the unmodified reference (tests/golden/ks_hunt_r2.npz and
ks_cfg0_table.npz, oracle/make_golden_r2.py): BASELINE configs[2] / configs[3] hunts (GMRES restart 50), the fully constrained
square GMRES branch, lsmr, standalone derivatives / preconditioning, and the configs[0] status table."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from oracle import ks_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(GOLDEN, "ks_hunt_r2.npz"))


def _cls(cid):
    from orbithunter_b200.batch import CLASS_NAMES
    from orbithunter_b200.ks import CLASSES

    return CLASSES[CLASS_NAMES[cid]]


def _batch(cid, M, p, cons=None):
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.batch import CLASS_NAMES

    c = None if cons is None else dict(zip("txs", (bool(x) for x in cons)))
    return BatchedOrbitKS(np.ascontiguousarray(M), np.ascontiguousarray(p), CLASS_NAMES[cid], "modes", constraints=c)


@pytest.mark.parametrize("cid", [1, 2, 3])
def test_cfg2_hunt_64x64_gmres50_matches_reference(g2, cid):
    """BASELINE configs[2]: 64x64 symmetry classes, (T, L) = (60, 44): 2000 adjoint steps, then Newton-GMRES (restart 50,
    4 cycles, rtol 1e-6, tol 1e-10), both seeds in one batch.  Identical status and nit; cost to 1e-9 relative (GMRES stage
    from the reference's own warm start) and 1e-8 (adjoint stage: 2000 data-dependent steps)."""
    from orbithunter_b200.optimize import hunt

    seeds = (0, 1)
    M0 = np.stack([g2[f"cfg2/{cid}_{s}/adj/in_state"] for s in seeds])
    p0 = np.stack([g2[f"cfg2/{cid}_{s}/adj/in_params"] for s in seeds])
    r1 = hunt(_batch(cid, M0, p0), "adj", maxiter=2000)
    for i, s in enumerate(seeds):
        key = f"cfg2/{cid}_{s}/adj"
        assert int(r1.status[i]) == int(g2[key + "/status"]) and int(r1.nit[i]) == int(g2[key + "/nit"])
        ref = float(g2[key + "/costs"][-1])
        assert abs(float(r1.costs[-1][i]) - ref) <= 1e-8 * ref
        assert relerr(r1.orbit.cpu_state()[i], g2[key + "/state"]) < 1e-7
    M1 = np.stack([g2[f"cfg2/{cid}_{s}/adj/state"] for s in seeds])
    p1 = np.stack([g2[f"cfg2/{cid}_{s}/adj/params"] for s in seeds])
    r2 = hunt(_batch(cid, M1, p1), "gmres", maxiter=3, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-6})
    for i, s in enumerate(seeds):
        key = f"cfg2/{cid}_{s}/gmres"
        assert int(r2.status[i]) == int(g2[key + "/status"]) and int(r2.nit[i]) == int(g2[key + "/nit"])
        ref = float(g2[key + "/costs"][-1])
        # 3 outer steps x 4 cycles x 50 modified-Gram-Schmidt Arnoldi steps on the normal equations: 1e-7 relative on a
        # cost of O(50) (measured 2e-9 .. 3e-8; the oracle, which calls SciPy's gmres like the reference, shows the same)
        assert abs(float(r2.costs[-1][i]) - ref) <= 1e-7 * max(1.0, ref)
        assert relerr(r2.orbit.cpu_state()[i], g2[key + "/state"]) < 1e-6
        assert np.allclose(r2.orbit.cpu_parameters()[i], g2[key + "/params"], rtol=1e-7)  # measured 2e-9 (same drift as the state)


def test_cfg3_hunt_256x256_gmres50_matches_reference(g2):
    """BASELINE configs[3]: OrbitKS 256x256, (T, L) = (200, 200), one Newton step of GMRES(50) (cluster GMRES step, tile path)."""
    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.optimize import hunt

    o = OrbitKS(parameters=(200.0, 200.0, 0.0)).populate(attr="state", seed=int(g2["cfg3/0/seed"]), discretization=(256, 256))
    r = hunt(o, "gmres", maxiter=1, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 1, "rtol": 1e-6})
    assert r.status == int(g2["cfg3/0/status"]) and r.nit == int(g2["cfg3/0/nit"])
    ref = g2["cfg3/0/costs"]
    assert abs(r.costs[0] - ref[0]) <= 1e-12 * ref[0] and abs(r.costs[-1] - ref[-1]) <= 1e-9 * ref[-1]
    sy, sx = (int(v) for v in g2["cfg3/0/sub"])
    assert relerr(r.orbit.state[::sy, ::sx], g2["cfg3/0/state"]) < 1e-8
    assert np.allclose(r.orbit.parameters, g2["cfg3/0/params"], rtol=1e-9)


@pytest.mark.parametrize("tag", ["0", "0_pc", "2", "2_pc", "1", "1_pc"])
def test_square_gmres_when_every_parameter_is_constrained(g2, tag):
    """cdof().size == eqn().size: the reference runs GMRES on J dx = -F, not on the normal equations (optimize.py:1803-1850)."""
    from orbithunter_b200.optimize import hunt

    cid, pc = int(tag[0]), tag.endswith("_pc")
    key = f"square/{tag}"
    o = _cls(cid)(state=g2[key + "/in_state"], basis="modes", parameters=tuple(g2[key + "/in_params"]),
                  constraints={"t": True, "x": True, "s": True})
    r = hunt(o, "gmres", maxiter=3, preconditioning=pc, scipy_kwargs={"restart": 20, "maxiter": 2})
    assert r.status == int(g2[key + "/status"]) and r.nit == int(g2[key + "/nit"])
    ref = float(g2[key + "/costs"][-1])
    assert abs(r.costs[-1] - ref) <= 1e-9 * max(1.0, ref) and relerr(r.orbit.state, g2[key + "/state"]) < 1e-8
    assert np.array_equal(np.asarray(r.orbit.parameters), g2[key + "/in_params"])


@pytest.mark.parametrize("cid", range(4))
def test_lsmr_and_lsqr_match_reference(g2, cid):
    """hunt('lsmr') (optimize.py:1319-1321) and hunt('lsqr') end states.  Short runs reproduce the reference's cost to 1e-9;
    the 300-step LSQR runs are round-off sensitive in the reference itself (tests/test_oracle_golden_r2.py measures SciPy
    against SciPy under 1e-15 input noise: 1e-4 .. 1e-2), so there status / nit must agree and the cost within 5e-2."""
    from orbithunter_b200.optimize import hunt

    key = f"lsq/{cid}/lsmr_maxiter50"
    o = _cls(cid)(state=g2[key + "/in_state"], basis="modes", parameters=tuple(g2[key + "/in_params"]))
    r = hunt(o, "lsmr", maxiter=3, scipy_kwargs={"maxiter": 50})
    assert r.status == int(g2[key + "/status"]) and r.nit == int(g2[key + "/nit"])
    ref = float(g2[key + "/costs"][-1])
    assert abs(r.costs[-1] - ref) <= 1e-9 * max(1.0, ref)
    assert relerr(r.orbit.state, g2[key + "/state"]) < 1e-8
    key = f"lsq/{cid}/lsqr_iter_lim300"
    r = hunt(o, "lsqr", maxiter=3, scipy_kwargs={"iter_lim": 300})
    assert r.status == int(g2[key + "/status"]) and r.nit == int(g2[key + "/nit"])
    ref = float(g2[key + "/costs"][-1])
    assert abs(r.costs[-1] - ref) <= 5e-2 * ref
    key = f"lsq/{cid}/lsmr"
    r = hunt(o, "lsmr", maxiter=3)  # default maxiter = min(m, n): 930 steps per outer iteration
    assert r.status == int(g2[key + "/status"]) and r.nit == int(g2[key + "/nit"])
    assert abs(r.costs[-1] - float(g2[key + "/costs"][-1])) <= 5e-2 * float(g2[key + "/costs"][-1])
    with pytest.raises(NotImplementedError):
        hunt(o, "lsmr", maxiter=1, scipy_kwargs={"damp": 0.1})


@pytest.mark.parametrize("cid", range(4))
def test_derivatives_and_preconditioning_on_device(g2, cid):
    """ks_deriv / ks_precondition on the sm_100a binary (not the CPU emulation): OrbitKS.dt / dx orders 1-4 with the SR / Anti
    odd orders through spatial_modes, return_basis, precondition and preconditioner(), against the reference's outputs."""
    for (n, m) in ((32, 32), (16, 24)):
        key = f"deriv/{cid}_{n}x{m}"
        M, p = g2[key + "/modes"], g2[key + "/params"]
        o = _cls(cid)(state=M, basis="modes", parameters=tuple(p))
        b = _batch(cid, np.stack([M, 2.0 * M]), np.stack([p, p]))
        for order in (1, 2, 3, 4):
            ref = g2[key + f"/dt{order}"]
            sc = max(float(np.max(np.abs(ref))), 1e-300)
            assert relerr(o.dt(order=order).state, ref, scale=sc) < 1e-13
            got = b.dt(order).cpu_state()
            assert relerr(got[0], ref, scale=sc) < 1e-13 and relerr(got[1], 2.0 * ref, scale=sc) < 1e-13
            ref = g2[key + f"/dx{order}"]
            sc = max(float(np.max(np.abs(ref))), float(np.max(np.abs(M))))
            assert relerr(o.dx(order=order).state, ref, scale=sc) < 1e-13
            assert relerr(b.dx(order).cpu_state()[0], ref, scale=sc) < 1e-13
            ref = g2[key + f"/dx{order}_field"]
            assert relerr(o.dx(order=order, return_basis="field").state, ref, scale=max(float(np.max(np.abs(ref))), 1e-300)) < 1e-12
        # the reference's non-in-place dt ignores return_basis (ks/orbits.py:435-440): the golden is in the modes basis
        d = o.dt(return_basis="field")
        assert d.basis == "modes" and relerr(d.state, g2[key + "/dt1_field"]) < 1e-13
        oi = o.copy()
        assert oi.dt(inplace=True, return_basis="field") is oi and oi.basis == "field"
        assert relerr(oi.state, o.dt().transform(to="field").state) < 1e-13
        g = o.rmatvec(o)
        assert relerr(g.state, g2[key + "/rmatvec_self"]) < 1e-12
        pc = g.precondition(pmult=o.preconditioning_parameters())
        assert relerr(pc.state, g2[key + "/precond"]) < 1e-12 and np.allclose(pc.parameters, g2[key + "/precond_params"], rtol=1e-12)
        P = o.preconditioner()
        assert relerr(np.asarray(P.matvec(np.ones(P.shape[0]))).ravel(), g2[key + "/preconditioner_diag"]) < 1e-13
        bg = b.rmatvec(b)
        bp = bg.precondition(b.parameters)
        assert relerr(bp.cpu_state()[0], g2[key + "/precond"]) < 1e-12


def test_newton_descent_step_is_the_pseudo_inverse_step():
    """'newton_descent' restated matrix-free: ONE outer iteration from a point where the full step is accepted and from one
    where it is not -- in both cases the returned orbit must be the reference logic's (optimize.py:785-815: dx = pinv(J)(-F),
    halve while the cost increases)."""
    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.optimize import hunt

    checked = 0
    for seed, scale in ((3, 1.0), (3, 0.05), (5, 0.2)):
        M = scale * ko.populate_state(44.0, 44.0, 16, 16, seed, 0)
        o = OrbitKS(state=M, basis="modes", parameters=(44.0, 44.0, 0.0))
        J = o.jacobian()
        dx = o.from_numpy_array(np.linalg.pinv(J) @ (-o.eqn().state.ravel()))
        c0, step = o.cost(), 1.0
        nxt = o.increment(dx, step_size=step)
        while nxt.cost() > c0 and step > 1e-6:
            step /= 2.0
            nxt = o.increment(dx, step_size=step)
        r = hunt(o, "newton_descent", maxiter=1)
        assert r.nit == 1
        if step > 1e-6:  # a decreasing step exists: status 2 (maxiter) returns the better orbit, which is the trial
            assert relerr(r.orbit.state, nxt.state) < 1e-6 and abs(r.costs[-1] - nxt.cost()) < 1e-6 * max(1.0, nxt.cost())
            assert np.allclose(r.orbit.parameters, nxt.parameters, rtol=1e-6)
            checked += 1
        else:
            assert r.status == 0 and np.array_equal(r.orbit.state, o.state)
    assert checked >= 2


def test_cfg0_status_table_matches_reference():
    """BASELINE configs[0] (hunt(('adj', 'lstsq'), maxiter=[10000, 200], tol=[1e-6, 1e-10]), OrbitKS 32x32, (T, L) = (44, 33.4))
    on seeds 0..191 against the unmodified reference's table (tests/golden/ks_cfg0_table.npz, oracle/make_golden_r2.py --only
    cfg0).  The adjoint stage must agree exactly (status, nit; cost to 1e-7).  The dense least-squares stage is Newton's method
    started far from a solution: the reference truncates singular values in gelsd where the device factors J J^T, iterates
    differ in the last bits and the iteration amplifies that, so a per-seed exit code cannot be pinned.  Measured on these 192
    seeds: the reference converges 29, the device 28-29, 26-27 of them the same seeds (the others need 180+ of the 200
    iterations in the reference and run out here; which ones depends on the summation order of the factorisation -- the
    rank-64 and the two-level rank-256 trailing updates differ in one seed), 89-91 % of the exit codes coincide.  Asserted: the
    same converged fraction (+-15 %), >= 85 % of the reference's converged seeds converge here, >= 85 % identical codes,
    converged costs below tol, and the common converged orbits lie on the same families ((T, L) to 1e-3 for >= 85 % of them)."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    tab = np.load(os.path.join(GOLDEN, "ks_cfg0_table.npz"))["table"]
    seeds = tab[:, 0].astype(int)
    assert len(seeds) >= 192
    b = BatchedOrbitKS.from_seeds(seeds, (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    r1 = hunt(b, "adj", maxiter=10000, tol=1e-6)
    assert np.array_equal(r1.status.cpu().numpy(), tab[:, 1].astype(int))
    assert np.array_equal(r1.nit.cpu().numpy(), tab[:, 2].astype(int))
    assert np.allclose(r1.costs[-1].cpu().numpy(), tab[:, 3], rtol=1e-7)
    r2 = hunt(r1.orbit, "lstsq", maxiter=200, tol=1e-10)
    got, ref = r2.status.cpu().numpy(), tab[:, 4].astype(int)
    confusion = {(int(a), int(c)): int(np.sum((ref == a) & (got == c))) for a in (0, 1, 2, 3) for c in (0, 1, 2, 3)}
    conv, mine = ref == 1, got == 1
    both = conv & mine
    assert abs(int(mine.sum()) - int(conv.sum())) <= 0.15 * conv.sum(), confusion
    assert both.sum() >= 0.85 * conv.sum(), confusion
    assert np.mean(got == ref) >= 0.85, confusion
    assert np.all(r2.costs[-1].cpu().numpy()[mine] <= 1e-10)
    P = r2.orbit.cpu_parameters()
    same_family = np.isclose(P[both, 0], tab[both, 7], rtol=1e-3) & np.isclose(P[both, 1], tab[both, 8], rtol=1e-3)
    assert same_family.mean() >= 0.85, (P[both, :2], tab[both, 7:9])
