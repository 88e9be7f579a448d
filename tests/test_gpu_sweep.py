"""GPU (-m gpu): the sweep driver on one GPU (the N>1 host logic is covered on CPU with gloo in test_sharding_gloo.py)."""
import numpy as np
import pytest

from oracle import ks_oracle as ko

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_hunt_sweep_single_gpu_records():
    from orbithunter_b200.sweep import RECORD_FIELDS, hunt_sweep

    rec = hunt_sweep("OrbitKS", 10, 44.0, 33.4, (32, 32), "adj", chunk=4, maxiter=50)
    assert rec.shape == (10, len(RECORD_FIELDS))
    rec = rec.cpu().numpy()
    assert np.array_equal(rec[:, 0], np.arange(10))
    for row in rec[:3]:
        M = ko.populate_state(44.0, 33.4, 32, 32, int(row[0]), ko.ORBIT)
        m1, p1, st = ko.adjoint_descent(M, np.array([44.0, 33.4, 0.0]), ko.ORBIT, maxiter=50)
        assert int(row[1]) == ko.hunt_status(st, st["cost"], 1e-6) and int(row[2]) == st["nit"]
        assert abs(row[3] - st["cost"]) < 1e-9 * max(1.0, st["cost"]) and np.allclose(row[4:7], p1, rtol=1e-10)


def test_full_size_sweep_properties():
    """BASELINE configs[4] size (65 536 seeds, 32x32) through size-independent properties: records sorted by seed, the cost
    never increases, every orbit either counts maxiter iterations or exits with a documented status, the run is
    deterministic, and chunking the seed range does not change a single bit."""
    from orbithunter_b200.sweep import hunt_sweep

    n_seeds, maxiter = 65536, 30
    a = hunt_sweep("OrbitKS", n_seeds, 44.0, 33.4, (32, 32), "adj", chunk=16384, maxiter=maxiter).cpu().numpy()
    b = hunt_sweep("OrbitKS", n_seeds, 44.0, 33.4, (32, 32), "adj", chunk=4096, maxiter=maxiter).cpu().numpy()
    assert a.shape == (n_seeds, 7) and np.array_equal(a, b)
    assert np.array_equal(a[:, 0], np.arange(n_seeds))
    status, nit, cost = a[:, 1], a[:, 2], a[:, 3]
    assert set(np.unique(status)) <= {-1.0, 0.0, 1.0, 2.0, 3.0}
    assert np.all((nit == maxiter) | (status != 2)) and np.all(nit <= maxiter) and np.all(np.isfinite(cost)) and np.all(cost >= 0)
    # spot-check against the oracle at both ends of the range
    for seed in (0, n_seeds - 1):
        M = ko.populate_state(44.0, 33.4, 32, 32, seed, ko.ORBIT)
        m1, p1, st = ko.adjoint_descent(M, np.array([44.0, 33.4, 0.0]), ko.ORBIT, maxiter=maxiter)
        assert int(status[seed]) == ko.hunt_status(st, st["cost"], 1e-6) and int(nit[seed]) == st["nit"]
        assert abs(cost[seed] - st["cost"]) < 1e-9 * max(1.0, st["cost"])


def test_two_stage_sweep_records():
    """sweep.two_stage_sweep (BASELINE configs[4], the bench's `sweep` block) on one GPU: one record per seed in seed order; the
    refined seeds carry the dense stage's outcome and a preprocess code, the others the adjoint stage's; every number equals
    what the two hunt() calls give when made directly."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt
    from orbithunter_b200.sweep import SWEEP_FIELDS, two_stage_sweep

    n_seeds, refine = 96, 24
    rec, tm = two_stage_sweep(n_seeds, refine, adj_maxiter=300, dense_maxiter=6, dense_chunk=16)
    assert rec.shape == (n_seeds, len(SWEEP_FIELDS)) and SWEEP_FIELDS[-1] == "code"
    r = rec.cpu().numpy()
    assert np.array_equal(r[:, 0], np.arange(n_seeds))
    assert np.all(r[refine:, 7] == -1) and set(np.unique(r[:refine, 7])) <= {0.0, 1.0, 2.0}
    assert tm["refined"] == refine and tm["seeds"] == n_seeds and tm["adjoint_steps"] == 300 * n_seeds and tm["dense_solves"] > 0
    b = BatchedOrbitKS.from_seeds(np.arange(n_seeds), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    r1 = hunt(b, "adj", maxiter=300, tol=1e-6)
    assert np.array_equal(r[refine:, 1], r1.status.cpu().numpy()[refine:]) and np.array_equal(r[refine:, 3], r1.costs[-1].cpu().numpy()[refine:])
    sub = r1.orbit._like(r1.orbit.state[:refine].contiguous(), r1.orbit.parameters[:refine].contiguous())
    r2 = hunt(sub, "lstsq", maxiter=6, tol=1e-10, chunk=16)
    assert np.array_equal(r[:refine, 1], r2.status.cpu().numpy())
    assert np.array_equal(r[:refine, 2], 300 + r2.nit.cpu().numpy())
    assert np.allclose(r[:refine, 3], r2.costs[-1].cpu().numpy(), rtol=1e-12) and np.allclose(r[:refine, 4:7], r2.orbit.cpu_parameters(), rtol=1e-13)
