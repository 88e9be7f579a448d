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
