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
