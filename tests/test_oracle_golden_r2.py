"""CPU: the oracle's solver loops against round-2 goldens of the unmodified reference (tests/golden/ks_hunt_r2.npz,
oracle/make_golden_r2.py): BASELINE configs[2] (64x64 symmetry classes, adjoint warm-up then Newton-GMRES(50)) and
configs[3] (256x256, GMRES(50)) settings, and the fully constrained square GMRES branch."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from oracle import ks_oracle as ko


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(GOLDEN, "ks_hunt_r2.npz"))


@pytest.mark.parametrize("cid", [1, 2, 3])
def test_cfg2_gmres50_from_reference_warm_start(g2, cid):
    key = f"cfg2/{cid}_0"
    m0, p0 = g2[key + "/adj/state"], g2[key + "/adj/params"]
    rm, rp, st = ko.newton_krylov(m0, p0, cid, method="gmres", maxiter=3, tol=1e-10,
                                  scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-6})
    assert ko.hunt_status(st, st["cost"], 1e-10) == int(g2[key + "/gmres/status"]) and st["nit"] == int(g2[key + "/gmres/nit"])
    ref = float(g2[key + "/gmres/costs"][-1])
    assert abs(st["cost"] - ref) <= 1e-9 * max(1.0, ref)
    assert relerr(rm, g2[key + "/gmres/state"]) < 1e-8 and np.allclose(rp, g2[key + "/gmres/params"], rtol=1e-9)


def test_cfg2_adjoint_warm_up(g2):
    key = "cfg2/2_0"
    rm, rp, st = ko.adjoint_descent(g2[key + "/adj/in_state"], g2[key + "/adj/in_params"], 2, maxiter=2000)
    assert (st["status"], st["nit"]) == (int(g2[key + "/adj/status"]), int(g2[key + "/adj/nit"]))
    ref = float(g2[key + "/adj/costs"][-1])
    assert abs(st["cost"] - ref) <= 1e-9 * ref and relerr(rm, g2[key + "/adj/state"]) < 1e-8


def test_cfg3_gmres50_256(g2):
    M = ko.populate_state(200.0, 200.0, 256, 256, int(g2["cfg3/0/seed"]), ko.ORBIT)
    rm, rp, st = ko.newton_krylov(M, np.array([200.0, 200.0, 0.0]), ko.ORBIT, method="gmres", maxiter=1, tol=1e-10,
                                  scipy_kwargs={"restart": 50, "maxiter": 1, "rtol": 1e-6})
    assert ko.hunt_status(st, st["cost"], 1e-10) == int(g2["cfg3/0/status"]) and st["nit"] == int(g2["cfg3/0/nit"])
    ref = float(g2["cfg3/0/costs"][-1])
    assert abs(st["cost"] - ref) <= 1e-9 * ref
    sy, sx = (int(v) for v in g2["cfg3/0/sub"])
    assert relerr(rm[::sy, ::sx], g2["cfg3/0/state"]) < 1e-8 and np.allclose(rp, g2["cfg3/0/params"], rtol=1e-9)


@pytest.mark.parametrize("tag", ["0", "0_pc", "2", "2_pc", "1", "1_pc"])
def test_square_gmres_branch(g2, tag):
    cid, pc = int(tag[0]), tag.endswith("_pc")
    key = f"square/{tag}"
    rm, rp, st = ko.newton_krylov(g2[key + "/in_state"], g2[key + "/in_params"], cid, constraints=(True, True, True), method="gmres",
                                  maxiter=3, preconditioning=pc, scipy_kwargs={"restart": 20, "maxiter": 2})
    assert ko.hunt_status(st, st["cost"], 1e-6) == int(g2[key + "/status"]) and st["nit"] == int(g2[key + "/nit"])
    ref = float(g2[key + "/costs"][-1])
    assert abs(st["cost"] - ref) <= 1e-9 * max(1.0, ref) and relerr(rm, g2[key + "/state"]) < 1e-8
