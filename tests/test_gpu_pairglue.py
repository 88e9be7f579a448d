"""GPU (-m gpu): symmetry operations, concat, group orbits and the pairwise group-orbit gluing search against outputs of the
unmodified reference (tests/golden/ks_pairglue.npz, oracle/make_golden_pairglue.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
CLASS_NAMES = ("OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS")


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(GOLDEN, "ks_pairglue.npz"))


def _pair(gp, cid):
    import orbithunter_b200 as ob

    C = getattr(ob, CLASS_NAMES[cid])
    a = C(state=gp[f"{cid}/a"], basis="field", parameters=tuple(gp[f"{cid}/a_params"]))
    b = C(state=gp[f"{cid}/b"], basis="field", parameters=tuple(gp[f"{cid}/b_params"]))
    return a, b


@pytest.mark.parametrize("cid", range(4))
def test_roll_cell_shift_reflection_concat(gp, cid):
    a, b = _pair(gp, cid)
    assert relerr(a.roll(3, axis=0).roll(-5, axis=1).state, gp[f"{cid}/roll"]) < 1e-13
    assert relerr(a.cell_shift(2, axis=0).state, gp[f"{cid}/cell_shift_t"]) < 1e-13
    assert relerr(a.cell_shift(-4, axis=1).state, gp[f"{cid}/cell_shift_x"]) < 1e-13
    r = a.reflection()
    assert relerr(r.state, gp[f"{cid}/reflection"]) < 1e-13 and np.allclose(r.parameters, gp[f"{cid}/reflection_params"])
    rm = a.transform(to="modes").reflection()
    assert rm.basis == "modes" and relerr(rm.state, gp[f"{cid}/reflection_modes"]) < 1e-12
    for axis in (0, 1):
        c = a.concat(b, axis=axis)
        assert c.state.shape == gp[f"{cid}/concat{axis}"].shape and tuple(c.discretization) == tuple(gp[f"{cid}/concat{axis}_disc"])
        assert relerr(c.state, gp[f"{cid}/concat{axis}"]) < 1e-13
        assert np.allclose(c.parameters, gp[f"{cid}/concat{axis}_params"], rtol=1e-14)


@pytest.mark.parametrize("cid", range(4))
def test_group_orbit_members_and_device_tensor(gp, cid):
    """The generator (single-orbit API) and the device tensor the pairwise search is built on give the reference's members in
    the reference's order."""
    from orbithunter_b200.gluing import _group_fields

    a, _ = _pair(gp, cid)
    for tag, kw in (("default", dict(rolls=(8, 4))), ("continuous", dict(continuous=True, rolls=(4, 8))), ("discrete", dict(discrete=True))):
        want, wp = gp[f"{cid}/group/{tag}"], gp[f"{cid}/group/{tag}_params"]
        G, sgn = _group_fields(a, kw)
        assert tuple(G.shape) == want.shape
        assert relerr(G.cpu().numpy(), want) < 1e-13
        assert np.allclose(sgn.cpu().numpy() * a.parameters[2], wp[:, 2])
        members = list(a.group_orbit(**kw))
        assert len(members) == len(want)
        for g, w, p in zip(members, want, wp):
            assert relerr(g.state, w) < 1e-13 and np.allclose(g.parameters, p)


@pytest.mark.parametrize("cid", range(4))
def test_expensive_pairwise_glue_matches_reference(gp, cid):
    """gluing.expensive_pairwise_glue (reference gluing.py:424-496): 1024 (rolls of 4) and 64 (discrete) combinations per
    call, scored on the device in one batch; the winner's cost, parameters and norm are the reference's."""
    from orbithunter_b200.gluing import expensive_pairwise_glue

    a, b = _pair(gp, cid)
    for axis in (0, 1):
        for tag, kw in (("default", dict(rolls=(4, 4))), ("discrete", dict(discrete=True))):
            arr = np.array([a, b], dtype=object).reshape((2, 1) if axis == 0 else (1, 2))
            best = expensive_pairwise_glue(arr, objective="cost", axis=axis, **kw)
            key = f"{cid}/pair/{tag}{axis}"
            assert type(best) is type(a) and best.basis == "field"
            assert best.state.shape == gp[key].shape
            # translating both tiles along the axis they are NOT joined on translates the joined orbit, so whole families of
            # combinations share the minimal cost up to rounding: the winner is pinned by its cost, not by its array
            assert best.cost() == pytest.approx(float(gp[key + "_cost"]), rel=1e-11), key
            assert np.allclose(np.abs(best.parameters), np.abs(gp[key + "_params"]), rtol=1e-14), key
            assert np.linalg.norm(best.state) == pytest.approx(np.linalg.norm(gp[key]), rel=1e-12), key
            small = expensive_pairwise_glue(arr, objective="cost", axis=axis, chunk=100, **kw)  # several chunks
            assert small.cost() == pytest.approx(best.cost(), rel=1e-11)


def test_pairwise_boundary_cost_and_generator(gp):
    """objective="boundary_cost" (the reference raises AttributeError there; the evident intent is implemented): the winner
    minimises the seam distance computed independently on the host; pairwise_group_orbit yields |Ga| x |Gb| pairs."""
    from orbithunter_b200.gluing import expensive_pairwise_glue, pairwise_group_orbit

    a, b = _pair(gp, 0)
    kw = dict(rolls=(4, 4))
    pairs = list(pairwise_group_orbit([a, b], **kw))
    assert len(pairs) == 32 * 32
    for axis in (0, 1):
        best = expensive_pairwise_glue([a, b], objective="boundary_cost", axis=axis, **kw)
        dist = []
        for ga, gb in pairs:
            A, B = ga.state, gb.state
            if axis == 0:
                dist.append(np.sqrt(np.sum((A[-1] - B[0]) ** 2) + np.sum((A[0] - B[-1]) ** 2)))
            else:
                dist.append(np.sqrt(np.sum((A[:, -1] - B[:, 0]) ** 2) + np.sum((A[:, 0] - B[:, -1]) ** 2)))
        ga, gb = pairs[int(np.argmin(dist))]
        assert relerr(best.state, ga.concat(gb, axis=axis).state) < 1e-13


@pytest.mark.parametrize("cid", range(4))
def test_slicing_and_fundamental_domains(gp, cid):
    """Orbit.__getitem__ (core.py:553-668), RelativeOrbitKS.change_reference_frame (ks/orbits.py:2770-2833) and the
    fundamental domains of the discrete-symmetry classes (:3229-3252, :3688-3711) against the reference's outputs, in the
    field and in the modes basis, and group_orbit(discrete=True, fundamental_domain=True)."""
    a, _ = _pair(gp, cid)
    cut = a[2:10, 4:16]
    assert np.array_equal(cut.state, gp[f"{cid}/getitem"]) and np.allclose(cut.parameters, gp[f"{cid}/getitem_params"], rtol=1e-14)
    assert tuple(cut.discretization) == (8, 12)
    assert np.allclose(a[3:4, :].parameters, gp[f"{cid}/getitem_row_params"])
    with pytest.raises(ValueError):
        a[3]
    if cid == 0:
        assert a.to_fundamental_domain() is a and a.from_fundamental_domain() is a
        return
    for half in ((0,) if cid == 1 else (0, 1)):
        kw = {} if cid == 1 else {"half": half}
        fd = a.to_fundamental_domain(**kw)
        assert fd.state.shape == gp[f"{cid}/to_fd{half}"].shape
        assert relerr(fd.state, gp[f"{cid}/to_fd{half}"]) < 1e-12 and np.allclose(fd.parameters, gp[f"{cid}/to_fd{half}_params"], rtol=1e-14)
        back = fd.from_fundamental_domain(**kw)
        assert relerr(back.state, gp[f"{cid}/from_fd{half}"]) < 1e-12
        assert np.allclose(back.parameters, gp[f"{cid}/from_fd{half}_params"], rtol=1e-14)
        fdm = a.transform(to="modes").to_fundamental_domain(**kw)
        assert fdm.basis == "modes" and relerr(fdm.state, gp[f"{cid}/to_fd{half}_modes"]) < 1e-12
    if cid == 1:
        assert fd.frame == "physical" and back.frame == "comoving" and relerr(back.state, a.state) < 1e-12
    members = list(a.group_orbit(discrete=True, fundamental_domain=True))
    want, wp = gp[f"{cid}/group/discrete_fd"], gp[f"{cid}/group/discrete_fd_params"]
    assert len(members) == len(want)
    for g, w, p in zip(members, want, wp):
        assert relerr(g.state, w) < 1e-12 and np.allclose(g.parameters, p)


@pytest.mark.parametrize("cid", [2, 3])
def test_pairwise_glue_over_fundamental_domains(gp, cid):
    """expensive_pairwise_glue(fundamental_domain=True) for the discrete-symmetry classes: members cut to their fundamental
    domains, joined, completed with the reflection, scored -- all combinations of a call in one device batch; winner's cost,
    parameters and norm are the reference's (ties as in test_expensive_pairwise_glue_matches_reference)."""
    from orbithunter_b200.gluing import expensive_pairwise_glue

    a, b = _pair(gp, cid)
    for axis in (0, 1):
        for tag, kw in (("default", dict(rolls=(4, 4))), ("discrete", dict(discrete=True))):
            arr = np.array([a, b], dtype=object).reshape((2, 1) if axis == 0 else (1, 2))
            best = expensive_pairwise_glue(arr, objective="cost", axis=axis, fundamental_domain=True, **kw)
            key = f"{cid}/pairfd/{tag}{axis}"
            assert type(best) is type(a) and best.state.shape == gp[key].shape
            assert best.cost() == pytest.approx(float(gp[key + "_cost"]), rel=1e-11), key
            assert np.allclose(np.abs(best.parameters), np.abs(gp[key + "_params"]), rtol=1e-14), key
            assert np.linalg.norm(best.state) == pytest.approx(np.linalg.norm(gp[key]), rel=1e-12), key
    r = _pair(gp, 1)
    with pytest.raises(NotImplementedError):
        expensive_pairwise_glue(list(r), fundamental_domain=True, discrete=True)
