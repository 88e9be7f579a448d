"""CPU: seed modulations, rescale, resize, dimension_based_discretization and the preprocess norms against outputs of the
unmodified reference (tests/golden/ks_aux_r3.npz, oracle/make_golden_r3.py) -- the host-side class logic, and the CUDA
kernel sources (ks_populate.cuh, ks_aux.cuh) compiled for the CPU-thread shim (tests/emul).  The same cases run on the
sm_100a binary in tests/test_gpu_r3.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_lib as E  # noqa: E402
from conftest import GOLDEN, relerr  # noqa: E402

CLASS_NAMES = ("OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS")


@pytest.fixture(scope="module")
def g3():
    return np.load(os.path.join(GOLDEN, "ks_aux_r3.npz"))


def _pop_keys(g3):
    for key in g3.files:
        if key.startswith("pop/") and key != "pop/override":
            _, tag, sm, tm, xs = key.split("/")
            cid, shape = tag.split("_")
            n, m = (int(v) for v in shape.split("x"))
            yield key, int(cid), n, m, sm, tm, xs


def test_host_populate_modulations_match_reference(g3):
    """OrbitKS._populate_state with every spatial / temporal profile and xshift (ks/orbits.py:1772-1842), host classes."""
    from orbithunter_b200.ks import CLASSES

    count = 0
    for key, cid, n, m, sm, tm, xs in _pop_keys(g3):
        o = CLASSES[CLASS_NAMES[cid]](parameters=(52.0, 41.0, 0.0)).populate(attr="state", seed=5 + cid, discretization=(n, m),
                                                                              spatial_modulation=sm, temporal_modulation=tm, xshift=xs)
        assert o.basis == "modes" and relerr(o.state, g3[key]) < 1e-14, key
        count += 1
    assert count == 4 * 2 * 16
    o = CLASSES["OrbitKS"](parameters=(52.0, 41.0, 0.0)).populate(attr="state", seed=9, discretization=(32, 32), tscale=3, xscale=4,
                                                                  xvar=2.5, tvar=1.5, spatial_modulation="gaussian", temporal_modulation="laplace")
    assert relerr(o.state, g3["pop/override"]) < 1e-14


def test_device_generator_modulations_match_reference(g3):
    """ks_populate_state_ex (kernel source on the CPU shim): same table; device exp/log/sqrt differ from NumPy's by an ulp."""
    for key, cid, n, m, sm, tm, xs in _pop_keys(g3):
        got = E.populate_ex(cid, n, m, [5 + cid], [[52.0, 41.0, 0.0]], sm, tm, xs)[0]
        assert relerr(got, g3[key]) < 1e-12, key
    got = E.populate_ex(0, 32, 32, [9], [[52.0, 41.0, 0.0]], "gaussian", "laplace", "sqrt", tscale=3, xscale=4, xvar=2.5, tvar=1.5)[0]
    assert relerr(got, g3["pop/override"]) < 1e-12


def test_dimension_based_discretization_table(g3):
    """ks/orbits.py:1638-1694 for every resolution keyword over a grid of (T, L), all classes, plus a pinned axis."""
    from orbithunter_b200.ks import CLASSES

    for cid, res, T, L, n, m in g3["dbd/table"]:
        got = CLASSES[CLASS_NAMES[int(cid)]].dimension_based_discretization((T, L), resolution=("default", "coarse", "fine", "power")[int(res)])
        assert tuple(int(v) for v in got) == (int(n), int(m)), (cid, res, T, L, got)
    assert tuple(CLASSES["OrbitKS"].dimension_based_discretization((44.0, 33.4), discretization=(20, None))) == tuple(int(v) for v in g3["dbd/pinned"])


@pytest.mark.parametrize("cid", range(4))
def test_resize_kernel_matches_reference(g3, cid):
    """ks_resize_modes: padding and truncation of both axes in one pass, bit for bit the reference's axis-by-axis result."""
    M = g3[f"resize/{cid}/in"]
    for (n2, m2) in ((48, 40), (16, 24), (32, 64), (24, 32), (64, 16)):
        got = E.resize_modes(cid, 32, 32, n2, m2, M[None])[0]
        assert np.array_equal(got, g3[f"resize/{cid}/{n2}x{m2}"]), (n2, m2)
    n2, m2 = (int(v) for v in g3[f"resize/{cid}/auto_shape"])
    got = E.resize_modes(cid, 32, 32, n2, m2, g3[f"resize/{cid}/auto_in"][None])[0]
    assert np.array_equal(got, g3[f"resize/{cid}/auto"])
    from orbithunter_b200.ks import CLASSES

    assert CLASSES[CLASS_NAMES[cid]].dimension_based_discretization((150.0, 70.0)) == (n2, m2)


@pytest.mark.parametrize("cid", range(4))
def test_rescale_kernel_matches_reference(g3, cid):
    """ks_rescale_field on the field of the golden input (transform by the kernel sources as well)."""
    M = g3[f"rescale/{cid}/in"]
    field = E.transform(cid, 32, 32, M[None], 2, 0)
    for method, mag in (("inf", 5.0), ("L1", 7.0), ("L2", 3.0), ("LP", 1.5)):
        got, norms = E.rescale_field(field, mag, method)
        assert relerr(got[0], g3[f"rescale/{cid}/{method}_field"]) < 1e-13, method
        back = E.transform(cid, 32, 32, got, 0, 2)[0]
        assert relerr(back, g3[f"rescale/{cid}/{method}"]) < 1e-13, method


@pytest.mark.parametrize("cid", range(4))
def test_preprocess_kernel_matches_reference(g3, cid):
    """ks_preprocess: |u| and |u_t| from the modes equal the reference's field norms; the decision is the reference's."""
    for tag in ("generic", "tiny", "steady", "nearly_steady"):
        key = f"prep/{cid}/{tag}"
        code, norms = E.preprocess(cid, 32, 32, g3[key + "/state"][None], g3[key + "/params"][None])
        ref = g3[key + "/norms"]
        # the reference takes |u_t| after a round trip through the field, which leaves round-off of size eps |u| in it
        assert np.allclose(norms[0], ref, rtol=1e-12, atol=1e-14 * ref[0]), (tag, norms, ref)
        kept = str(g3[key + "/result_class"]) == CLASS_NAMES[cid]
        assert (code[0] == 0) == kept, (tag, code)
    code, _ = E.preprocess(cid, 32, 32, g3[f"prep/{cid}/generic/state"][None], np.array([[0.0, 33.4, 0.0]]))
    assert code[0] == (1 if cid == 1 else 2)  # T == 0: equilibrium (ks/orbits.py:1574; Relative :2987)
