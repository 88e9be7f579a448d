"""GPU (-m gpu): the auxiliary device operations through the product API against outputs of the unmodified reference
(tests/golden/ks_aux_r3.npz): seed generation with every modulation, the tutorial's catalog seed (cost 8807.93719632008),
rescale, resize, preprocess."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
CLASS_NAMES = ("OrbitKS", "RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS")


@pytest.fixture(scope="module")
def g3():
    return np.load(os.path.join(GOLDEN, "ks_aux_r3.npz"))


def test_from_seeds_with_every_modulation(g3):
    from orbithunter_b200 import BatchedOrbitKS

    n_checked = 0
    for key in g3.files:
        if not key.startswith("pop/") or key == "pop/override":
            continue
        _, tag, sm, tm, xs = key.split("/")
        cid, shape = tag.split("_")
        cid, (n, m) = int(cid), (int(v) for v in shape.split("x"))
        b = BatchedOrbitKS.from_seeds([5 + cid, 5 + cid], (52.0, 41.0, 0.0), CLASS_NAMES[cid], (n, m), spatial_modulation=sm,
                                      temporal_modulation=tm, xshift=xs)
        got = b.cpu_state()
        assert relerr(got[0], g3[key]) < 1e-12 and np.array_equal(got[0], got[1]), key
        n_checked += 1
    assert n_checked == 128
    b = BatchedOrbitKS.from_seeds([9], (52.0, 41.0, 0.0), "OrbitKS", (32, 32), tscale=3, xscale=4, xvar=2.5, tvar=1.5,
                                  spatial_modulation="gaussian", temporal_modulation="laplace")
    assert relerr(b.cpu_state()[0], g3["pop/override"]) < 1e-12


def test_catalog_seed_of_the_published_timing(g3):
    """notebooks/tutorial.ipynb: OrbitKS().populate(seed=0, attr='all', discretization=(64, 64), spatial_modulation='gaussian',
    temporal_modulation='truncate').rescale(5) -- the seed of numerical_method_benchmarks.csv (cost 8807.93719632008)."""
    from orbithunter_b200 import BatchedOrbitKS, OrbitKS

    np.random.seed(0)
    o = OrbitKS().populate(seed=0, attr="all", discretization=(64, 64), spatial_modulation="gaussian",
                           temporal_modulation="truncate").rescale(5)
    assert o.basis == "modes" and np.allclose(o.parameters, g3["catalog/params"], rtol=1e-15)
    assert relerr(o.state, g3["catalog/state"]) < 1e-13
    assert abs(o.cost() - 8807.93719632008) < 1e-9 * 8807.93719632008 and abs(o.cost() - float(g3["catalog/cost"])) < 1e-8
    # the same on the device generator
    b = BatchedOrbitKS.from_seeds([0], tuple(g3["catalog/params"]), "OrbitKS", (64, 64), spatial_modulation="gaussian",
                                  temporal_modulation="truncate").rescale(5.0)
    assert relerr(b.cpu_state()[0], g3["catalog/state"]) < 1e-12
    assert abs(float(b.cost()[0]) - 8807.93719632008) < 1e-8 * 8807.93719632008


@pytest.mark.parametrize("cid", range(4))
def test_rescale_resize_preprocess_on_device(g3, cid):
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.ks import CLASSES

    C = CLASSES[CLASS_NAMES[cid]]
    M = g3[f"rescale/{cid}/in"]
    o = C(state=M, basis="modes", parameters=(44.0, 33.4, 0.0))
    b = BatchedOrbitKS(np.stack([M, 0.5 * M]), np.tile([44.0, 33.4, 0.0], (2, 1)), CLASS_NAMES[cid])
    for method, mag in (("inf", 5.0), ("L1", 7.0), ("L2", 3.0), ("LP", 1.5)):
        ref = g3[f"rescale/{cid}/{method}"]
        assert relerr(o.rescale(mag, method=method).state, ref) < 1e-13
        got = b.rescale(mag, method).cpu_state()
        assert relerr(got[0], ref) < 1e-13
        if method != "LP":
            assert relerr(got[1], ref) < 1e-13  # norm-based rescaling forgets the input's amplitude
        f = o.transform(to="field").rescale(mag, method=method)
        assert f.basis == "field" and relerr(f.state, g3[f"rescale/{cid}/{method}_field"]) < 1e-13
    with pytest.raises(ValueError):
        o.rescale(1.0, method="L7")
    # resize: batched device kernel and the single-orbit class
    M = g3[f"resize/{cid}/in"]
    o = C(state=M, basis="modes", parameters=(44.0, 33.4, 0.0))
    b = BatchedOrbitKS(M[None], [[44.0, 33.4, 0.0]], CLASS_NAMES[cid])
    for (n2, m2) in ((48, 40), (16, 24), (32, 64), (24, 32), (64, 16)):
        ref = g3[f"resize/{cid}/{n2}x{m2}"]
        r = b.resize(n2, m2)
        assert r.discretization == (n2, m2) and np.array_equal(r.cpu_state()[0], ref)
        assert relerr(o.resize(n2, m2).state, ref) < 1e-15
        rf = b.transform("field").resize((n2, m2))
        assert rf.basis == "field" and relerr(rf.transform("modes").cpu_state()[0], ref) < 1e-13
    big = C(state=g3[f"resize/{cid}/auto_in"], basis="modes", parameters=(150.0, 70.0, 0.0))
    r = big.resize()
    assert tuple(r.discretization) == tuple(int(v) for v in g3[f"resize/{cid}/auto_shape"]) and relerr(r.state, g3[f"resize/{cid}/auto"]) < 1e-15
    # preprocess
    for tag in ("generic", "tiny", "steady", "nearly_steady"):
        key = f"prep/{cid}/{tag}"
        o = C(state=g3[key + "/state"], basis="modes", parameters=tuple(g3[key + "/params"]))
        res = o.preprocess()
        kept = str(g3[key + "/result_class"]) == CLASS_NAMES[cid]
        if kept:
            assert not getattr(res, "equilibrium", False) and np.allclose(res.parameters, g3[key + "/result_params"])
            assert np.array_equal(res.state, o.state)
        elif tag == "tiny":
            assert not res.state.any()
        else:
            assert res.equilibrium


def test_batched_gluing_matches_reference():
    """glue_batch / tile_batch: all golden layouts of tests/golden/ks_gluing.npz (outputs of the unmodified reference's glue and
    tile) glued in ONE call -- shared tiles uploaded and resized once, one batched resize per (class, grid, new grid) and pass."""
    from orbithunter_b200 import OrbitKS
    from orbithunter_b200.gluing import aspect_ratio_correction, glue_batch, rediscretize_tileset, tile_batch

    g = np.load(os.path.join(GOLDEN, "ks_gluing.npz"))
    o = {t: OrbitKS(state=g[f"in/{t}/field"], basis="field", parameters=tuple(g[f"in/{t}/params"])) for t in "abc"}
    a, b, c = o["a"], o["b"], o["c"]

    def arr(rows):
        x = np.empty((len(rows), len(rows[0])), dtype=object)
        for i, r in enumerate(rows):
            for j, v in enumerate(r):
                x[i, j] = v
        return x

    strip_tags = ["time_strip", "mixed_strip"]
    out = glue_batch([arr([[a], [b]]), arr([[a], [c]])], OrbitKS, strip_wise=True)
    seen = 0
    for batch in out:
        fields, params = batch.cpu_state(), batch.cpu_parameters()
        for row, li in enumerate(batch.layout_index):
            tag = strip_tags[li]
            assert batch.discretization == tuple(g[tag + "/shape"]) and np.allclose(params[row, :2], g[tag + "/params"][:2], rtol=1e-14)
            assert relerr(fields[row], g[tag + "/field"]) < 1e-13, tag
            seen += 1
    assert seen == 2
    plain = glue_batch([arr([[a], [b]]), arr([[b], [a]])], OrbitKS)
    assert len(plain) == 1 and plain[0].batch == 2 and relerr(plain[0].cpu_state()[0], g["time/field"]) < 1e-13
    assert np.array_equal(plain[0].cpu_state()[1], np.concatenate((b.state, a.state), axis=0))
    tb = tile_batch([np.array([[0, 1], [1, 1]]), np.array([[0, 1], [1, 0]])], {0: a, 1: b}, OrbitKS, strip_wise=True)
    got = {li: (bt, row) for bt in tb for row, li in enumerate(bt.layout_index)}
    bt, row = got[0]
    assert relerr(bt.cpu_state()[row], g["tile/field"]) < 1e-13 and np.allclose(bt.cpu_parameters()[row, :2], g["tile/params"][:2], rtol=1e-14)
    bt, row = got[1]
    assert relerr(bt.cpu_state()[row], g["grid_strip/field"]) < 1e-13
    # the glued batch is a usable initial condition on the device path
    assert float(bt.transform("modes").cost()[row]) > 0.0
    # strip correction on its own, and the dictionary-wide default grid
    fixed = aspect_ratio_correction(arr([[a], [c]])[:, 0], axis=0)
    assert sum(f.discretization[0] for f in fixed) == 48 and all(f.basis == "field" for f in fixed)
    same = rediscretize_tileset({0: a, 1: c})
    assert {tuple(v.discretization) for v in same.values()} == {OrbitKS.dimension_based_discretization((35.0, 26.0))}
    assert relerr(same[1].state, c.resize(same[1].discretization).state) < 1e-13


def test_span_family_and_hdf5_round_trip(tmp_path):
    """span_family (continuation.py:280-384) against the unmodified reference's family of the polished rpo
    (tests/golden/ks_family.npz): same branches, same members, same saved names; the file it writes is read back by read_h5;
    to_h5 / to_h5_batch / read_h5 round trips with the reference's attribute set."""
    from orbithunter_b200 import BatchedOrbitKS, RelativeOrbitKS
    from orbithunter_b200.continuation import span_family
    from orbithunter_b200.h5io import h5_tree, read_h5, to_h5_batch

    g = np.load(os.path.join(GOLDEN, "ks_family.npz"))
    gc = np.load(os.path.join(GOLDEN, "ks_continuation.npz"))
    root = RelativeOrbitKS(state=gc["start/modes"], basis="modes", parameters=tuple(gc["start/params"]), discretization=(32, 32))
    assert root.filename() == str(g["root_filename"])
    fname = str(tmp_path / "family.h5")
    fam = span_family(root, bounds={"t": tuple(g["bounds_t"]), "x": tuple(g["bounds_x"])}, step_sizes={"t": 0.01, "x": 0.01}, leafnames=True,
                      filename=fname, methods="gmres", maxiter=40, tol=1e-6, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})
    assert len(fam) == int(g["n_branches"]) and fam[0][0] is root
    names = []
    for b in range(1, len(fam)):
        assert len(fam[b]) == int(g[f"branch{b}/size"]), b
        for i, o in enumerate(fam[b]):
            assert np.allclose(o.parameters, g[f"branch{b}/{i}/params"], rtol=1e-6)
            assert relerr(o.state, g[f"branch{b}/{i}/modes"]) < 1e-5
            assert o.filename(cls_name=False, extension="").lstrip("_") == str(g[f"branch{b}/{i}/leafname"])
            names.append(str(g[f"branch{b}/{i}/leafname"]))
    assert names == [str(s) for s in g["saved_names"]]
    stored = h5_tree(fname)
    assert sorted(stored) == sorted(names)
    back = read_h5(fname, names[0])
    assert isinstance(back, RelativeOrbitKS) and back.basis == "modes" and tuple(back.discretization) == (32, 32) and back.frame == "comoving"
    assert np.array_equal(back.state, fam[1][0].state) and np.allclose(back.parameters, fam[1][0].parameters, rtol=0, atol=0)
    assert stored[names[0]][1]["cost"] < 1e-6
    # single-orbit to_h5 with the reference's collision rule, and the batch writer
    f2 = str(tmp_path / "orbits.h5")
    assert root.to_h5(filename=f2) == "0" and root.to_h5(filename=f2) == "1" and root.to_h5(filename=f2, dataname="rpo") == "rpo"
    assert root.to_h5(filename=f2, dataname="rpo") == "rpo_0"
    b = BatchedOrbitKS.from_seeds(np.arange(6), (44.0, 33.4, 0.0), "OrbitKS", (32, 32))
    written = to_h5_batch(b, f2, groupname="sweep", select=[1, 3, 4])
    assert written == ["sweep/0", "sweep/1", "sweep/2"]
    group = read_h5(f2, "sweep")
    assert len(group) == 3 and np.array_equal(group[1].state, b.cpu_state()[3])
    costs = b.cost().cpu().numpy()
    assert abs(h5_tree(f2)["sweep"]["2"][1]["cost"] - costs[4]) <= 1e-12 * costs[4]
    everything = read_h5(f2)
    assert len(everything) == 5  # '0', '1', 'rpo', 'rpo_0' and the 'sweep' group
    assert read_h5(f2, "rpo", validate=True).parameters[:2] == root.parameters[:2]  # preprocess keeps a genuine orbit


@pytest.mark.parametrize("cid,shape,B", [(0, (256, 256), 5), (1, (128, 128), 9), (2, (64, 128), 7), (3, (128, 64), 7)])
def test_tma_staged_tile_kernels_match_the_plain_ones(cid, shape, B):
    """ks_tile_tma.cuh (cp.async.bulk.tensor boxes / cp.async.bulk lines, persistent CTAs, mbarrier ring) against the plain pass
    kernels: the per-tile arithmetic is the same, so every op must agree bit for bit -- and with the oracle to 1e-12."""
    from oracle import ks_oracle as ko
    from orbithunter_b200 import BatchedOrbitKS, _lib

    n, m = shape
    L = _lib.lib()
    u = BatchedOrbitKS.from_seeds(np.arange(B), (60.0, 44.0, 0.4 if cid == 1 else 0.0), CLASS_NAMES[cid], (n, m))
    v = BatchedOrbitKS.from_seeds(np.arange(B) + 50, (60.0, 44.0, 0.0), CLASS_NAMES[cid], (n, m))
    v.parameters[:, :] = 0.25
    out = {}
    try:
        L.ks_debug_tile(2, 256)
        for mode in (0, 1, 2, 3):
            L.ks_debug_tile_tma(mode)
            cost, grad = u.costgrad()
            out[mode] = [t.cpu().numpy() for t in (cost, grad.state, grad.parameters, u.eqn().state, u.matvec(v).state, u.rmatvec(v).state,
                                                    u.rmatvec(v).parameters, u.costgrad(preconditioning=True)[1].state)]
    finally:
        L.ks_debug_tile_tma(0)
        L.ks_debug_tile(1, 256)
    for mode in (1, 2, 3):
        for a, b in zip(out[0], out[mode]):
            assert np.array_equal(a, b), mode
    M, p = u.cpu_state(), u.cpu_parameters()
    oc, ogs, ogp, _ = ko.costgrad(M, p, cid)
    tol = max(1e-12, 64 * 2.2e-16 * (2 * np.pi * (m // 2 - 1) / 44.0) ** 4)  # the adjoint amplifies FFT round-off by q_max^4
    assert relerr(out[1][1], ogs) < tol and relerr(out[1][0], oc) < 1e-12


@pytest.mark.parametrize("cid,shape", [(0, (32, 32)), (1, (64, 64)), (2, (16, 24)), (3, (128, 64))])
def test_empty_and_single_orbit_batches(cid, shape):
    """Edge cases of the batch axis on every kernel family (register-resident 32x32, CTA-resident 64x64, generic, multi-pass
    tiles): an EMPTY batch goes through every entry point without a launch error and returns empty results; a batch of ONE
    orbit gives the same numbers as that orbit inside a larger batch (no dependence on the neighbours or on the grid size)."""
    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    n, m = shape
    name = CLASS_NAMES[cid]
    big = BatchedOrbitKS.from_seeds(np.arange(5), (60.0, 44.0, 0.2 if cid == 1 else 0.0), name, (n, m))
    one = big._like(big.state[2:3].contiguous(), big.parameters[2:3].contiguous())
    empty = big._like(big.state[:0].contiguous(), big.parameters[:0].contiguous())
    # empty batch
    assert empty.batch == 0 and BatchedOrbitKS.from_seeds([], (60.0, 44.0, 0.0), name, (n, m)).batch == 0
    c, g = empty.costgrad()
    assert c.shape == (0,) and g.state.shape == (0,) + tuple(big.state.shape[1:]) and g.parameters.shape == (0, 3)
    assert empty.eqn().state.shape[0] == 0 and empty.matvec(empty).state.shape[0] == 0 and empty.rmatvec(empty).state.shape[0] == 0
    assert empty.transform("field").state.shape == (0, n, m) and empty.resize(n + 2, m + 2).state.shape[0] == 0
    assert empty.rescale(2.0).state.shape[0] == 0 and empty.preprocess_codes()[0].shape == (0,)
    for method in ("adj", "gmres"):
        r = hunt(empty, method, maxiter=3)
        assert r.status.shape == (0,) and r.orbit.batch == 0
    # a batch of one
    cb, gb = big.costgrad()
    c1, g1 = one.costgrad()
    assert torch.equal(c1[0], cb[2]) and torch.equal(g1.state[0], gb.state[2]) and torch.equal(g1.parameters[0], gb.parameters[2])
    assert torch.equal(one.eqn().state[0], big.eqn().state[2])
    assert torch.equal(one.matvec(one).state[0], big.matvec(big).state[2])
    rb, r1 = hunt(big, "adj", maxiter=25), hunt(one, "adj", maxiter=25)
    assert int(r1.status[0]) == int(rb.status[2]) and int(r1.nit[0]) == int(rb.nit[2]) and torch.equal(r1.orbit.state[0], rb.orbit.state[2])


@pytest.mark.gpu
@pytest.mark.parametrize("cid", range(4))
def test_warp32x1_matvec_rmatvec_vs_oracle_and_variant3(cid):
    """matvec / rmatvec on the one-orbit-per-warp kernel (variant 4, the default) against the oracle and against the
    two-orbits-per-warp kernel (variant 3), on a batch spanning several rounds of the persistent warps."""
    from orbithunter_b200 import _lib
    from oracle import ks_oracle as ko
    from test_gpu_parity import _b

    B = 2 * 148 * 8 + 53
    rng = np.random.RandomState(10 + cid)
    M = np.stack([ko.populate_state(44.0, 33.4, 32, 32, s % 61, cid) for s in range(B)]) * (1.0 + 1e-3 * np.arange(B)[:, None, None])
    p = np.stack([44.0 + rng.rand(B), 33.4 + rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    V = rng.randn(*M.shape) * (M != 0)
    vp = rng.randn(B, 3)
    L = _lib.lib()
    out = {}
    try:
        for cons in (ko.default_constraints(cid), (True, False, True)):
            for variant in (4, 3):
                L.ks_debug_warp32_variant(variant)
                u, vb = _b(cid, M, p, cons), _b(cid, V, vp, cons)
                mv = u.matvec(vb).state.cpu().numpy()
                r = u.rmatvec(vb)
                out[variant] = (mv, r.state.cpu().numpy(), r.parameters.cpu().numpy())
            omv = ko.matvec(M, p, V, vp, cid, cons)
            ors, orp = ko.rmatvec(M, p, V, cid, cons)
            for variant in (4, 3):
                mv, rs, rp = out[variant]
                assert np.max(np.abs(mv - omv)) < 1e-12 * np.max(np.abs(omv))
                assert np.max(np.abs(rs - ors)) < 1e-12 * np.max(np.abs(ors))
                assert np.allclose(rp, orp, rtol=1e-10, atol=1e-10 * np.max(np.abs(orp)) + 1e-300)
            assert np.max(np.abs(out[4][0] - out[3][0])) < 1e-13 * np.max(np.abs(omv))
    finally:
        L.ks_debug_warp32_variant(4)


@pytest.mark.gpu
@pytest.mark.parametrize("cid,shape", [(0, (32, 32)), (1, (32, 32)), (2, (32, 32)), (3, (32, 32)), (1, (16, 16)), (0, (64, 64))])
def test_batched_jacobian_equals_matvec_on_unit_vectors(cid, shape):
    """BatchedOrbitKS.jacobian (ks_jacobian; OrbitKS.jacobian ks/orbits.py:557-657): 32x32 uses the Jacobian mode of the
    one-orbit-per-warp kernel, other sizes a stored identity.  J @ v must equal matvec(v) for random v (oracle), and a sample
    of columns must equal the oracle's matvec on unit vectors."""
    from oracle import ks_oracle as ko
    from test_gpu_parity import _b

    n, m = shape
    B = 3 if n == 32 else 1
    rng = np.random.RandomState(cid)
    M = np.stack([ko.populate_state(44.0, 44.0, n, m, 5 + s, cid) for s in range(B)])
    p = np.stack([44.0 + rng.rand(B), 44.0 + rng.rand(B), (rng.randn(B) if cid == 1 else np.zeros(B))], axis=1)
    for cons in (ko.default_constraints(cid), (False, True, True)):
        u = _b(cid, M, p, cons)
        J = u.jacobian().cpu().numpy()
        nm = M.shape[1] * M.shape[2]
        free = [i for i in range(3) if not cons[i] and (i < 2 or cid == 1)]
        assert J.shape == (B, nm, nm + len(free))
        v = rng.randn(B, nm + len(free))
        V = v[:, :nm].reshape(M.shape)
        vp = np.zeros((B, 3))
        for k, i in enumerate(free):
            vp[:, i] = v[:, nm + k]
        ref = ko.matvec(M, p, V, vp, cid, cons).reshape(B, nm)
        got = np.einsum("bij,bj->bi", J, v)
        assert np.max(np.abs(got - ref)) < 1e-11 * np.max(np.abs(ref))
        for col in list(range(0, nm, max(1, nm // 7))) + list(range(nm, nm + len(free))):
            e = np.zeros((B, nm + len(free)))
            e[:, col] = 1.0
            ep = np.zeros((B, 3))
            for k, i in enumerate(free):
                ep[:, i] = e[:, nm + k]
            refc = ko.matvec(M, p, e[:, :nm].reshape(M.shape), ep, cid, cons).reshape(B, nm)
            assert np.max(np.abs(J[:, :, col] - refc)) < 1e-12 * max(np.max(np.abs(J)), 1.0), col


@pytest.mark.gpu
def test_dense_newton_jacobian_mode_changes_no_result():
    """hunt('lstsq') at 32x32 by the three routes of ks_newton_dense: matrix-free J J^T (2, the default: rmatvec over unit
    vectors, matvec over the result, no Jacobian stored), Jacobian by the kernel's Jacobian mode + GEMM (1), Jacobian by a
    batched matvec over a stored identity + GEMM (0).  Routes 1 and 0 produce the same bits; route 2 sums in another order,
    and Newton far from an orbit amplifies that, so: at least 85 % of the exit codes agree and the sets of converged seeds
    overlap by 80 % or more (the bar of the comparison with the reference, test_cfg0_status_table_matches_reference)."""
    from orbithunter_b200 import BatchedOrbitKS, _lib
    from orbithunter_b200.optimize import hunt

    L = _lib.lib()
    u = BatchedOrbitKS.from_seeds(np.arange(96), (44.0, 44.0, 0.0), "OrbitKS", (32, 32))
    w = hunt(u, "adj", maxiter=10000).orbit
    res = {}
    try:
        for mode in (2, 1, 0):
            L.ks_debug_dense_jacobian_mode(mode)
            res[mode] = hunt(w, "lstsq", maxiter=200, tol=1e-10)
    finally:
        L.ks_debug_dense_jacobian_mode(2)
    assert torch.equal(res[1].status, res[0].status) and torch.equal(res[1].orbit.state, res[0].orbit.state)
    a, b = res[2], res[1]
    assert (a.status == b.status).float().mean().item() >= 0.85
    ca, cb = (a.status == 1).cpu().numpy(), (b.status == 1).cpu().numpy()
    assert ca.sum() > 0 and (ca & cb).sum() >= 0.8 * max(ca.sum(), cb.sum())
    assert np.all(a.costs[-1].cpu().numpy()[ca] <= 1e-10) and np.all(b.costs[-1].cpu().numpy()[cb] <= 1e-10)
