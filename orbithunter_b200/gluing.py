"""Gluing / tiling of initial conditions, batched on the device (behaviour of reference orbithunter/gluing.py:16-421).

A gluing is planned on the host as plain integer bookkeeping and executed on the GPU as a few batched operations:

* every tile involved in a call -- and in ``glue_batch`` / ``tile_batch`` every tile of EVERY tiling of the call -- is a
  "patch": a device field plus its class, (T, L) and current grid;
* a gluing pass along one axis asks, per strip, for new patch sizes (``_apportion``: each orbit's share of the strip's grid
  points follows its share of the strip's period / length, the reference's aspect-ratio correction, gluing.py:16-115) and
  for the strip's dimensions (``_combined_dimensions``, Orbit.glue_dimensions core.py:899-945);
* all resizes of a pass are grouped by (class, old grid, new grid) and run as ONE ``BatchedOrbitKS.resize`` per group
  (forward transform, ``ks_resize_modes``, inverse transform); identical requests (the same tile appearing many times
  in a symbol array, or in many tilings) are computed once;
* the resized fields of a strip are joined with one device concatenation; the result is a patch of the next pass.

``glue`` / ``tile`` / ``aspect_ratio_correction`` / ``rediscretize_tileset`` / ``generate_symbol_arrays`` keep the
reference's names, arguments and results for single calls.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from .batch import BatchedOrbitKS

__all__ = ["aspect_ratio_correction", "glue", "tile", "glue_batch", "tile_batch", "rediscretize_tileset", "generate_symbol_arrays",
           "pairwise_group_orbit", "expensive_pairwise_glue"]


class _Patch:
    """A field on the device with what the planner needs to know about it."""

    __slots__ = ("cls", "field", "dims", "key")

    def __init__(self, cls, field, dims, key=None):
        self.cls, self.field, self.dims = cls, field, tuple(float(d) for d in dims)
        self.key = key  # identity of the source orbit (for de-duplication); None for intermediate strips

    @property
    def shape(self):
        return tuple(self.field.shape)


def _as_patches(orbits):
    """Host orbits -> patches in the field basis; orbits given in another basis are transformed in class/grid batches."""
    out, pending = {}, {}
    for o in orbits:
        if id(o) in out:
            continue
        if o.basis == "field":
            out[id(o)] = _Patch(type(o), torch.as_tensor(np.ascontiguousarray(o.state, dtype=np.float64)).cuda(), o.dimensions(), id(o))
        else:
            out[id(o)] = None
            pending.setdefault((type(o), o.basis, tuple(o.discretization)), []).append(o)
    for (cls, basis, disc), group in pending.items():
        b = BatchedOrbitKS(np.stack([o.state for o in group]), np.array([[o.t, o.x, 0.0] for o in group]), cls._cls_name, basis, disc)
        fields = b.transform("field").state
        for i, o in enumerate(group):
            out[id(o)] = _Patch(cls, fields[i], o.dimensions(), id(o))
    return out


def _apportion(sizes, dims, floors, conserve_parity=True):
    """New grid sizes along one axis for the members of a strip (reference gluing.py:41-105).  The strip keeps its total
    number of points.  Every member starts from its class minimum (raised by one if that would flip the parity of its
    current size, when parity is conserved); the points left over are handed out in units of two (one, without parity) in
    order of increasing dimension, each member taking the integer part of its proportional share of what is still
    unassigned, and the member with the largest dimension takes the rest -- doubled, in both modes, as the reference does."""
    sizes, dims = np.asarray(sizes, dtype=np.int64), np.asarray(dims, dtype=float)
    base = np.asarray(floors, dtype=np.int64).copy()
    unit = 1
    if conserve_parity:
        base += (base % 2) != (sizes % 2)
        unit = 2
    spare = (int(sizes.sum()) - int(base.sum())) // unit
    ranking = np.argsort(dims)
    new = base.copy()
    unassigned = float(dims.sum())
    for member in ranking[:-1]:
        take = int(spare * (dims[member] / unassigned))
        new[member] += unit * take
        spare -= take
        unassigned -= dims[member]
    new[ranking[-1]] += 2 * spare
    return [int(v) for v in new]


def _combined_dimensions(patches, counts, include_zero_dimensions=True):
    """Orbit.glue_dimensions (core.py:899-945): per dimension, (number of tiles along it) x (mean over the tiles)."""
    if len(counts) < len(patches[0].dims):
        raise ValueError("Gluing shape must have as many elements as the orbit type has dimensions")
    table = np.array([p.dims for p in patches])
    out = []
    for axis in range(table.shape[1]):
        column = table[:, axis]
        if not include_zero_dimensions:
            column = column[column > 0.0]
        out.append(counts[axis] * column.mean())
    return tuple(out)


def _resize_all(requests):
    """requests: list of (patch, new_shape).  One batched device resize per (class, grid, new grid); duplicates once."""
    done, groups = {}, {}
    for patch, new_shape in requests:
        if new_shape == patch.shape:
            continue
        ident = (patch.key if patch.key is not None else id(patch), new_shape)
        if ident not in done:
            done[ident] = None
            groups.setdefault((patch.cls, patch.shape, new_shape), []).append((ident, patch))
    for (cls, shape, new_shape), members in groups.items():
        for a, (size, floor_) in enumerate(zip(new_shape, cls.minimal_shape())):
            if size < floor_:
                raise ValueError("minimum discretization requirements not met during resize.")
        batch = BatchedOrbitKS(torch.stack([p.field for _, p in members]), torch.zeros(len(members), 3, dtype=torch.float64),
                               cls._cls_name, "field", shape)
        fields = batch.resize(new_shape).state
        for i, (ident, _) in enumerate(members):
            done[ident] = fields[i]
    out = []
    for patch, new_shape in requests:
        if new_shape == patch.shape:
            out.append(patch)
        else:
            ident = (patch.key if patch.key is not None else id(patch), new_shape)
            out.append(_Patch(patch.cls, done[ident], patch.dims, None))
    return out


def _strips(grid_shape, axis):
    """Index tuples selecting every strip of a layout along ``axis`` (the other indices fixed), in row-major order."""
    others = [range(g) if i != axis else (slice(None),) for i, g in enumerate(grid_shape)]
    return list(itertools.product(*others))


def _glue_pass(layouts, axis, orbit_type, conserve_parity, include_zero_dimensions):
    """One strip-wise pass along ``axis`` over MANY layouts at once (object arrays of patches): plan every strip, run all the
    resizes of the pass together, concatenate.  Returns the collapsed layouts (extent 1 along ``axis``)."""
    plans, requests = [], []
    for li, layout in enumerate(layouts):
        for sel in _strips(layout.shape, axis):
            members = list(np.atleast_1d(layout[sel]).ravel())
            counts = tuple(len(members) if a == axis else 1 for a in range(layout.ndim))
            dims = _combined_dimensions(members, counts, include_zero_dimensions)
            along = [p.dims[axis] for p in members]
            if len(members) > 1 and sum(along) != 0.0:
                sizes = _apportion([p.shape[axis] for p in members], along, [p.cls.minimal_shape()[axis] for p in members], conserve_parity)
                targets = [tuple(s if a == axis else p.shape[a] for a in range(2)) for p, s in zip(members, sizes)]
            else:
                targets = [p.shape for p in members]
            plans.append((li, sel, dims, len(requests), len(members)))
            requests.extend(zip(members, targets))
    resized = _resize_all(requests)
    out = [np.empty(tuple(1 if a == axis else g for a, g in enumerate(layout.shape)), dtype=object) for layout in layouts]
    for li, sel, dims, start, count in plans:
        field = torch.cat([p.field for p in resized[start:start + count]], dim=axis)
        out[li][tuple(0 if isinstance(s, slice) else s for s in sel)] = _Patch(orbit_type, field, dims, None)
    return out


def _run(layouts, orbit_type, strip_wise, kwargs):
    """Glue every layout (object array of patches) of the call; returns one patch per layout."""
    conserve_parity = kwargs.get("conserve_parity", True)
    nzero = kwargs.get("include_zero_dimensions", True)
    if strip_wise:
        order = kwargs.get("gluing_order", None)
        if order is None:
            order = np.argsort(layouts[0].shape)
        finished = [None] * len(layouts)
        for axis in order:
            layouts = _glue_pass(layouts, int(axis), orbit_type, conserve_parity, nzero)
            for i, layout in enumerate(layouts):
                if layout.size == 1:
                    finished[i] = layout.ravel()[0]
        if any(f is None for f in finished):
            raise ValueError("strip-wise gluing did not produce a result; verify gluing_shape and gluing_order "
                             "are not empty or disable the strip-wise correction.")
        return finished
    out = []
    for layout in layouts:  # array-wise: every tile on the same grid, the dimensions averaged over the whole layout at once
        members = list(layout.ravel())
        dims = _combined_dimensions(members, layout.shape, nzero)
        if len({p.shape for p in members}) != 1:
            raise ValueError("array-wise gluing needs tiles of one discretization; use strip_wise=True or rediscretize_tileset")
        rows = [torch.cat([p.field for p in layout[i, :]], dim=1) for i in range(layout.shape[0])]
        out.append(_Patch(orbit_type, torch.cat(rows, dim=0), dims, None))
    return out


def _layout_of(orbit_array, patches):
    arr = np.asarray(orbit_array, dtype=object)
    if arr.ndim != 2:
        raise ValueError("Gluing shape must have as many elements as the orbit type has dimensions")
    layout = np.empty(arr.shape, dtype=object)
    for idx in np.ndindex(arr.shape):
        layout[idx] = patches[id(arr[idx])]
    return layout


def _ctor_kwargs(kwargs):
    return {k: v for k, v in kwargs.items() if k not in ("basis", "gluing_order", "conserve_parity", "include_zero_dimensions")}


def glue(orbit_array, orbit_type, strip_wise=False, **kwargs):
    """Combine an array of orbits, laid out as they are to be joined (axis 0 = time, axis 1 = space), into one ``orbit_type``
    instance (reference gluing.py:118-272): fields concatenated, dimensions = tiles-per-axis x mean tile dimension;
    ``strip_wise`` first re-proportions every strip (see module docstring)."""
    arr = np.asarray(orbit_array, dtype=object)
    patches = _as_patches(arr.ravel())
    result = _run([_layout_of(arr, patches)], orbit_type, strip_wise, kwargs)[0]
    return orbit_type(state=result.field.cpu().numpy(), basis=kwargs.get("basis", orbit_type.bases_labels()[0]), parameters=result.dims,
                      **_ctor_kwargs(kwargs))


def tile(symbol_array, tiling_dictionary, orbit_type, **kwargs):
    """``glue`` for an array of dictionary keys (reference gluing.py:275-302)."""
    symbols = np.asarray(symbol_array)
    orbits = np.empty(symbols.shape, dtype=object)
    for idx in np.ndindex(symbols.shape):
        orbits[idx] = tiling_dictionary[symbols[idx].item() if hasattr(symbols[idx], "item") else symbols[idx]]
    return glue(orbits, orbit_type, **kwargs)


def glue_batch(orbit_arrays, orbit_type, strip_wise=False, **kwargs):
    """Many gluings in one go (the sweep over symbol arrays that the reference does one ``glue`` at a time): every layout of
    ``orbit_arrays`` is glued with the same settings; tiles shared between layouts are uploaded and resized once, and every
    pass runs its resizes as one batch per (class, grid, new grid).  Returns a list of BatchedOrbitKS (field basis), one per
    distinct result grid, with ``.layout_index`` = positions of its members in ``orbit_arrays``."""
    arrays = [np.asarray(a, dtype=object) for a in orbit_arrays]
    patches = _as_patches([o for a in arrays for o in a.ravel()])
    results = _run([_layout_of(a, patches) for a in arrays], orbit_type, strip_wise, kwargs)
    by_shape = {}
    for i, r in enumerate(results):
        by_shape.setdefault(r.shape, []).append(i)
    out = []
    for shape, idx in by_shape.items():
        params = torch.tensor([list(results[i].dims) + [0.0] for i in idx], dtype=torch.float64)
        b = BatchedOrbitKS(torch.stack([results[i].field for i in idx]), params, orbit_type._cls_name, "field", shape)
        b.layout_index = idx
        out.append(b)
    return out


def tile_batch(symbol_arrays, tiling_dictionary, orbit_type, **kwargs):
    """``glue_batch`` for a list of symbol arrays over one tiling dictionary."""
    arrays = []
    for sa in symbol_arrays:
        sa = np.asarray(sa)
        arr = np.empty(sa.shape, dtype=object)
        for idx in np.ndindex(sa.shape):
            arr[idx] = tiling_dictionary[sa[idx].item() if hasattr(sa[idx], "item") else sa[idx]]
        arrays.append(arr)
    return glue_batch(arrays, orbit_type, **kwargs)


def aspect_ratio_correction(orbit_array, axis=0, conserve_parity=True):
    """Resize a strip of orbits along ``axis`` in proportion to their dimensions along it (reference gluing.py:16-115); returns
    an object array of resized orbits (the input itself for a single orbit or all-zero dimensions)."""
    strip = np.asarray(orbit_array, dtype=object)
    members = list(strip.ravel())
    along = [o.dimensions()[axis] for o in members]
    if sum(along) == 0.0 or len(members) == 1:
        return orbit_array
    sizes = _apportion([o.shapes()[0][axis] for o in members], along, [o.minimal_shape()[axis] for o in members], conserve_parity)
    patches = _as_patches(members)
    requests = [(patches[id(o)], tuple(s if a == axis else o.shapes()[0][a] for a in range(2))) for o, s in zip(members, sizes)]
    out = np.empty(len(members), dtype=object)
    for i, (o, p) in enumerate(zip(members, _resize_all(requests))):
        out[i] = o._new(state=p.field.cpu().numpy(), basis="field", discretization=p.shape).transform(to=o.basis)
    return out


def rediscretize_tileset(tiling_dictionary, new_shape=None, **kwargs):
    """Bring every tile of a dictionary to one grid (reference gluing.py:355-421): ``new_shape``, or the grid the class
    conventions give the dictionary's average dimensions (``dimension_based_discretization`` of its most common class)."""
    orbits = list(tiling_dictionary.values())
    if new_shape is None:
        mean_dims = tuple(float(np.mean(col)) for col in zip(*(o.dimensions() for o in orbits)))
        classes = [type(o) for o in orbits]
        commonest = max(set(classes), key=classes.count)
        new_shape = commonest.dimension_based_discretization(mean_dims, **kwargs)
    new_shape = tuple(int(v) for v in new_shape)
    patches = _as_patches(orbits)
    resized = _resize_all([(patches[id(o)], new_shape) for o in orbits])
    return {k: o._new(state=p.field.cpu().numpy(), basis="field", discretization=new_shape).transform(to=o.basis)
            for (k, o), p in zip(tiling_dictionary.items(), resized)}


def generate_symbol_arrays(tiling_dictionary, glue_shape, unique=True):
    """All symbol arrays of a shape over a dictionary's keys (reference gluing.py:305-352); with ``unique`` one representative
    per class of arrays equal up to cyclic translation along the axes."""
    keys, count = list(tiling_dictionary.keys()), int(np.prod(glue_shape))
    arrays = (np.reshape(combo, glue_shape) for combo in itertools.product(keys, repeat=count))
    if not unique:
        return list(arrays)
    seen, out = set(), []
    shifts = list(itertools.product(*(range(g) for g in glue_shape)))
    axes = tuple(range(len(glue_shape)))
    for arr in arrays:
        images = {tuple(np.roll(arr, s, axis=axes).ravel().tolist()) for s in shifts}
        if seen.isdisjoint(images):
            out.append(arr)
        seen |= images
    return out


# ---- pairwise search over group orbits (reference gluing.py:395-496) -----------------------------------------------------
def pairwise_group_orbit(orbit_pair, **kwargs):
    """All pairs of members of the two group orbits (reference gluing.py:395-421), first orbit's members in the outer loop."""
    pair = np.asarray(orbit_pair, dtype=object).ravel()
    yield from itertools.product(pair[0].group_orbit(**kwargs), pair[1].group_orbit(**kwargs))


def _group_fields(orbit, kwargs):
    """The fields of ``orbit.group_orbit(**kwargs)`` as one device tensor [G, n, m], in the generator's order, and the sign
    each member's spatial shift carries (reflections negate it; ks/orbits.py:1143-1170, :1355-1401)."""
    field = orbit.transform(to="field")
    F = torch.as_tensor(np.ascontiguousarray(field.state, dtype=np.float64)).cuda()
    n, m = F.shape

    def reflect(X):
        return -torch.roll(torch.flip(X, dims=(-1,)), 1, dims=-1)

    if kwargs.get("discrete", False):
        half = torch.roll(F, m // 2, dims=1)
        members, signs = [], []
        for g, sgn in ((F, 1.0), (reflect(F), -1.0), (half, 1.0), (reflect(half), -1.0)):
            members += [g, torch.roll(g, n // 2, dims=0)]
            signs += [sgn, sgn]
        return torch.stack(members), torch.tensor(signs, dtype=torch.float64, device=F.device)
    r0, r1 = kwargs.get("rolls", (1, 1))
    bases = [(F, 1.0)] if kwargs.get("continuous", False) else [(F, 1.0), (reflect(F), -1.0)]
    Ns = torch.arange(0, n, r0, device=F.device)
    Ms = torch.arange(0, m, r1, device=F.device)
    rows = (torch.arange(n, device=F.device)[None, :] - Ns[:, None]) % n  # np.roll(a, N)[i] = a[(i - N) mod n]
    cols = (torch.arange(m, device=F.device)[None, :] - Ms[:, None]) % m
    out, signs = [], []
    for g, sgn in bases:
        out.append(g[rows[:, None, :, None], cols[None, :, None, :]].reshape(-1, n, m))
        signs.append(torch.full((Ns.numel() * Ms.numel(),), sgn, dtype=torch.float64, device=F.device))
    return torch.cat(out), torch.cat(signs)


def expensive_pairwise_glue(orbit_pair, objective="cost", axis=0, **kwargs):
    """Search every pair of group-orbit members of two orbits for the best join along ``axis`` (reference gluing.py:424-496).

    The reference builds and scores one combination at a time; here the group orbits become two device tensors (rolls and
    reflections as one gather each), and the combinations are joined, transformed and scored ``chunk`` (kwarg, default 8192)
    at a time by the batched transform and eqn kernels.  ``objective``: "cost" -- 1/2 |F|^2 of the joined orbit -- or
    "boundary_cost" -- the L2 distance of the rows / columns that meet (both seams when the axis is periodic); the reference
    raises AttributeError for the latter (gluing.py:466 calls ``periodic_dimension``, which does not exist), the evident
    intent is implemented.  Ties go to the first combination in the reference's iteration order, as its strict ``<`` does.
    Group-orbit kwargs: ``rolls``, ``continuous``, ``discrete`` (ks/orbits.py:1355-1401); ``return_type``;
    ``fundamental_domain`` for the ShiftReflection and Antisymmetric classes (OrbitKS: no effect, as in the reference)."""
    pair = np.asarray(orbit_pair, dtype=object).ravel()
    a, b = pair[0], pair[1]
    return_type = kwargs.get("return_type", type(a))
    chunk = int(kwargs.get("chunk", 8192))
    GA, sa = _group_fields(a, kwargs)
    GB, _ = _group_fields(b, kwargs)
    # fundamental_domain=True (reference gluing.py:470-487 with ks/orbits.py:3229-3252, :3688-3711): the members are cut to their
    # fundamental domains, joined, and the join is completed with its reflection before it is scored
    fd = bool(kwargs.get("fundamental_domain", False))
    cls_name = type(a)._cls_name
    if fd and cls_name == "RelativeOrbitKS":
        raise NotImplementedError("expensive_pairwise_glue(fundamental_domain=True) is not provided for RelativeOrbitKS "
                                  "(its fundamental domain is the physical reference frame)")
    fd = fd and cls_name in ("ShiftReflectionOrbitKS", "AntisymmetricOrbitKS")
    if fd and objective == "boundary_cost":
        raise NotImplementedError("objective='boundary_cost' with fundamental_domain=True is not provided")
    fd_axis = 0 if cls_name == "ShiftReflectionOrbitKS" else 1
    if fd:
        if fd_axis == 0:
            GA, GB = GA[:, -(GA.shape[1] // 2):, :], GB[:, -(GB.shape[1] // 2):, :]
        else:
            GA, GB = GA[:, :, :-(GA.shape[2] // 2)], GB[:, :, :-(GB.shape[2] // 2)]
    if GA.shape[2 - axis] != GB.shape[2 - axis]:
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    ta, xa, s0 = (a.parameters + (0.0,) * 3)[:3]
    tb, xb = b.parameters[:2]
    # dimensions of the join (glue_dimensions): summed along the axis, averaged across it; cutting to the fundamental domain
    # halves the symmetric dimension of both tiles and the completion doubles it again, so the result is the same either way
    dims = tuple((2 if i == axis else 1) * 0.5 * (p + q) for i, (p, q) in enumerate(((ta, tb), (xa, xb))))
    periodic = a.periodic_dimensions()[axis]
    sr_roll = GB.shape[2] // 2 if (return_type._cls_name == "ShiftReflectionOrbitKS" and axis == 1) else 0
    nb_members = GB.shape[0]
    total = GA.shape[0] * nb_members
    best_cost, best_idx = float("inf"), -1
    for c0 in range(0, total, chunk):
        idx = torch.arange(c0, min(c0 + chunk, total), device=GA.device)
        ia, ib = idx // nb_members, idx % nb_members
        if objective == "boundary_cost":
            A, B = GA[ia], GB[ib]
            if axis == 0:
                d = (A[:, -1, :] - B[:, 0, :]).square().sum(1)
                if periodic:
                    d = d + (A[:, 0, :] - B[:, -1, :]).square().sum(1)
            else:
                d = (A[:, :, -1] - B[:, :, 0]).square().sum(1)
                if periodic:
                    d = d + (A[:, :, 0] - B[:, :, -1]).square().sum(1)
            cost = d.sqrt()
        else:
            glued = torch.cat((GA[ia], GB[ib]), dim=1 + axis)
            if sr_roll:
                glued = torch.roll(glued, sr_roll, dims=2)
            if fd:
                mirror = -torch.roll(torch.flip(glued, dims=(-1,)), 1, dims=-1)
                glued = torch.cat((mirror, glued), dim=1) if fd_axis == 0 else torch.cat((glued, mirror), dim=2)
            params = torch.stack((torch.full_like(sa[ia], dims[0]), torch.full_like(sa[ia], dims[1]), sa[ia] * s0), dim=1)
            cost = BatchedOrbitKS(glued, params, return_type._cls_name, "field").transform("modes").cost()
        cmin = float(cost.min())
        if cmin < best_cost:
            best_cost, best_idx = cmin, c0 + int((cost == cmin).nonzero()[0])
    ia, ib = best_idx // nb_members, best_idx % nb_members
    fa = a.transform(to="field")
    fb = b.transform(to="field")
    if fd:  # the winner once more through the single-orbit methods: cut, join, complete
        full_a, full_b = _group_fields(a, kwargs)[0][ia].cpu().numpy(), _group_fields(b, kwargs)[0][ib].cpu().numpy()
        ga = fa._new(state=full_a, parameters=(ta, xa, float(sa[ia]) * s0)).to_fundamental_domain()
        gb = fb._new(state=full_b).to_fundamental_domain()
        joined = ga.concat(gb, axis=axis).from_fundamental_domain()
    else:
        ga = fa._new(state=GA[ia].cpu().numpy(), parameters=(ta, xa, float(sa[ia]) * s0))
        gb = fb._new(state=GB[ib].cpu().numpy())
        joined = ga.concat(gb, axis=axis)
    return return_type(**vars(joined)).transform(to=a.basis)
