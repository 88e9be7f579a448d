"""Gluing / tiling initial conditions from smaller orbits (reference orbithunter/gluing.py:16-302).

``glue`` combines an array of orbits into one larger orbit by concatenating their fields (time along axis 0, space along
axis 1) and averaging their dimensions; with ``strip_wise=True`` every strip is first re-proportioned by
``aspect_ratio_correction`` (each orbit gets a share of the strip's grid points proportional to its period / length).
``tile`` is the same with a symbol array and a dictionary of orbits.  The array bookkeeping is host-side NumPy as in the
reference; the basis changes and rediscretisations run through the orbit classes (device transforms).  This is what builds
the large-domain inputs (BASELINE config 4) that ``hunt`` then converges.
"""
from __future__ import annotations

import itertools

import numpy as np

__all__ = ["aspect_ratio_correction", "glue", "tile", "rediscretize_tileset"]


def glue_dimensions(dimension_tuples, glue_shape, include_zero_dimensions=True):
    """Orbit.glue_dimensions (core.py:899-945): per dimension, glue_shape[i] times the mean over the orbits."""
    try:
        arrays = [np.array(p) for p in dimension_tuples]
        if include_zero_dimensions:
            return tuple(glue_shape[i] * p.mean() for i, p in enumerate(arrays))
        return tuple(glue_shape[i] * p[p > 0.0].mean() for i, p in enumerate(arrays))
    except IndexError as ie:
        raise ValueError("Gluing shape must have as many elements as the orbit type has dimensions") from ie


def _object_array(orbits, shape=None):
    arr = np.empty(len(orbits), dtype=object)
    for i, o in enumerate(orbits):
        arr[i] = o
    return arr if shape is None else arr.reshape(shape)


def aspect_ratio_correction(orbit_array, axis=0, conserve_parity=True):
    """gluing.py:16-115: resize a strip of orbits along ``axis`` in proportion to their dimensions along that axis."""
    dims = np.array([o.dimensions()[axis] for o in orbit_array])
    sizes = np.array([o.shapes()[0][axis] for o in orbit_array])
    disc_total, dim_total = np.sum(sizes), np.sum(dims)
    if dim_total == 0.0 or len(orbit_array.ravel()) == 1:
        return orbit_array
    new_sizes = np.zeros(len(orbit_array))
    order = np.argsort(dims)
    if conserve_parity:
        min_sizes = np.array([ms + 1 if (ms % 2) != (sizes[i] % 2) else ms
                              for i, ms in enumerate(o.minimal_shape()[axis] for o in orbit_array)])
        disc_remainder = (disc_total - np.sum(min_sizes)) // 2
        dim_remainder = dim_total
        for idx in order[:-1]:
            share = int(disc_remainder * (dims[idx] / dim_remainder))
            new_sizes[idx] = min_sizes[idx] + 2 * share
            dim_remainder -= dims[idx]
            disc_remainder -= share
        new_sizes[order[-1]] = min_sizes[order[-1]] + 2 * disc_remainder
    else:
        min_sizes = np.array([o.minimal_shape()[axis] for o in orbit_array])
        disc_remainder = disc_total - np.sum(min_sizes)
        dim_remainder = dim_total
        for idx in order[:-1]:
            share = int(disc_remainder * (dims[idx] / dim_remainder))
            new_sizes[idx] = min_sizes[idx] + share
            dim_remainder -= dims[idx]
            disc_remainder -= share
        new_sizes[order[-1]] = min_sizes[order[-1]] + 2 * disc_remainder  # (sic, gluing.py:103-105)
    new_shapes = [tuple(int(new_sizes[j]) if i == axis else o.shapes()[0][i] for i in range(len(o.shape)))
                  for j, o in enumerate(orbit_array)]
    return _object_array([o.resize(shp) for o, shp in zip(orbit_array, new_shapes)])


def glue(orbit_array, orbit_type, strip_wise=False, **kwargs):
    """gluing.py:118-272.  ``orbit_array``: NumPy object array of orbits (field basis) whose shape is the gluing layout."""
    glue_shape = orbit_array.shape
    tile_shape = orbit_array.ravel()[0].shapes()[0]
    gluing_order = kwargs.get("gluing_order", np.argsort(glue_shape))
    conserve_parity = kwargs.get("conserve_parity", True)
    nzero = kwargs.get("include_zero_dimensions", True)
    gluing_basis = kwargs.get("basis", orbit_type.bases_labels()[0])
    ctor_kwargs = {k: v for k, v in kwargs.items() if k != "basis"}
    if strip_wise:
        glued_orbit = None
        for gluing_axis in gluing_order:
            slices = itertools.product(*(range(g) if i != gluing_axis else [slice(None)] for i, g in enumerate(glue_shape)))
            strips = []
            for gs in slices:
                strip_shape = tuple(len(orbit_array[gs]) if n == gluing_axis else 1 for n in range(len(orbit_array.shape)))
                zipped = tuple(zip(*(o.dimensions() for o in orbit_array[gs].ravel())))
                strip_parameters = glue_dimensions(zipped, glue_shape=strip_shape, include_zero_dimensions=nzero)
                corrected = aspect_ratio_correction(orbit_array[gs].ravel(), axis=gluing_axis, conserve_parity=conserve_parity)
                state = np.concatenate(tuple(x.state for x in corrected), axis=gluing_axis)
                strips.append(orbit_type(state=state, basis=gluing_basis, parameters=strip_parameters, **ctor_kwargs))
            glue_shape = tuple(glue_shape[i] if i != gluing_axis else 1 for i in range(len(glue_shape)))
            orbit_array = _object_array(strips, glue_shape)
            if orbit_array.size == 1:
                glued_orbit = orbit_array.ravel()[0]
        if glued_orbit is None:
            raise ValueError("strip-wise gluing did not produce a result; verify gluing_shape and gluing_order "
                             "are not empty or disable the strip-wise correction.")
        return glued_orbit
    zipped = tuple(zip(*(o.dimensions() for o in orbit_array.ravel())))
    glued_dimensions = glue_dimensions(zipped, glue_shape=glue_shape, include_zero_dimensions=nzero)
    gluing_axis = len(glue_shape) - 1
    fields = np.array([o.transform(to=orbit_type.bases_labels()[0]).state for o in orbit_array.ravel()])
    state = fields.reshape(*orbit_array.shape, *tile_shape)
    while len(state.shape) > len(tile_shape):
        state = np.concatenate(state, axis=gluing_axis)
    return orbit_type(state=state, basis=gluing_basis, parameters=glued_dimensions, **ctor_kwargs)


def tile(symbol_array, tiling_dictionary, orbit_type, **kwargs):
    """gluing.py:275-302: glue the orbits that the symbols of ``symbol_array`` stand for."""
    symbol_array = np.asarray(symbol_array)
    orbits = _object_array([tiling_dictionary[s] for s in symbol_array.ravel()], symbol_array.shape)
    return glue(orbits, orbit_type, **kwargs)


def rediscretize_tileset(tiling_dictionary, new_shape=None, **kwargs):
    """gluing.py:355-421, explicit ``new_shape`` only: bring every tile to the same discretization (as non-strip-wise glue needs)."""
    if new_shape is None:
        raise NotImplementedError("rediscretize_tileset needs an explicit new_shape here (the reference's default derives it "
                                  "from dimension_based_discretization of the average dimensions)")
    return {k: o.resize(*new_shape) for k, o in tiling_dictionary.items()}
