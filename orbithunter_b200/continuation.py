"""Numerical continuation in one parameter (reference orbithunter/continuation.py:10-136), on the device hunt().

``continuation`` is the single-orbit drop-in: same arguments, same loop (constrain, step the constrained parameter
towards its target, re-converge with ``hunt``, stop at the first non-converged step).  ``continuation_batch`` runs the
same loop for a whole BatchedOrbitKS in lock step, one family member per orbit slot, each with its own current value and
target: an orbit stops moving as soon as one of its hunts does not converge (its last converged member is kept), the
loop ends when every orbit has either reached its target or failed.  ``save=True`` (HDF5 output, io.py) is outside the
hot path and raises.
"""
from __future__ import annotations

import numpy as np
import torch

from .batch import PARAMETER_LABELS, BatchedOrbitKS
from .optimize import OrbitResult, hunt

__all__ = ["continuation", "continuation_batch", "discretization_continuation", "span_family"]


def _equals_target(value, target):
    """continuation.py:10-18: equality up to 13 decimals."""
    return np.round(value, 13) == np.round(target, 13)


def _next_extent(current, target, increment):
    """continuation.py:45-50: step, but never past the target."""
    if np.sign(target - (current + increment)) != np.sign(increment):
        return target
    return current + increment


def _parse_item(constraint_item):
    if isinstance(constraint_item, type({}.items())):
        return tuple(*constraint_item)
    if isinstance(constraint_item, dict):
        return tuple(*constraint_item.items())
    if isinstance(constraint_item, tuple):
        return constraint_item
    raise TypeError("constraint_item is expected to be dict, dict_item, tuple containing a single key, value pair.")


def _constrained(cls_defaults, labels):
    """Orbit.constrain (core.py:2069-2115): the given labels are held fixed, every other parameter gets its default."""
    return {key: True if key in labels else cls_defaults.get(key, True) for key in PARAMETER_LABELS}


def continuation(orbit_instance, constraint_item, *extra_constraints, step_size=0.01, **kwargs):
    """Drop-in for orbithunter.continuation.continuation on an OrbitKS-family instance (continuation.py:59-136)."""
    if kwargs.get("save", False):
        raise NotImplementedError("save=True writes HDF5 through io.py, which is outside the B200 hot path")
    label, target = _parse_item(constraint_item)
    orbit_instance.constraints = _constrained(orbit_instance._default_constraints(), (label, *extra_constraints))
    result = OrbitResult(orbit=orbit_instance, status=1)
    step = np.sign(target - getattr(result.orbit, label)) * np.abs(step_size)
    while result.status == 1 and not _equals_target(getattr(result.orbit, label), target):
        o = result.orbit
        nxt = _next_extent(getattr(o, label), target, step)
        params = tuple(nxt if lab == label else getattr(o, lab) for lab in o.parameter_labels())
        result = hunt(o._new(parameters=params), **kwargs)
    return result


def continuation_batch(batch: BatchedOrbitKS, label, targets, *extra_constraints, step_size=0.01, max_steps=100000, **kwargs):
    """Lock-step continuation of every orbit of ``batch`` in parameter ``label`` towards ``targets`` (scalar or (B,)).

    Returns OrbitResult(orbit=BatchedOrbitKS of the last converged member of each family, status=(B,) int32 with 1 where
    the target was reached and otherwise the status of the failing hunt, steps=(B,) accepted continuation steps,
    reached=(B,) bool).  The starting orbits are assumed converged (the reference assumes the same, continuation.py:101).
    """
    if kwargs.get("save", False):
        raise NotImplementedError("save=True writes HDF5 through io.py, which is outside the B200 hot path")
    if label not in PARAMETER_LABELS:
        raise ValueError(f"unknown parameter label {label!r}")
    col = PARAMETER_LABELS.index(label)
    cur = batch.transform("modes").copy()
    from .batch import default_constraints

    cur.constraints = _constrained(default_constraints(cur.cls_id), (label, *extra_constraints))
    B, dev = cur.batch, cur.state.device
    target = torch.as_tensor(np.broadcast_to(np.asarray(targets, dtype=np.float64), (B,)).copy(), device=dev)
    value = cur.parameters[:, col]
    sign = torch.sign(target - value)
    step = sign * abs(float(step_size))
    status = torch.ones(B, dtype=torch.int32, device=dev)
    steps = torch.zeros(B, dtype=torch.int32, device=dev)

    def at_target(v):
        return torch.round(v, decimals=13) == torch.round(target, decimals=13)

    for _ in range(int(max_steps)):
        value = cur.parameters[:, col]
        moving = (status == 1) & ~at_target(value)
        if not bool(moving.any()):
            break
        stepped = value + step
        overshoot = torch.sign(target - stepped) != sign
        nxt = torch.where(overshoot, target, stepped)
        trial = cur.copy()
        trial.parameters[:, col] = torch.where(moving, nxt, value)
        res = hunt(trial, **{k: (v.copy() if hasattr(v, "copy") else v) for k, v in kwargs.items()})
        ok = moving & (res.status == 1)
        failed = moving & (res.status != 1)
        keep = ok[:, None, None]
        cur.state = torch.where(keep, res.orbit.state, cur.state)
        cur.parameters = torch.where(ok[:, None], res.orbit.parameters, cur.parameters)
        status = torch.where(failed, res.status.to(torch.int32), status)
        steps = steps + ok.to(torch.int32)
    reached = at_target(cur.parameters[:, col])
    return OrbitResult(orbit=cur, status=status, steps=steps, reached=reached)


def _increment_discretization(orbit_instance, target_size, increment, axis=0):
    """continuation.py:139-167: one step of the discretization along ``axis``, never past the target."""
    current = orbit_instance.shapes()[0][axis]
    nxt = target_size if np.sign(target_size - (current + increment)) != np.sign(increment) else current + increment
    shape = tuple(int(d) if i != axis else int(nxt) for i, d in enumerate(orbit_instance.shapes()[0]))
    return orbit_instance.resize(*shape)


def discretization_continuation(orbit_instance, target_discretization, cycle=False, **kwargs):
    """Drop-in for orbithunter.continuation.discretization_continuation (continuation.py:170-279): change the grid of a
    converged orbit step by step (zero padding / truncation of the modes), re-converging with ``hunt`` after every step."""
    if kwargs.get("save", False):
        raise NotImplementedError("save=True writes HDF5 through io.py, which is outside the B200 hot path")
    target_discretization = tuple(int(x) for x in target_discretization)
    result = OrbitResult(orbit=orbit_instance, status=1)
    axes_order = np.array(kwargs.get("axes_order", np.argsort(target_discretization)[::-1]))
    step_sizes = kwargs.get("step_sizes", np.array(orbit_instance.minimal_shape_increments())[axes_order])
    hunt_kwargs = {k: v for k, v in kwargs.items() if k not in ("axes_order", "step_sizes")}

    def fresh():
        return {k: (v.copy() if hasattr(v, "copy") else v) for k, v in hunt_kwargs.items()}

    if cycle:
        idx = 0
        while result.status == 1 and tuple(result.orbit.shapes()[0]) != target_discretization:
            ax = axes_order[idx]
            step = np.sign(target_discretization[ax] - result.orbit.shapes()[0][ax]) * np.abs(step_sizes[ax])
            result = hunt(_increment_discretization(result.orbit, target_discretization[ax], step, axis=ax), **fresh())
            idx = np.mod(idx + 1, len(axes_order))
    else:
        for ax in axes_order:
            step = np.abs(step_sizes[ax]) * np.sign(target_discretization[ax] - result.orbit.shapes()[0][ax])
            while result.status == 1 and not result.orbit.shapes()[0][ax] == target_discretization[ax]:
                result = hunt(_increment_discretization(result.orbit, target_discretization[ax], step, axis=ax), **fresh())
    return result


def span_family(orbit_instance, **kwargs):
    """Drop-in for orbithunter.continuation.span_family (continuation.py:280-384): from a converged root orbit, walk every
    dimension up and down in steps of ``step_sizes[dim]`` (default 0.01), re-converging each step, until a step fails or the
    dimension leaves ``bounds[dim]``; returns [[root], branch of the first dimension, ...] with each branch ordered by
    increasing dimension (root excluded), and stores every ``sampling_rate``-th member of every branch in the HDF5 file
    ``filename`` (default 'family_' + root.filename()).

    All walkers (two per dimension) advance in LOCK STEP: one batched ``hunt`` per step re-converges the next member of every
    branch that is still alive, instead of the reference's one-orbit-at-a-time walk; a walker stops exactly where the
    reference's loop does.  The members are written in one pass (``save=False`` skips the file).  Keyword arguments of
    ``hunt`` pass through, as in the reference."""
    from .h5io import _free_name, _insert, _open_tree, _orbit_attrs, write_tree

    step_sizes = kwargs.get("step_sizes", {})
    bounds = kwargs.get("bounds", {k: (0, np.inf) for k in orbit_instance.dimension_labels()})
    filename = kwargs.get("filename", "".join(["family_", orbit_instance.filename()]))
    skip = ("step_sizes", "bounds", "filename", "strides", "sampling_rate", "leafnames", "save", "step_size", "h5mode")
    hunt_kwargs = {k: v for k, v in kwargs.items() if k not in skip}
    root = orbit_instance.transform(to="modes")
    labels = root.parameter_labels()
    dims = [d for d in bounds if getattr(root, d) != 0.0]
    # one walker per (dimension, direction): its current member, whether it still moves, the members it has collected
    walkers = []
    for d in dims:
        for sign in (+1.0, -1.0):
            walkers.append({"dim": d, "sign": sign, "step": abs(step_sizes.get(d, 0.01)), "orbit": root, "alive": True, "members": []})

    def in_bounds(w):
        v = getattr(w["orbit"], w["dim"])
        return v < bounds.get(w["dim"])[1] if w["sign"] > 0 else v > bounds.get(w["dim"])[0]

    while True:
        moving = []
        for w in walkers:
            if not w["alive"]:
                continue
            if not in_bounds(w):  # the reference's while-condition: converged AND strictly inside the bounds
                w["alive"] = False
                continue
            if getattr(w["orbit"], w["dim"]) != getattr(root, w["dim"]):
                w["members"].append(w["orbit"])
            moving.append(w)
        if not moving:
            break
        # next member of every live branch: constrain its dimension, step it, re-converge -- batched per constraint set
        groups = {}
        for w in moving:
            groups.setdefault(w["dim"], []).append(w)
        for d, ws in groups.items():
            col = labels.index(d)
            states = np.stack([w["orbit"].state for w in ws])
            params = np.array([[p + (w["sign"] * w["step"] if i == col else 0.0) for i, p in enumerate(w["orbit"].parameters)] for w in ws])
            trial = BatchedOrbitKS(states, params, root._cls_name, "modes", root.discretization,
                                   _constrained(root._default_constraints(), (d,)))
            res = hunt(trial, **{k: (v.copy() if hasattr(v, "copy") else v) for k, v in hunt_kwargs.items()})
            status = res.status.cpu().numpy()
            for i, w in enumerate(ws):
                if status[i] == 1:
                    w["orbit"] = res.orbit[i]
                else:
                    w["alive"] = False
    family = [[orbit_instance]]
    tree = None
    for d in dims:
        up = [w for w in walkers if w["dim"] == d and w["sign"] > 0][0]["members"]
        down = [w for w in walkers if w["dim"] == d and w["sign"] < 0][0]["members"]
        branch = list(reversed(down)) + up
        family.append(branch)
        if kwargs.get("save", True) and branch:
            if tree is None:
                tree = _open_tree(filename, kwargs.get("h5mode", "a"))
            for leaf in branch[:: int(kwargs.get("sampling_rate", 1))]:
                name = leaf.filename(cls_name=False, extension="").lstrip("_") if kwargs.get("leafnames", False) else ""
                _insert(tree, _free_name(tree, "", name), (leaf.state, _orbit_attrs(leaf, leaf.cost())))
    if tree is not None:
        write_tree(filename, tree)
    return family
