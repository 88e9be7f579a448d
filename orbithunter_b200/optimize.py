"""hunt(): the reference's solver front end (orbithunter/optimize.py:82-364) for single orbits and whole batches.

Methods on the device path: 'adj' (adjoint descent, optimize.py:367-509), 'gmres', 'lsqr' and 'lsmr' (Newton-Krylov,
optimize.py:1196-1404 with _sparse_linalg_factory :1682-1850), 'newton_descent' (matrix-free restatement of
:653-852: the pseudo-inverse step -J^+ F is the minimum-norm least-squares solution LSQR converges to from x0 = 0) and
'lstsq' (:855-1022, without preconditioning; the second stage of BASELINE configs[0]): dense Jacobian by one batched
matvec, minimum-norm step by an SVD pseudo-inverse on the device, orbit by orbit as the reference does.
Everything else the reference offers (the SciPy minimize/root pass-throughs, dense solve, preconditioned lstsq) is outside the
north star's path and raises NotImplementedError here.

Keyword handling follows the reference: tol / ftol / maxiter / min_step / preconditioning may be scalars or
per-method lists consumed front to back; scipy_kwargs may be a dict or a list of dicts; unknown kwargs are ignored.
Status codes: -1 has-not-failed, 0 minimum step, 1 converged, 2 maxiter, 3 insufficient decrease (optimize.py:51-56).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .batch import BatchedOrbitKS, workspace
from .ks import OrbitKS

__all__ = ["hunt", "OrbitResult"]

_DEVICE_METHODS = ("adj", "gmres", "lsqr", "lsmr", "newton_descent", "lstsq")
_MESSAGES = {
    -1: "Has not failed yet: tolerance not met but no exit condition triggered.",
    0: "Failed to converge: minimum backtracking step size reached.",
    1: "Converged: cost below tolerance.",
    2: "Failed to converge: maximum number of iterations reached.",
    3: "Failed to converge: insufficient cost decrease.",
}


class OrbitResult(dict):
    """dict with attribute access, same keys as the reference's OrbitResult (optimize.py:24-79)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __dir__(self):
        return list(self.keys())


def _pop(value, name):
    if isinstance(value, list):
        try:
            return value.pop(0)
        except IndexError as ie:
            raise IndexError(": (tol, ftol, maxiter, min_step, preconditioning) need to be either scalar (or bool)"
                             " or list of same length as number of methods.") from ie
    return value


def _check_types(tol, ftol, maxiter, min_step, preconditioning):
    ok = (int, float, list, np.float64, np.int32)
    if not all(type(v) in ok for v in (tol, ftol, maxiter, min_step)) or type(preconditioning) not in (bool, list):
        raise TypeError("tol, ftol, min_step and maxiter must be scalar numbers (or list of scalars if multiple methods provided)")


def _adjoint_descent(orb: BatchedOrbitKS, tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6, preconditioning=False, **kwargs):
    """Batched optimize.py:367-509.  Returns (orbit batch, stats) with per-orbit tensors in stats."""
    _check_types(tol, ftol, maxiter, min_step, preconditioning)
    tol, ftol, maxiter = _pop(tol, "tol"), _pop(ftol, "ftol"), _pop(maxiter, "maxiter")
    min_step, preconditioning = _pop(min_step, "min_step"), _pop(preconditioning, "preconditioning")
    if not kwargs.get("backtracking", True):
        min_step = 1
    if maxiter is None or int(maxiter) <= 0:
        raise ValueError("maxiter must be a positive integer")  # the reference's loop would never stop (optimize.py:2183)
    orb._need_modes()
    L = _lib.lib()
    out = orb.copy()
    B, dev = out.batch, out.state.device
    status = torch.empty(B, dtype=torch.int32, device=dev)
    nit = torch.empty(B, dtype=torch.int32, device=dev)
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    step = torch.empty(B, dtype=torch.float64, device=dev)
    initial_cost = orb.cost()
    nbytes = L.ks_adjoint_descent_workspace_bytes(out.cls_id, out.n, out.m, B)
    ws = workspace(dev, nbytes)
    passes = ctypes.c_int(0)
    with torch.cuda.device(dev):
        _lib.check(L.ks_adjoint_descent(out.cls_id, out.n, out.m, B, out.state.data_ptr(), out.parameters.data_ptr(), out.cmask,
                                        float(tol), float(ftol), int(maxiter), float(min_step), int(bool(preconditioning)),
                                        float(kwargs.get("step_size", 1)), status.data_ptr(), nit.data_ptr(), cost.data_ptr(),
                                        step.data_ptr(), int(kwargs.get("check_every", 64)), ctypes.byref(passes), ws.data_ptr(),
                                        ws.numel(), torch.cuda.current_stream().cuda_stream))
    stats = {"method": "adj", "nit": nit, "costs": [initial_cost, cost], "maxiter": maxiter, "tol": tol, "status": status,
             "step_size": step, "passes": passes.value}
    return out, stats


def _run_method(orb, method, hunt_kwargs):
    if method == "adj":
        return _adjoint_descent(orb, **hunt_kwargs)
    if method in ("gmres", "lsqr", "lsmr", "newton_descent", "lstsq"):
        from .krylov import newton_krylov

        return newton_krylov(orb, method=method, **hunt_kwargs)
    raise NotImplementedError(
        f"hunt method '{method}' is outside the B200 hot path (device methods: {_DEVICE_METHODS}); "
        "use the reference implementation for SciPy minimize/root pass-throughs and dense lstsq/solve")


def hunt(orbit_instance, methods="adj", **kwargs):
    """Drop-in for orbithunter.optimize.hunt on an OrbitKS-family instance, or batched on a BatchedOrbitKS.

    Single orbit: returns OrbitResult(orbit, status, nit, costs, maxiter, tol, method, message) with the reference's
    scalar/list conventions.  Batch: the same keys hold per-orbit tensors (status/nit int32, costs a list of (B,)
    tensors: [initial, after method 1, ...]).
    """
    single = isinstance(orbit_instance, OrbitKS)
    if single:
        batch = orbit_instance.transform(to="modes")._batched()
    elif isinstance(orbit_instance, BatchedOrbitKS):
        batch = orbit_instance.transform("modes")
    else:
        raise TypeError("hunt expects an OrbitKS-family instance or a BatchedOrbitKS")
    hunt_kwargs = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in kwargs.items()}
    if isinstance(methods, str):
        methods = (methods,)
    hunt_kwargs.pop("methods", None)
    scipy_kwargs = hunt_kwargs.pop("scipy_kwargs", {})
    merged = {}
    for method in methods:
        if isinstance(scipy_kwargs, list):
            try:
                hunt_kwargs["scipy_kwargs"] = scipy_kwargs.pop(0)
            except IndexError as ie:
                raise IndexError("Providing keyword arguments for multiple methods requires a dict for each method,"
                                 " even if empty. This avoids accidentally passing incorrect keyword arguments") from ie
        else:
            hunt_kwargs["scipy_kwargs"] = scipy_kwargs
        batch, stats = _run_method(batch, method, hunt_kwargs)
        if not merged:
            merged = {k: ([v] if k in ("method", "maxiter", "tol", "nit") else v) for k, v in stats.items()}
            merged["costs"] = list(stats["costs"])
        else:
            for k in ("method", "maxiter", "tol", "nit"):
                merged[k].append(stats[k])
            merged["costs"].extend(stats["costs"])
            merged["status"] = stats["status"]
    final_tol = kwargs.get("tol", 1e-6)
    if isinstance(final_tol, list):
        final_tol = min(final_tol)
    final_cost = merged["costs"][-1]
    status = merged["status"].clone()
    status[(status == -1) & (final_cost <= final_tol)] = 1                          # optimize.py:355-358
    merged["status"] = status
    for k in ("method", "maxiter", "tol", "nit"):
        if len(merged[k]) == 1:
            merged[k] = merged[k][0]
    merged.pop("step_size", None)
    if not single:
        merged["message"] = None
        return OrbitResult(orbit=batch, **merged)
    # reference conventions for one orbit: python scalars
    def scal(v):
        return v[0].item() if torch.is_tensor(v) else v

    res = {k: ([scal(x) for x in v] if isinstance(v, list) else scal(v)) for k, v in merged.items()}
    res["status"] = int(res["status"])
    res["message"] = _MESSAGES[res["status"]]
    return OrbitResult(orbit=batch[0], **res)
