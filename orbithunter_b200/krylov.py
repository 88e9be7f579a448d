"""Newton-Krylov methods of hunt() on the device: 'gmres', 'lsqr', 'lsmr' and the matrix-free 'newton_descent' / 'lstsq'.

Reference: orbithunter/optimize.py _scipy_sparse_linalg_solver_wrapper :1196-1404 (+ _sparse_linalg_factory :1682-1850)
and _newton_descent :653-852.  ``scipy_kwargs`` keeps SciPy's names: gmres -> restart, maxiter (restart cycles), rtol,
atol; lsqr -> atol, btol, conlim, iter_lim (damp must be 0); lsmr -> atol, btol, conlim, maxiter (damp 0, no x0).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .batch import BatchedOrbitKS, workspace


def _pop(value):
    return value.pop(0) if isinstance(value, list) else value


def newton_krylov(orb: BatchedOrbitKS, method="gmres", tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6,
                  preconditioning=False, scipy_kwargs=None, **kwargs):
    from .optimize import _check_types

    _check_types(tol, ftol, maxiter, min_step, preconditioning)
    tol, ftol, maxiter, min_step, preconditioning = (_pop(v) for v in (tol, ftol, maxiter, min_step, preconditioning))
    if not kwargs.get("backtracking", True):
        min_step = 1
    if maxiter is None or int(maxiter) <= 0:
        # the reference's loops only stop at nit == maxiter (optimize.py:2183), so maxiter <= 0 never terminates there
        raise ValueError("maxiter must be a positive integer")
    if method != "lstsq" and kwargs.get("step_size", 1) != 1:
        # the reference's Newton-type loops restart every outer iteration at step_size (optimize.py:756, :1315); the device
        # line search starts at the full step
        raise NotImplementedError(f"hunt('{method}', step_size != 1) is not provided on the device path")
    if method == "newton_descent" and (preconditioning or kwargs.get("approximation", False)):
        # optimize.py:757-815 right-preconditions the pseudo-inverse step / runs the approximate-inverse inner loop
        raise NotImplementedError("hunt('newton_descent') with preconditioning=True or approximation=True is not provided")
    sk = dict(scipy_kwargs or {})
    orb._need_modes()
    out = orb.copy()
    B, dev = out.batch, out.state.device
    N = out.state.shape[1] * out.state.shape[2] + len(out.free_indices())
    if method == "gmres":
        mid, restart = 0, int(sk.pop("restart", 20) or 20)
        inner = sk.pop("maxiter", None)
        inner = 10 * N if inner is None else int(inner)
        rtol, atol, btol, conlim = float(sk.pop("rtol", 1e-5)), float(sk.pop("atol", 0.0)), 0.0, 0.0
    elif method == "lsqr":
        mid, restart = 1, 1
        if sk.pop("damp", 0.0):
            raise NotImplementedError("lsqr damp != 0 is not provided on the device path")
        inner = sk.pop("iter_lim", None)
        inner = 2 * N if inner is None else int(inner)
        rtol, atol, btol, conlim = 0.0, float(sk.pop("atol", 1e-6)), float(sk.pop("btol", 1e-6)), float(sk.pop("conlim", 1e8))
        preconditioning = False  # the reference never preconditions lsqr/lsmr (optimize.py:1317-1321)
    elif method == "lsmr":
        mid, restart = 2, 1
        if sk.pop("damp", 0.0):
            raise NotImplementedError("lsmr damp != 0 is not provided on the device path")
        if sk.pop("x0", None) is not None:
            raise NotImplementedError("lsmr x0 is not provided on the device path")
        inner = sk.pop("maxiter", None)
        inner = min(N, out.state.shape[1] * out.state.shape[2]) if inner is None else int(inner)  # lsmr.py: min(m, n)
        rtol, atol, btol, conlim = 0.0, float(sk.pop("atol", 1e-6)), float(sk.pop("btol", 1e-6)), float(sk.pop("conlim", 1e8))
        preconditioning = False  # the reference never preconditions lsqr/lsmr (optimize.py:1317-1321)
    elif method == "lstsq":
        return _dense_lstsq(orb, tol, ftol, maxiter, min_step, preconditioning, **kwargs)
    elif method == "newton_descent":
        # dx = pinv(J) (-F): the minimum-norm least-squares step, which LSQR from x0 = 0 converges to.  Stated
        # tolerance of the restatement: atol = btol = 1e-13, up to 4 N Golub-Kahan steps, no condition limit.
        mid, restart, preconditioning = 1, 1, False
        inner, rtol, atol, btol, conlim = 4 * N, 0.0, 1e-13, 1e-13, 0.0
    else:
        raise ValueError(f"Unknown solver {method}")
    if sk:
        raise TypeError(f"unsupported scipy_kwargs for {method}: {sorted(sk)}")
    L = _lib.lib()
    status = torch.empty(B, dtype=torch.int32, device=dev)
    nit = torch.empty(B, dtype=torch.int32, device=dev)
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    initial_cost = orb.cost()
    nbytes = L.ks_newton_krylov_workspace_bytes(out.cls_id, out.n, out.m, B, mid, restart)
    ws = workspace(dev, nbytes)
    counters = (ctypes.c_longlong * 4)()
    with torch.cuda.device(dev):
        _lib.check(L.ks_newton_krylov(out.cls_id, out.n, out.m, B, out.state.data_ptr(), out.parameters.data_ptr(), out.cmask, mid,
                                      float(tol), float(ftol), int(maxiter), float(min_step), int(bool(preconditioning)), restart,
                                      inner, rtol, atol, btol, conlim, status.data_ptr(), nit.data_ptr(), cost.data_ptr(), counters,
                                      int(kwargs.get("check_every", 8)), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream))
    stats = {"method": method, "nit": nit, "costs": [initial_cost, cost], "maxiter": maxiter, "tol": tol, "status": status,
             "operator_applications": counters[0], "solver_steps": counters[1], "line_search_evals": counters[2],
             "outer_iterations": counters[3]}
    return out, stats


def _dense_lstsq(orb: BatchedOrbitKS, tol, ftol, maxiter, min_step, preconditioning, **kwargs):
    """hunt('lstsq') (reference optimize.py:855-1022, the second stage of BASELINE configs[0]) for the whole batch on the
    device: ks_newton_dense builds the dense Jacobians of the orbits still in the loop with one batched matvec over the
    identity, takes the minimum-norm solution of J dx = -F (the reference: scipy.linalg.lstsq) as J^T (J J^T)^-1 (-F) through a
    blocked Cholesky factorisation with iterative refinement, then the reference's line search (full step, halve while the cost
    increases) and _process_correction (:2115-2239).  ``chunk`` (kwarg; default 1 024 at 32x32, fewer for larger grids: 24 GB of workspace) orbits are factored at a time."""
    if preconditioning:
        raise NotImplementedError("hunt('lstsq', preconditioning=True) (right-preconditioned dense least squares) is not provided")
    if not kwargs.get("backtracking", True):
        min_step = 1
    if kwargs.get("step_size", 1) != 1:
        raise NotImplementedError("hunt('lstsq', step_size != 1) is not provided on the device path")
    out = orb.copy()
    B, dev = out.batch, out.state.device
    L = _lib.lib()
    nm = out.state.shape[1] * out.state.shape[2]
    # orbits factored at a time: three Nm x N arrays each; the default keeps them within 24 GB (1 024 orbits at 32x32, where
    # larger chunks measured faster: 38.5 / 35.4 / 33.8 us per dense solve at 256 / 512 / 1 024)
    default_chunk = max(1, min(1024, int(24e9 // (3 * nm * (nm + 3) * 8))))
    chunk = max(1, min(int(kwargs.get("chunk", default_chunk)), B))
    refine = int(kwargs.get("refine", 2))
    status = torch.empty(B, dtype=torch.int32, device=dev)
    nit = torch.empty(B, dtype=torch.int32, device=dev)
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    initial_cost = orb.cost()
    nbytes = L.ks_newton_dense_workspace_bytes(out.cls_id, out.n, out.m, B, chunk)
    ws = workspace(dev, nbytes)
    counters = (ctypes.c_longlong * 4)()
    with torch.cuda.device(dev):
        _lib.check(L.ks_newton_dense(out.cls_id, out.n, out.m, B, out.state.data_ptr(), out.parameters.data_ptr(), out.cmask,
                                     float(tol), float(ftol), int(maxiter), float(min_step), chunk, refine, status.data_ptr(),
                                     nit.data_ptr(), cost.data_ptr(), counters, ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream))
    stats = {"method": "lstsq", "nit": nit, "costs": [initial_cost, cost], "maxiter": maxiter, "tol": tol, "status": status,
             "jacobian_columns": counters[0], "dense_solves": counters[1], "line_search_evals": counters[2],
             "outer_iterations": counters[3]}
    return out, stats
