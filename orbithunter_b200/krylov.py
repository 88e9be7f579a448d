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
    """hunt('lstsq') (reference optimize.py:855-1022, the second stage of BASELINE configs[0]), orbit by orbit as the reference
    does: dense Jacobian (one batched matvec over the identity on the device), dx = the minimum-norm least-squares solution
    of J dx = -F (the reference calls scipy.linalg.lstsq, LAPACK gelsd with cond = eps; here the SVD pseudo-inverse of
    torch.linalg on the device with the same cut-off), full step then halving while the cost increases (`>`), and
    _process_correction (:2115-2239).  Cost O(N^3) per iteration and orbit, like the reference: meant for a handful of
    orbits; large batches belong to 'lsqr' / 'gmres' / 'newton_descent' (matrix-free)."""
    if preconditioning:
        raise NotImplementedError("hunt('lstsq', preconditioning=True) (right-preconditioned dense least squares) is not provided")
    if not kwargs.get("backtracking", True):
        min_step = 1
    out = orb.copy()
    B, dev = out.batch, out.state.device
    initial_cost = orb.cost()
    status = torch.full((B,), -1, dtype=torch.int32, device=dev)
    nit = torch.zeros(B, dtype=torch.int32, device=dev)
    cost_out = initial_cost.clone()
    eps = float(np.finfo(np.float64).eps)
    for b in range(B):
        o = out[b]
        F = o.eqn()
        cost = F.cost(evaleqn=False)
        st, it = -1, 0
        while cost > tol and st == -1:
            step = kwargs.get("step_size", 1)
            J = torch.as_tensor(o.jacobian(), device=dev)
            rhs = torch.as_tensor(-1.0 * F.state.ravel(), device=dev)
            dx = (torch.linalg.pinv(J, rtol=eps) @ rhs).cpu().numpy()
            dxo = o.from_numpy_array(dx)
            nxt = o.increment(dxo, step_size=step)
            nF = nxt.eqn()
            ncost = nF.cost(evaleqn=False)
            while ncost > cost and step > min_step:
                step /= 2.0
                nxt = o.increment(dxo, step_size=step)
                nF = nxt.eqn()
                ncost = nF.cost(evaleqn=False)
            it += 1                                                                        # _process_correction
            if ncost <= tol:
                st, o, F, cost = -1, nxt, nF, ncost
            elif step <= min_step:
                st = 0
            elif it == int(maxiter):
                st = 2
                if ncost < cost:
                    o, F, cost = nxt, nF, ncost
            elif (cost - ncost) / max([cost, ncost, 1]) < ftol:
                st = 3
            else:
                o, F, cost = nxt, nF, ncost
        out.state[b] = torch.as_tensor(o.state, device=dev)
        out.parameters[b] = torch.as_tensor(np.asarray(o.parameters, dtype=np.float64), device=dev)
        status[b], nit[b], cost_out[b] = st, it, float(cost)
    stats = {"method": "lstsq", "nit": nit, "costs": [initial_cost, cost_out], "maxiter": maxiter, "tol": tol, "status": status}
    return out, stats
