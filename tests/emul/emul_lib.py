"""Load the CPU-thread emulation build of the kernels and call its C ABI with NumPy (host) pointers -- tests only."""
import ctypes as C
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

_lib = None


def lib():
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("build_emul", os.path.join(HERE, "build_emul.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        path = mod.build()
        spec2 = importlib.util.spec_from_file_location("_cabi", os.path.join(ROOT, "orbithunter_b200", "_cabi.py"))
        cabi = importlib.util.module_from_spec(spec2)
        spec2.loader.exec_module(cabi)
        _lib = cabi.bind(C.CDLL(path))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _ws(L, cls, n, m, B):
    nbytes = L.ks_workspace_bytes(cls, n, m, B)
    return np.zeros(max(nbytes // 8, 1)), nbytes


def _chk(L, rc):
    if rc != 0:
        raise RuntimeError(f"ks error {rc}: {L.ks_last_error().decode()}")


def mode_shape(cls, n, m):
    return (n - 1, m - 2) if cls in (0, 1) else (n - 1, m // 2 - 1)


def basis_shape(cls, n, m, basis):
    return {0: (n, m), 1: (n, m - 2), 2: mode_shape(cls, n, m)}[basis]


def populate(cls, n, m, seeds, params):
    L = lib()
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    params = np.ascontiguousarray(params, dtype=float)
    out = np.zeros((len(seeds),) + mode_shape(cls, n, m))
    _chk(L, L.ks_populate_state(cls, n, m, len(seeds), _p(seeds), _p(params), _p(out), None))
    return out


def transform(cls, n, m, arr, frm, to):
    L = lib()
    arr = np.ascontiguousarray(arr, dtype=float)
    B = arr.shape[0]
    out = np.zeros((B,) + basis_shape(cls, n, m, to))
    ws, nb = _ws(L, cls, n, m, B)
    _chk(L, L.ks_transform(cls, n, m, B, frm, to, _p(arr), _p(out), _p(ws), nb, None))
    return out


def eqn(cls, n, m, modes, params):
    L = lib()
    modes = np.ascontiguousarray(modes, dtype=float)
    params = np.ascontiguousarray(params, dtype=float)
    B = modes.shape[0]
    F = np.zeros_like(modes)
    cost = np.zeros(B)
    ws, nb = _ws(L, cls, n, m, B)
    _chk(L, L.ks_eqn(cls, n, m, B, _p(modes), _p(params), _p(F), _p(cost), _p(ws), nb, None))
    return F, cost


def matvec(cls, n, m, modes, params, v, vp, cmask):
    L = lib()
    modes, params, v, vp = (np.ascontiguousarray(x, dtype=float) for x in (modes, params, v, vp))
    B = modes.shape[0]
    out = np.zeros_like(modes)
    ws, nb = _ws(L, cls, n, m, B)
    _chk(L, L.ks_matvec(cls, n, m, B, _p(modes), _p(params), _p(v), _p(vp), cmask, _p(out), _p(ws), nb, None))
    return out


def jacobian(cls, n, m, modes, params, cmask):
    """[B, cdof, Nm]: column-major Jacobians."""
    L = lib()
    modes, params = (np.ascontiguousarray(x, dtype=float) for x in (modes, params))
    B, nm = modes.shape[0], modes.shape[1] * modes.shape[2]
    ncol = nm + sum(1 for i in range(3) if not (cmask >> i) & 1)
    out = np.zeros((B, ncol, nm))
    nb = L.ks_jacobian_workspace_bytes(cls, n, m, B, cmask)
    ws = np.zeros(nb // 8 + 1)
    _chk(L, L.ks_jacobian(cls, n, m, B, _p(modes), _p(params), cmask, _p(out), _p(ws), nb, None))
    return out


def rmatvec(cls, n, m, modes, params, v, cmask):
    L = lib()
    modes, params, v = (np.ascontiguousarray(x, dtype=float) for x in (modes, params, v))
    B = modes.shape[0]
    out = np.zeros_like(modes)
    op = np.zeros((B, 3))
    ws, nb = _ws(L, cls, n, m, B)
    _chk(L, L.ks_rmatvec(cls, n, m, B, _p(modes), _p(params), _p(v), cmask, _p(out), _p(op), _p(ws), nb, None))
    return out, op


def costgrad(cls, n, m, modes, params, cmask, precond=False, want_f=True):
    L = lib()
    modes, params = (np.ascontiguousarray(x, dtype=float) for x in (modes, params))
    B = modes.shape[0]
    g = np.zeros_like(modes)
    gp = np.zeros((B, 3))
    cost = np.zeros(B)
    F = np.zeros_like(modes) if want_f else None
    ws, nb = _ws(L, cls, n, m, B)
    _chk(L, L.ks_costgrad(cls, n, m, B, _p(modes), _p(params), cmask, int(precond), _p(cost), _p(g), _p(gp), _p(F), _p(ws),
                          nb, None))
    return cost, g, gp, F


def adjoint_descent(cls, n, m, modes, params, cmask, tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6, precond=False,
                    step_size=1.0, check_every=16):
    L = lib()
    modes = np.array(modes, dtype=float, order="C")
    params = np.array(params, dtype=float, order="C")
    B = modes.shape[0]
    status = np.zeros(B, dtype=np.int32)
    nit = np.zeros(B, dtype=np.int32)
    cost = np.zeros(B)
    step = np.zeros(B)
    nb = L.ks_adjoint_descent_workspace_bytes(cls, n, m, B)
    ws = np.zeros(nb // 8 + 1)
    passes = C.c_int(0)
    _chk(L, L.ks_adjoint_descent(cls, n, m, B, _p(modes), _p(params), cmask, tol, ftol, maxiter, min_step, int(precond), step_size,
                                 _p(status), _p(nit), _p(cost), _p(step), check_every, C.byref(passes), _p(ws), nb, None))
    return modes, params, status, nit, cost, step, passes.value


def newton_krylov(cls, n, m, modes, params, cmask, method, tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6, precond=False,
                  restart=20, inner=None, rtol=1e-5, atol=0.0, btol=1e-6, conlim=1e8, check_every=4):
    L = lib()
    modes = np.array(modes, dtype=float, order="C")
    params = np.array(params, dtype=float, order="C")
    B = modes.shape[0]
    status = np.zeros(B, dtype=np.int32)
    nit = np.zeros(B, dtype=np.int32)
    cost = np.zeros(B)
    nb = L.ks_newton_krylov_workspace_bytes(cls, n, m, B, method, restart)
    ws = np.zeros(nb // 8 + 1)
    counters = (C.c_longlong * 4)()
    _chk(L, L.ks_newton_krylov(cls, n, m, B, _p(modes), _p(params), cmask, method, tol, ftol, maxiter, min_step, int(precond), restart,
                               inner, rtol, atol, btol, conlim, _p(status), _p(nit), _p(cost), counters, check_every, _p(ws), nb, None))
    return modes, params, status, nit, cost, list(counters)


def newton_dense(cls, n, m, modes, params, cmask, tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6, chunk=2, refine=2):
    L = lib()
    modes = np.array(modes, dtype=float, order="C")
    params = np.array(params, dtype=float, order="C")
    B = modes.shape[0]
    status = np.zeros(B, dtype=np.int32)
    nit = np.zeros(B, dtype=np.int32)
    cost = np.zeros(B)
    nb = L.ks_newton_dense_workspace_bytes(cls, n, m, B, chunk)
    ws = np.zeros(nb // 8 + 1)
    counters = (C.c_longlong * 4)()
    _chk(L, L.ks_newton_dense(cls, n, m, B, _p(modes), _p(params), cmask, tol, ftol, maxiter, min_step, chunk, refine, _p(status), _p(nit),
                              _p(cost), counters, _p(ws), nb, None))
    return modes, params, status, nit, cost, list(counters)


def deriv(cls, n, m, arr, params, axis, order):
    L = lib()
    arr = np.ascontiguousarray(arr, dtype=float)
    params = np.ascontiguousarray(params, dtype=float)
    out = np.zeros_like(arr)
    _chk(L, L.ks_deriv(cls, n, m, arr.shape[0], axis, order, arr.shape[1], arr.shape[2], _p(arr), _p(params), _p(out), None))
    return out


def precondition(cls, n, m, state, gparams, params, cmask, pexp=(1.0, 4.0)):
    L = lib()
    state, gparams, params = (np.ascontiguousarray(x, dtype=float) for x in (state, gparams, params))
    out, gp = np.zeros_like(state), np.zeros_like(gparams)
    _chk(L, L.ks_precondition(cls, n, m, state.shape[0], _p(state), _p(gparams), _p(params), cmask, pexp[0], pexp[1], _p(out), _p(gp), None))
    return out, gp


def populate_ex(cls, n, m, seeds, params, spatial="laplace", temporal="gaussian", xshift="sqrt", tscale=None, xscale=None, xvar=None,
                tvar=None):
    import ctypes

    L = lib()
    spec = importlib.util.spec_from_file_location("_cabi", os.path.join(ROOT, "orbithunter_b200", "_cabi.py"))
    cabi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cabi)
    nan = float("nan")
    opt = cabi.KsPopulateOptions(cabi.SPATIAL_MODULATION_IDS.get(spatial, 0), cabi.TEMPORAL_MODULATION_IDS.get(temporal, 0),
                                 cabi.XSHIFT_IDS.get(xshift, 0), 0, *(nan if v is None else float(v) for v in (tscale, xscale, xvar, tvar)))
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    params = np.ascontiguousarray(params, dtype=float)
    out = np.zeros((len(seeds),) + mode_shape(cls, n, m))
    _chk(L, L.ks_populate_state_ex(cls, n, m, len(seeds), _p(seeds), _p(params), ctypes.addressof(opt), _p(out), None))
    return out


def resize_modes(cls, n, m, n2, m2, modes):
    L = lib()
    modes = np.ascontiguousarray(modes, dtype=float)
    out = np.zeros((modes.shape[0],) + mode_shape(cls, n2, m2))
    _chk(L, L.ks_resize_modes(cls, n, m, n2, m2, modes.shape[0], _p(modes), _p(out), None))
    return out


def rescale_field(field, magnitude, method):
    L = lib()
    f = np.array(field, dtype=float, order="C")
    norms = np.zeros(f.shape[0])
    _chk(L, L.ks_rescale_field(f.shape[1], f.shape[2], f.shape[0], _p(f), float(magnitude), {"inf": 0, "L1": 1, "L2": 2, "LP": 3}[method],
                               _p(norms), None))
    return f, norms


def preprocess(cls, n, m, modes, params):
    L = lib()
    modes = np.ascontiguousarray(modes, dtype=float)
    params = np.ascontiguousarray(params, dtype=float)
    code = np.zeros(modes.shape[0], dtype=np.int32)
    norms = np.zeros((modes.shape[0], 2))
    _chk(L, L.ks_preprocess(cls, n, m, modes.shape[0], _p(modes), _p(params), _p(code), _p(norms), None))
    return code, norms
