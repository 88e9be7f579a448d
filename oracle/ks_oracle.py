"""CPU oracle for the spatiotemporal Kuramoto-Sivashinsky hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A batched NumPy/SciPy restatement of the reference's algorithm (chaotic-systems/orbithunter v1.3.4).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this module, and only as the checker or as the timed CPU arm; the product package ``orbithunter_b200`` never
imports it and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against golden vectors that
``oracle/make_golden.py`` produced by running the *unmodified* reference in the build container (through
``oracle/refshim.py``), and against the reference's own known-answer values (reference ``tests/test_basic.py:390-436``
transform norms, ``:622-651`` rmatvec/matvec/costgrad norms).  Krylov iteration *histories* are the exception:
no reference test pins them -- "parity unpinned" for GMRES/LSQR trajectories; the oracle's
Newton-Krylov loop calls the same third-party routine the reference calls (``scipy.sparse.linalg.gmres``/``lsqr``,
reference ``orbithunter/optimize.py:1321,1346``).

All arrays are float64 with arbitrary leading batch axes: ``modes[..., rows, cols]``, ``params[..., 3] = (T, L, S)``.
Reference file:line citations are into ``/root/reference/orbithunter``.

Third-party arithmetic restated/used: ``scipy.fft.rfft/irfft`` (norm="ortho"), requirements/default.txt pins only
``scipy>=1.4.1``; this image has SciPy 1.18.1 / NumPy 2.3.
"""
from __future__ import annotations

import numpy as np
from scipy.fft import irfft, rfft

ORBIT, RELATIVE, SHIFT_REFLECTION, ANTISYMMETRIC = 0, 1, 2, 3
CLASS_NAMES = {
    ORBIT: "OrbitKS",
    RELATIVE: "RelativeOrbitKS",
    SHIFT_REFLECTION: "ShiftReflectionOrbitKS",
    ANTISYMMETRIC: "AntisymmetricOrbitKS",
}
CLASS_IDS = {v: k for k, v in CLASS_NAMES.items()}
SQRT2 = np.sqrt(2.0)


def default_constraints(cls):
    """(t, x, s) constrained flags; ks/orbits.py:1731-1736 (s fixed) and :3035-3036 (Relative frees s)."""
    return (False, False, cls != RELATIVE)


def mode_shape(cls, n, m):
    """ks/orbits.py:1403-1412; SR :3677-3686; Anti :3254-3263."""
    k = m // 2 - 1
    return (max(n - 1, 1), 2 * k) if cls in (ORBIT, RELATIVE) else (max(n - 1, 1), k)


# ----------------------------------------------------------------------------------------------------------------
# frequency tables -- ks/orbits.py:5239-5307 (temporal_frequencies, spatial_frequencies, so2_coefficients)
# ----------------------------------------------------------------------------------------------------------------
def _so2(order):
    return {0: (1, 1), 1: (1, -1), 2: (-1, -1), 3: (-1, 1)}[order % 4]


def omega(T, n):
    """w_j = -2 pi j / T, j = 1..n/2-1, shape (..., h).  Negative: last field row is t=0 (ks/orbits.py:5265-5267)."""
    T = np.asarray(T, dtype=float)[..., None]
    j = np.fft.rfftfreq(n)[1:-1]
    return -1.0 * (2 * np.pi * n / T) * j


def wavenumber(L, m):
    """q_k = 2 pi k / L, k = 1..m/2-1, shape (..., k) (ks/orbits.py:5303)."""
    L = np.asarray(L, dtype=float)[..., None]
    return (2 * np.pi * m / L) * np.fft.rfftfreq(m)[1:-1]


def temporal_table(T, n, order):
    """[0, c1 w^p, c2 w^p] as a (..., n-1, 1) column (ks/orbits.py:5267-5272)."""
    w = omega(T, n) ** order
    c1, c2 = _so2(order)
    z = np.zeros(w.shape[:-1] + (1,))
    return np.concatenate((z, c1 * w, c2 * w), axis=-1)[..., :, None]


def spatial_table(L, m, order):
    """[c1 q^p | c2 q^p] as a (..., 1, 2k) row (ks/orbits.py:5303-5307)."""
    q = wavenumber(L, m) ** order
    c1, c2 = _so2(order)
    return np.concatenate((c1 * q, c2 * q), axis=-1)[..., None, :]


def swap_time(modes):
    """swap_modes(axis=0), ks/orbits.py:5185-5190: rows [0 | Re 1..h | Im 1..h] -> [0 | Im | Re]."""
    h = (modes.shape[-2] + 1) // 2 - 1
    return np.concatenate((modes[..., :1, :], modes[..., -h:, :], modes[..., 1:-h, :]), axis=-2) if h else modes


def swap_space(modes):
    """swap_modes(axis=1), ks/orbits.py:5179-5184."""
    k = modes.shape[-1] // 2
    return np.concatenate((modes[..., k:], modes[..., :k]), axis=-1)


# ----------------------------------------------------------------------------------------------------------------
# transforms -- ks/orbits.py:2212-2420 (OrbitKS), :3472-3579 (Antisymmetric), :3869-4005 (ShiftReflection)
# ----------------------------------------------------------------------------------------------------------------
def space_transform(field):
    """field (..., n, m) -> spatial_modes (..., n, m-2); ks/orbits.py:2350-2356."""
    X = SQRT2 * rfft(field, norm="ortho", axis=-1)[..., 1:-1]
    return np.concatenate((X.real, X.imag), axis=-1)


def inv_space_transform(smodes):
    """spatial_modes (..., n, m-2) -> field (..., n, m); ks/orbits.py:2403-2414."""
    k = smodes.shape[-1] // 2
    Z = smodes[..., :k] + 1j * smodes[..., k:]
    z = np.zeros(Z.shape[:-1] + (1,))
    return (1.0 / SQRT2) * irfft(np.concatenate((z, Z, z), axis=-1), norm="ortho", axis=-1)


def _pack_time(C):
    """rows [Re C[0..n/2-1] ; Im C[1..n/2-1]], sqrt2 on rows >= 1; ks/orbits.py:2246-2247."""
    out = np.concatenate((C.real[..., :-1, :], C.imag[..., 1:-1, :]), axis=-2)
    out[..., 1:, :] *= SQRT2
    return out


def _unpack_time(modes, n):
    """inverse of _pack_time with a zero Nyquist row; ks/orbits.py:2295-2306."""
    h = n // 2 - 1
    C = np.zeros(modes.shape[:-2] + (n // 2 + 1, modes.shape[-1]), dtype=complex)
    C[..., 0, :] = modes[..., 0, :]
    if h > 0:
        C[..., 1 : h + 1, :] = (modes[..., 1 : h + 1, :] + 1j * modes[..., h + 1 :, :]) / SQRT2
    return C


def _sr_mask(C, k):
    """ShiftReflection projection: zero C[even j, :k] and C[odd j, k:]; ks/orbits.py:3916-3917, :3997-3998."""
    C[..., ::2, :k] = 0
    C[..., 1::2, k:] = 0
    return C


def time_transform(smodes, cls=ORBIT):
    """spatial_modes (..., n, 2k) -> modes; ks/orbits.py:2245-2247, Anti :3505-3513, SR :3914-3933."""
    k = smodes.shape[-1] // 2
    if cls == ANTISYMMETRIC:
        return _pack_time(rfft(smodes[..., k:], norm="ortho", axis=-2))
    C = rfft(smodes, norm="ortho", axis=-2)
    if cls == SHIFT_REFLECTION:
        C = _sr_mask(C, k)
        C = C[..., :k] + C[..., k:]
    return _pack_time(C)


def inv_time_transform(modes, cls=ORBIT):
    """modes -> spatial_modes (..., n, 2k); ks/orbits.py:2295-2307, Anti :3560-3573, SR :3984-3999."""
    n = modes.shape[-2] + 1
    C = _unpack_time(modes, n)
    if cls == SHIFT_REFLECTION:
        k = modes.shape[-1]
        C = _sr_mask(np.concatenate((C, C), axis=-1), k)
    S = irfft(C, n=n, norm="ortho", axis=-2)
    if cls == ANTISYMMETRIC:
        S = np.concatenate((0 * S, S), axis=-1)
    return S


def to_field(modes, cls=ORBIT):
    """modes -> field; ks/orbits.py:2500-2524."""
    return inv_space_transform(inv_time_transform(modes, cls))


def to_modes(field, cls=ORBIT):
    """field -> modes; ks/orbits.py:2526-2551."""
    return time_transform(space_transform(field), cls)


# ----------------------------------------------------------------------------------------------------------------
# spectral derivatives -- ks/orbits.py:379-524 (dt, dx), SR/Anti overrides :3190-3227, :3583-3619
# ----------------------------------------------------------------------------------------------------------------
def dt(modes, T, order=1):
    """Time derivative in the modes basis; ks/orbits.py:423-430."""
    n = modes.shape[-2] + 1
    out = temporal_table(T, n, order) * modes
    return swap_time(out) if order % 2 else out


def dx_spatial(smodes, L, order=1):
    """Spatial derivative acting on a [Re | Im] array (spatial_modes, or modes of OrbitKS/Relative); :499-516."""
    m = smodes.shape[-1] + 2
    out = spatial_table(L, m, order) * smodes
    return swap_space(out) if order % 2 else out


def dx(modes, L, cls=ORBIT, order=1):
    """dx in the modes basis, returned in the modes basis (ks/orbits.py:442-524).

    For SR/Anti odd orders are computed in the spatial_modes basis and re-projected by the class's time transform
    (:3222-3224, :3614-3616) -- which gives zero; even orders use the first k table entries (:505-508).
    """
    if cls in (ORBIT, RELATIVE):
        return dx_spatial(modes, L, order)
    if order % 2:
        return time_transform(dx_spatial(inv_time_transform(modes, cls), L, order), cls)
    k = modes.shape[-1]
    m = 2 * k + 2
    return spatial_table(L, m, order)[..., :k] * modes


def _T(params):
    return np.asarray(params, dtype=float)[..., 0]


def _L(params):
    return np.asarray(params, dtype=float)[..., 1]


def _S(params):
    return np.asarray(params, dtype=float)[..., 2]


def _bc(x):
    """scalar-per-orbit -> broadcastable against (..., rows, cols)."""
    return np.asarray(x, dtype=float)[..., None, None]


# ----------------------------------------------------------------------------------------------------------------
# governing equation -- ks/orbits.py:526-555, :1899-1950; Rel :3038-3057; Anti :3283-3297; SR :3713-3727
# ----------------------------------------------------------------------------------------------------------------
def eqn_linear(modes, params, cls=ORBIT):
    """u_t + u_xx + u_xxxx [ - (S/T) u_x for Relative ]; ks/orbits.py:1918-1922, :3052-3057."""
    T, L, S = _T(params), _L(params), _S(params)
    out = dt(modes, T) + dx(modes, L, cls, 2) + dx(modes, L, cls, 4)
    if cls == RELATIVE:
        out = out - _bc(S / T) * dx(modes, L, cls, 1)
    return out


def rmatvec_linear(vmodes, params, cls=ORBIT):
    """-v_t + v_xx + v_xxxx [ + (S/T) v_x ]; ks/orbits.py:2041-2045, :3110-3115."""
    T, L, S = _T(params), _L(params), _S(params)
    out = -1.0 * dt(vmodes, T) + dx(vmodes, L, cls, 2) + dx(vmodes, L, cls, 4)
    if cls == RELATIVE:
        out = out + _bc(S / T) * dx(vmodes, L, cls, 1)
    return out


def nonlinear(ufield, vfield, params, cls=ORBIT):
    """1/2 T_cls( Dx( Fx(u .* v) ) ), no dealiasing; ks/orbits.py:1950, :3291-3293, :3721-3723."""
    return 0.5 * time_transform(dx_spatial(space_transform(ufield * vfield), _L(params), 1), cls)


def rnonlinear(ufield, vmodes, params, cls=ORBIT):
    """-T_cls( Fx( u .* Fx^-1( Dx( T_cls^-1 v ) ) ) ); ks/orbits.py:1977-1979 with dx(return_basis='field')."""
    w = inv_space_transform(dx_spatial(inv_time_transform(vmodes, cls), _L(params), 1))
    return -1.0 * to_modes(ufield * w, cls)


def eqn(modes, params, cls=ORBIT):
    """F(u); ks/orbits.py:546-551."""
    u = to_field(modes, cls)
    return eqn_linear(modes, params, cls) + nonlinear(u, u, params, cls)


def cost(modes, params, cls=ORBIT):
    """1/2 F.F; core.py:1004-1008."""
    F = eqn(modes, params, cls)
    return 0.5 * np.sum(F * F, axis=(-2, -1))


def _dot(a, b):
    return np.sum(a * b, axis=(-2, -1))


def matvec(modes, params, vmodes, vparams, cls=ORBIT, constraints=None):
    """J v; ks/orbits.py:680-709, Relative :2645-2660.  ``vparams`` are the increments (dT, dL, dS)."""
    ct, cx, cs = default_constraints(cls) if constraints is None else constraints
    T, L, S = _T(params), _L(params), _S(params)
    u = to_field(modes, cls)
    v = to_field(vmodes, cls)
    out = eqn_linear(vmodes, params, cls) + 2 * nonlinear(u, v, params, cls)
    vt, vx, vs = _T(vparams), _L(vparams), _S(vparams)
    if not ct:
        out = out + _bc(vt * (-1.0 / T)) * dt(modes, T)
    if not cx:
        dfdl = (
            _bc(-2.0 / L) * dx(modes, L, cls, 2)
            + _bc(-4.0 / L) * dx(modes, L, cls, 4)
            + _bc(-1.0 / L) * nonlinear(u, u, params, cls)
        )
        out = out + _bc(vx) * dfdl
    if cls == RELATIVE:
        udx = dx(modes, L, cls, 1)
        if not ct:
            out = out + _bc(vt * (-1.0 / T) * (-S / T)) * udx
        if not cx:
            out = out + _bc(vx * (-1.0 / L) * (-S / T)) * udx
        if not cs:
            out = out + _bc(vs * (-1.0 / T)) * udx
    return out


def rmatvec(modes, params, vmodes, cls=ORBIT, constraints=None):
    """J^T v -> (state, params3); ks/orbits.py:738-755, :1984-2024, Relative :3059-3094."""
    ct, cx, cs = default_constraints(cls) if constraints is None else constraints
    T, L, S = _T(params), _L(params), _S(params)
    u = to_field(modes, cls)
    state = rmatvec_linear(vmodes, params, cls) + rnonlinear(u, vmodes, params, cls)
    zero = np.zeros(np.shape(T))
    rt, rx, rs = zero, zero, zero
    udx = dx(modes, L, cls, 1) if cls == RELATIVE else None
    if not ct:
        g = dt(modes, T)
        if cls == RELATIVE:
            g = g + _bc(-S / T) * udx
        rt = (-1.0 / T) * _dot(g, vmodes)
    if not cx:
        nl = nonlinear(u, u, params, cls)
        if cls == RELATIVE:
            nl = nl + _bc(-S / T) * udx
        g = _bc(-2.0 / L) * dx(modes, L, cls, 2) + _bc(-4.0 / L) * dx(modes, L, cls, 4) + _bc(-1.0 / L) * nl
        rx = _dot(g, vmodes)
    if cls == RELATIVE and not cs:
        rs = (-1.0 / T) * _dot(udx, vmodes)
    return state, np.stack(np.broadcast_arrays(rt, rx, rs), axis=-1)


def precondition_multipliers(params, n, m, cls=ORBIT):
    """1 / (|w_j| + q^2 + q^4), rows [0, w, w]; ks/orbits.py:1061-1065."""
    T, L = _T(params), _L(params)
    ncols = mode_shape(cls, n, m)[1]
    return 1.0 / (
        np.abs(temporal_table(T, n, 1))
        + np.abs(spatial_table(L, m, 2)[..., :ncols])
        + spatial_table(L, m, 4)[..., :ncols]
    )


def precondition(state, gparams, params, n, m, cls=ORBIT, constraints=None):
    """state * multipliers; t *= T^-1, x *= L^-4 (pexp=(1,4)), s untouched; ks/orbits.py:1060-1089."""
    ct, cx, _ = default_constraints(cls) if constraints is None else constraints
    out = state * precondition_multipliers(params, n, m, cls)
    gp = np.array(gparams, dtype=float, copy=True)
    if not ct:
        gp[..., 0] = gp[..., 0] * _T(params) ** -1
    if not cx:
        gp[..., 1] = gp[..., 1] * _L(params) ** -4
    return out, gp


def costgrad(modes, params, cls=ORBIT, constraints=None, preconditioning=False, F=None):
    """(cost, grad_state, grad_params, F); ks/orbits.py:780-791.  grad = J^T F (optionally preconditioned)."""
    n, m = field_shape(modes, cls)
    if F is None:
        F = eqn(modes, params, cls)
    gs, gp = rmatvec(modes, params, F, cls, constraints)
    if preconditioning:
        gs, gp = precondition(gs, gp, params, n, m, cls, constraints)
    return 0.5 * _dot(F, F), gs, gp, F


def field_shape(modes, cls=ORBIT):
    """inverse of mode_shape; ks/orbits.py:1877-1879, SR/Anti :3417-3418."""
    r, c = modes.shape[-2:]
    return (r + 1, c + 2) if cls in (ORBIT, RELATIVE) else (r + 1, 2 * c + 2)


# ----------------------------------------------------------------------------------------------------------------
# cdof packing -- core.py:1617-1664, ks/orbits.py:305-377, core.py:1760-1796
# ----------------------------------------------------------------------------------------------------------------
def cdof(state, params, constraints):
    """[state.ravel(), free params in label order]; core.py:1631-1643."""
    free = [i for i in range(3) if not constraints[i]]
    flat = state.reshape(state.shape[:-2] + (-1,))
    return np.concatenate((flat, np.asarray(params, dtype=float)[..., free]), axis=-1)


def from_cdof(vec, shape, constraints, constants=None):
    """Unpack; constrained components come from ``constants`` (full (...,3) array) or 0.0; ks/orbits.py:343-377."""
    size = shape[0] * shape[1]
    state = vec[..., :size].reshape(vec.shape[:-1] + tuple(shape))
    params = np.zeros(vec.shape[:-1] + (3,))
    free = [i for i in range(3) if not constraints[i]]
    for slot, i in enumerate(free):
        params[..., i] = vec[..., size + slot]
    if constants is not None:
        for i in range(3):
            if constraints[i]:
                params[..., i] = np.asarray(constants, dtype=float)[..., i]
    return state, params


# ----------------------------------------------------------------------------------------------------------------
# solver drivers (single orbit; Python control flow mirrors the reference loop structure)
# ----------------------------------------------------------------------------------------------------------------
def _process_correction(stats, tol, maxiter, ftol, step_size, min_step, cost_, next_cost):
    """Returns accept flag; order tol -> min_step -> maxiter -> ftol; optimize.py:2174-2239."""
    stats["nit"] += 1
    if next_cost <= tol:
        stats["status"] = -1
        return True
    if step_size <= min_step:
        stats["status"] = 0
        return False
    if int(stats["nit"]) == int(maxiter):
        stats["status"] = 2
        return bool(next_cost < cost_)
    if (cost_ - next_cost) / max(cost_, next_cost, 1) < ftol:
        stats["status"] = 3
        return False
    return True


def adjoint_descent(modes, params, cls=ORBIT, constraints=None, tol=1e-6, ftol=1e-9, maxiter=10000, min_step=1e-6,
                    preconditioning=False, step_size=1.0):
    """optimize.py:367-509 for ONE orbit.  Returns (modes, params, stats) with stats = {nit, status, cost, step_size}.

    ``status`` is the value *after* the loop-exit fix-up (optimize.py:506-507) but before hunt's final promotion.
    """
    constraints = default_constraints(cls) if constraints is None else constraints
    modes = np.array(modes, dtype=float)
    params = np.array(params, dtype=float)
    stats = {"nit": 0, "status": -1}
    F = eqn(modes, params, cls)
    c = 0.5 * float(np.sum(F * F))
    step = float(step_size)
    while c > tol and stats["status"] == -1:
        _, gs, gp, _ = costgrad(modes, params, cls, constraints, preconditioning, F=F)
        nm, npar = modes - step * gs, params - step * gp
        nF = eqn(nm, npar, cls)
        nc = 0.5 * float(np.sum(nF * nF))
        while nc >= c and step > min_step:
            step /= 2
            nm, npar = modes - step * gs, params - step * gp
            nF = eqn(nm, npar, cls)
            nc = 0.5 * float(np.sum(nF * nF))
        if _process_correction(stats, tol, maxiter, ftol, step, min_step, c, nc):
            modes, params, c = nm, npar, nc
        F = nF
    final = float(cost(modes, params, cls))
    if final <= tol:
        stats["status"] = -1
    stats["cost"] = final
    stats["step_size"] = step
    return modes, params, stats


def normal_operator(modes, params, cls, constraints):
    """v -> A^T A v on cdof vectors, and A^T b with b = -F; optimize.py:1764-1797."""
    shape = modes.shape[-2:]

    def ata(v):
        vs, vp = from_cdof(np.asarray(v).ravel(), shape, constraints)
        jv = matvec(modes, params, vs, vp, cls, constraints)
        gs, gp = rmatvec(modes, params, jv, cls, constraints)
        return cdof(gs, gp, constraints)

    F = eqn(modes, params, cls)
    gs, gp = rmatvec(modes, params, -1.0 * F, cls, constraints)
    return ata, cdof(gs, gp, constraints)


def newton_krylov(modes, params, cls=ORBIT, constraints=None, method="gmres", tol=1e-6, ftol=1e-9, maxiter=10000,
                  min_step=1e-6, preconditioning=False, scipy_kwargs=None):
    """optimize.py:1196-1404 for ONE orbit with method in {gmres, lsqr, lsmr, lstsq}; inner solver is SciPy's, as in the reference."""
    from scipy.sparse.linalg import LinearOperator, gmres, lsmr, lsqr

    constraints = default_constraints(cls) if constraints is None else constraints
    scipy_kwargs = dict(scipy_kwargs or {})
    modes = np.array(modes, dtype=float)
    params = np.array(params, dtype=float)
    shape = modes.shape[-2:]
    n, m = field_shape(modes, cls)
    stats = {"nit": 0, "status": -1}
    c = float(cost(modes, params, cls))
    while c > tol and stats["status"] == -1:
        step = 1.0
        N = cdof(modes, params, constraints).size
        if method == "lstsq":
            # optimize.py:855-1022 (no preconditioning): dense Jacobian (ks/orbits.py:557-657), scipy.linalg.lstsq
            from scipy.linalg import lstsq as _dense_lstsq

            Jd = np.empty((modes.size, N))
            for col in range(N):
                e = np.zeros(N)
                e[col] = 1.0
                vs, vp = from_cdof(e, shape, constraints)
                Jd[:, col] = matvec(modes, params, vs, vp, cls, constraints).ravel()
            x = _dense_lstsq(Jd, -1.0 * eqn(modes, params, cls).ravel())[0]
        elif method in ("lsqr", "lsmr"):
            def mv(v):
                vs, vp = from_cdof(np.asarray(v).ravel(), shape, constraints)
                return matvec(modes, params, vs, vp, cls, constraints).ravel()

            def rmv(v):
                gs, gp = rmatvec(modes, params, np.asarray(v).reshape(shape), cls, constraints)
                return cdof(gs, gp, constraints)

            A = LinearOperator((modes.size, N), matvec=mv, rmatvec=rmv, dtype=float)
            b = -1.0 * eqn(modes, params, cls).ravel()
            x = (lsqr if method == "lsqr" else lsmr)(A, b, **scipy_kwargs)[0]
        else:
            if N == modes.size:
                # square system (every parameter constrained): J dx = -F itself, optimize.py:1803-1850
                def ata(v):
                    vs, vp = from_cdof(np.asarray(v).ravel(), shape, constraints)
                    return matvec(modes, params, vs, vp, cls, constraints).ravel()

                atb = -1.0 * eqn(modes, params, cls).ravel()
            else:
                ata, atb = normal_operator(modes, params, cls, constraints)
            A = LinearOperator((N, N), matvec=ata, rmatvec=ata, dtype=float)
            kw = dict(scipy_kwargs)
            if preconditioning:
                ones_s, ones_p = precondition(np.ones(shape), np.ones(3), params, n, m, cls, constraints)
                d = cdof(ones_s, ones_p, constraints)
                kw["M"] = LinearOperator((N, N), matvec=lambda v, d=d: np.asarray(v).ravel() * d, dtype=float)
            x = gmres(A, atb, **kw)[0]
        ds, dp = from_cdof(np.asarray(x).ravel(), shape, constraints)
        nm, npar = modes + ds, params + dp
        nc = float(cost(nm, npar, cls))
        while nc > c and step > min_step:
            step /= 2.0
            nm, npar = modes + step * ds, params + step * dp
            nc = float(cost(nm, npar, cls))
        if _process_correction(stats, tol, maxiter, ftol, step, min_step, c, nc):
            modes, params, c = nm, npar, nc
    final = float(cost(modes, params, cls))
    if final <= tol:
        stats["status"] = -1
    stats["cost"] = final
    return modes, params, stats


def hunt_status(stats, final_cost, tol):
    """hunt's final promotion -1 -> 1 when cost <= min(tol); optimize.py:351-358."""
    tol = min(tol) if isinstance(tol, (list, tuple)) else tol
    return 1 if (stats["status"] == -1 and final_cost <= tol) else stats["status"]


# ----------------------------------------------------------------------------------------------------------------
# seed generator -- ks/orbits.py:1738-1849 with default modulations ("laplace" space, "gaussian" time, xshift sqrt)
# ----------------------------------------------------------------------------------------------------------------
def populate_state(T, L, n, m, seed, cls=ORBIT):
    """Random modes from the legacy global MT19937 stream, as the reference's ``populate(attr='state', seed=..)``."""
    rows, cols = mode_shape(cls, n, m)
    tscale = int(np.round(T / 20.0))
    xscale = int(np.round(L / (2 * np.pi * np.sqrt(2))))
    xvar = max(np.sqrt(xscale), 1)
    tvar = max(np.sqrt(tscale), 1)
    rs = np.random.RandomState(seed)  # same stream as np.random.seed(seed); np.random.randn (:1779, :1793)
    space_ = np.abs(spatial_table(2 * np.pi, m, 1)[..., :cols]).astype(int)
    time_ = np.abs(temporal_table(2 * np.pi, n, 1)).astype(int)
    random_modes = rs.randn(rows, cols)
    xs = xscale + np.sqrt(time_)
    norm0 = np.linalg.norm(random_modes)
    modes_ = np.exp(-1.0 * np.abs(space_ - xs) / np.sqrt(xvar)) * random_modes
    modes_ = np.exp(-((time_ - tscale) ** 2) / (2 * tvar)) * modes_
    return (norm0 / np.linalg.norm(modes_)) * modes_
