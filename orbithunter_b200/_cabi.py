"""ctypes signatures of the C ABI declared in include/ks_b200.h (one place, used by the loader and by the tests)."""
import ctypes as C

_P = C.c_void_p
_SIG = {
    "ks_b200_abi_version": (C.c_int, []),
    "ks_last_error": (C.c_char_p, []),
    "ks_debug_force_generic": (None, [C.c_int]),
    "ks_debug_warp32_variant": (None, [C.c_int]),
    "ks_debug_dense_jacobian_mode": (None, [C.c_int]),
    "ks_debug_dense_outer_blocks": (None, [C.c_int]),
    "ks_debug_warp32_stagger": (None, [C.c_int]),
    "ks_debug_warp32_pingpong": (None, [C.c_int]),
    "ks_debug_tile": (None, [C.c_int, C.c_int]),
    "ks_debug_gmres_cluster": (None, [C.c_int]),
    "ks_debug_tile_tma": (None, [C.c_int]),
    "ks_debug_operator_cache": (None, [C.c_int]),
    "ks_debug_adjoint_inner": (None, [C.c_int]),
    "ks_mode_size": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ks_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ks_populate_state": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "ks_populate_state_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "ks_resize_modes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "ks_rescale_field": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_double, C.c_int, _P, _P]),
    "ks_preprocess": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "ks_transform": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "ks_eqn": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "ks_matvec": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_uint32, _P, _P, C.c_size_t, _P]),
    "ks_jacobian_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]),
    "ks_jacobian": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_uint32, _P, _P, C.c_size_t, _P]),
    "ks_rmatvec": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_uint32, _P, _P, _P, C.c_size_t, _P]),
    "ks_costgrad": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_uint32, C.c_int, _P, _P, _P, _P, _P,
                              C.c_size_t, _P]),
    "ks_precondition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_uint32, C.c_double, C.c_double, _P, _P,
                                  _P]),
    "ks_deriv": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "ks_adjoint_descent_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ks_adjoint_descent": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_uint32, C.c_double, C.c_double, C.c_int,
                                     C.c_double, C.c_int, C.c_double, _P, _P, _P, _P, C.c_int, C.POINTER(C.c_int), _P,
                                     C.c_size_t, _P]),
    "ks_newton_krylov_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ks_newton_krylov": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_uint32, C.c_int, C.c_double, C.c_double,
                                   C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                   C.c_double, _P, _P, _P, C.POINTER(C.c_longlong), C.c_int, _P, C.c_size_t, _P]),
    "ks_newton_dense_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ks_newton_dense": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_uint32, C.c_double, C.c_double, C.c_int,
                                  C.c_double, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_longlong), _P, C.c_size_t, _P]),
}



class KsPopulateOptions(C.Structure):
    """include/ks_b200.h KsPopulateOptions (host struct)."""
    _fields_ = [("spatial_modulation", C.c_int), ("temporal_modulation", C.c_int), ("xshift", C.c_int), ("reserved", C.c_int),
                ("tscale", C.c_double), ("xscale", C.c_double), ("xvar", C.c_double), ("tvar", C.c_double)]


SPATIAL_MODULATION_IDS = {"gaussian": 1, "laplace": 2, "laplace_sqrt": 3, "plateau_linear": 4, "exponential_linear": 5, "flat_top": 6}
TEMPORAL_MODULATION_IDS = {"gaussian": 1, "laplace": 2, "flat_top": 3, "laplace_plateau": 4, "truncate": 5}
XSHIFT_IDS = {"sqrt": 1, "log": 2, "lin": 3}
EXPORTS = tuple(_SIG)


def bind(lib):
    """Attach restype/argtypes for every exported symbol; raises AttributeError if one is missing."""
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
