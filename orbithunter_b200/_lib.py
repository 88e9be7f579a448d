"""Loader for the hand-written sm_100a kernels (libks_b200.so).  There is deliberately NO fallback:
if the library is missing or a call fails, the caller gets an exception."""
import ctypes as C
import os

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libks_b200.so")
_lib = None


class KsB200Error(RuntimeError):
    pass


def lib():
    """The bound CDLL.  Raises KsB200Error when the CUDA extension has not been built (python -m orbithunter_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KsB200Error(
                f"CUDA extension not built: {LIB_PATH} is missing. Run `python -m orbithunter_b200.build` "
                "(needs nvcc). orbithunter_b200 has no CPU fallback."
            )
        try:
            _lib = _cabi.bind(C.CDLL(LIB_PATH))
        except (OSError, AttributeError) as exc:
            raise KsB200Error(f"cannot load {LIB_PATH}: {exc}") from exc
    return _lib


def check(rc):
    if rc != 0:
        raise KsB200Error(f"libks_b200 error {rc}: {lib().ks_last_error().decode()}")
