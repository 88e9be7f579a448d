"""CPU: the C-ABI library is present, loads, and exports every symbol include/ks_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ks_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ks_[a-z0-9_]+)\s*\(", text)))


def test_header_and_bindings_agree():
    from orbithunter_b200 import _cabi

    assert _declared() == sorted(_cabi.EXPORTS)


def test_library_builds_loads_and_exports_all_symbols():
    from orbithunter_b200 import build as _build

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    from orbithunter_b200 import _lib

    L = _lib.lib()
    assert L.ks_b200_abi_version() == 1
    assert L.ks_mode_size(0, 32, 32) == 31 * 30 and L.ks_mode_size(2, 64, 64) == 63 * 31
    assert L.ks_mode_size(0, 33, 32) == 0  # odd sizes are rejected
    assert L.ks_workspace_bytes(0, 32, 32, 0) == 0


def test_missing_extension_fails_loudly(monkeypatch):
    from orbithunter_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "orbithunter_b200", "lib", "does_not_exist.so"))
    with pytest.raises(_lib.KsB200Error):
        _lib.lib()


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import torch

    from orbithunter_b200 import BatchedOrbitKS, KsB200Error

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(KsB200Error):
        BatchedOrbitKS(np.zeros((1, 31, 30)), np.array([[44.0, 33.4, 0.0]]))
