"""Import the UNMODIFIED reference (chaotic-systems/orbithunter, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` to generate the committed golden vectors under
``tests/golden/`` and by the container-only cross-check tests (skipped when /root/reference is absent, as it is
on the GPU box).  Nothing in the product package imports this module.

The reference imports ``h5py`` (orbithunter/core.py:3, orbithunter/io.py:3) and matplotlib
(orbithunter/ks/orbits.py:6,10) at module import time; neither is installed here and both are only used for file
I/O and plotting, so empty stub modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ORBITHUNTER_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "orbithunter")):
    # the GPU box has no /root/reference: use the unmodified copy oracle/build_ref.py made (oracle/_ref, git-ignored)
    _carried = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isdir(os.path.join(_carried, "orbithunter")):
        REFERENCE_ROOT = _carried


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "orbithunter"))


def load():
    """Return the reference ``orbithunter`` package (imported once, unmodified)."""
    if "orbithunter" in sys.modules and getattr(sys.modules["orbithunter"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["orbithunter"]
    if not available():
        raise ImportError(f"reference tree not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for name in ("h5py", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                if name == "mpl_toolkits.axes_grid1":
                    stub.make_axes_locatable = None
                sys.modules[name] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import orbithunter  # noqa: E402

    return orbithunter
