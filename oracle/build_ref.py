"""Carry the UNMODIFIED reference along as oracle/_ref -- TEST / BENCH INFRASTRUCTURE.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is present)

chaotic-systems/orbithunter is pure Python, so "building" it is copying its package directory, file for file, from where it
lies (/root/reference/orbithunter) into oracle/_ref/orbithunter.  oracle/_ref/ is git-ignored (no reference source enters
the history) but not gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist.  Only
oracle/refshim.py imports it: for golden generation here, and for `bench.py --impl reference` / `cpu_baseline` on the box.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ORBITHUNTER_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build():
    src_pkg = os.path.join(SRC, "orbithunter")
    if not os.path.isdir(src_pkg):
        return os.path.isdir(os.path.join(DST, "orbithunter"))  # on the GPU box: use what travelled
    dst_pkg = os.path.join(DST, "orbithunter")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("LICENSE", "LICENSE.txt", "LICENSE.md"):
        if os.path.exists(os.path.join(SRC, extra)):
            shutil.copy(os.path.join(SRC, extra), os.path.join(DST, extra))
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "ready" if ok else "reference tree not found")
    sys.exit(0 if ok else 1)
