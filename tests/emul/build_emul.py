"""Build the CPU-thread emulation of the kernel sources -- TEST INFRASTRUCTURE (see cuda_emul.h).

    python tests/emul/build_emul.py   ->  tests/emul/libks_emul.so

The same .cu/.cuh files that nvcc compiles for sm_100a are compiled by g++ with -DKS_HOST_EMUL.  The resulting
library exports the identical C ABI but takes HOST pointers; only tests/test_emul_*.py load it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "orbithunter_b200", "csrc")
OUT = os.path.join(HERE, "libks_emul.so")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emul.h"),
                                                                   os.path.join(ROOT, "include", "ks_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not needs_build():
        return OUT
    flags = ["-std=c++20", "-O1", "-mfma", "-fPIC", "-pthread", "-DKS_HOST_EMUL", "-Wno-attributes", "-I", HERE, "-I", CSRC]
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in sources():  # one g++ per translation unit, all at once
        obj = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        cmd = ["g++", *flags, "-c", "-x", "c++", s, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", OUT, *objs])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
