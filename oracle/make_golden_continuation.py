"""Generate tests/golden/ks_continuation.npz by running the UNMODIFIED reference's continuation() -- TEST
INFRASTRUCTURE (container only; needs /root/reference).

    python oracle/make_golden_continuation.py

The reference's stored relative periodic orbit (tests/test_data.h5 'rpo', read with oracle/h5mini.py) is polished with
the reference's own hunt('gmres') and then continued in T with orbithunter.continuation.continuation
(continuation.py:59-136), every step re-converged by hunt('gmres').  Inputs, per-call keyword arguments and the
resulting family member are stored; tests/test_gpu_hunt.py repeats the same calls on the device path.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from oracle.h5mini import H5Mini  # noqa: E402

warnings.filterwarnings("ignore")
oh = refshim.load()
from orbithunter.continuation import continuation, discretization_continuation  # noqa: E402

KW = dict(methods="gmres", maxiter=40, tol=1e-6, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-8})


def main():
    h5 = H5Mini(os.path.join(refshim.REFERENCE_ROOT, "tests", "test_data.h5")).datasets()
    arr, attrs = h5["rpo/0"]
    o = oh.RelativeOrbitKS(state=arr, basis="modes", parameters=tuple(float(x) for x in attrs["parameters"]), discretization=(32, 32))
    r0 = oh.hunt(o, **KW)
    assert r0.status == 1, r0.status
    start = r0.orbit
    out = {"start/modes": start.state, "start/params": np.array(start.parameters, dtype=float), "start/cost": start.cost()}
    for tag, dT, step in (("up", 0.05, 0.02), ("down", -0.03, 0.02)):
        target = start.t + dT
        s = oh.RelativeOrbitKS(**vars(start))
        r = continuation(s, ("t", target), step_size=step, **{k: (dict(v) if isinstance(v, dict) else v) for k, v in KW.items()})
        out[f"{tag}/target"] = target
        out[f"{tag}/step"] = step
        out[f"{tag}/status"] = r.status
        out[f"{tag}/modes"] = r.orbit.state
        out[f"{tag}/params"] = np.array(r.orbit.parameters, dtype=float)
        out[f"{tag}/cost"] = r.orbit.cost()
        print(tag, r.status, r.orbit.parameters, r.orbit.cost())
    # discretization continuation (continuation.py:170-279) of the same orbit: 32x32 -> 32x36 (two re-converged steps)
    s = oh.RelativeOrbitKS(**vars(start))
    r = discretization_continuation(s, (32, 36), **{k: (dict(v) if isinstance(v, dict) else v) for k, v in KW.items()})
    out["disc/target"] = np.array([32, 36])
    out["disc/status"] = r.status
    out["disc/shape"] = np.array(r.orbit.discretization)
    out["disc/modes"] = r.orbit.state
    out["disc/params"] = np.array(r.orbit.parameters, dtype=float)
    out["disc/cost"] = r.orbit.cost()
    print("disc", r.status, r.orbit.discretization, r.orbit.parameters, r.orbit.cost())
    # resize (core.py:1167-1226; _pad / _truncate ks/orbits.py:1428-1564 and the subclass overrides) in the modes basis
    classes = {0: oh.OrbitKS, 1: oh.RelativeOrbitKS, 2: oh.ShiftReflectionOrbitKS, 3: oh.AntisymmetricOrbitKS}
    for cid, C in classes.items():
        o = C(parameters=(44.0, 44.0, 0.0)).populate(attr="state", seed=4, discretization=(16, 24))
        out[f"resize/{cid}/modes"] = o.state
        for shape in ((16, 24), (20, 24), (16, 30), (12, 20), (24, 16), (32, 32)):
            out[f"resize/{cid}/{shape[0]}x{shape[1]}"] = o.resize(*shape).state
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ks_continuation.npz"), **out)


if __name__ == "__main__":
    main()
