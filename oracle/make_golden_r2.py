"""Round-2 golden vectors from the UNMODIFIED reference -- TEST INFRASTRUCTURE (container only).

    python oracle/make_golden_r2.py [--only cfg2,cfg3,square,lsqr,modulation,resize,preprocess,cfg0] [--procs 8]

Writes tests/golden/ks_hunt_r2.npz (hunt end states at BASELINE configs[2]/[3] settings, the fully constrained square
GMRES branch, lsmr) and tests/golden/ks_cfg0_table.npz (per-seed status table of configs[0] on seeds 0..63).
Everything stored is an output of chaotic-systems/orbithunter v1.3.4 itself (oracle/refshim.py); see make_golden.py.
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

warnings.filterwarnings("ignore")
oh = refshim.load()
GOLD = os.path.join(ROOT, "tests", "golden")
CLASSES = {0: oh.OrbitKS, 1: oh.RelativeOrbitKS, 2: oh.ShiftReflectionOrbitKS, 3: oh.AntisymmetricOrbitKS}


def put(out, key, r, o_in=None, sub=None):
    st = r.orbit.state if sub is None else r.orbit.state[sub]
    out[key + "/state"] = st
    out[key + "/params"] = np.array(r.orbit.parameters, dtype=float)
    out[key + "/status"] = np.array(r.status)
    out[key + "/nit"] = np.array(r.nit, dtype=float)
    out[key + "/costs"] = np.array(r.costs, dtype=float)
    if o_in is not None:
        out[key + "/in_state"] = o_in.state
        out[key + "/in_params"] = np.array(o_in.parameters, dtype=float)


def gen_cfg2(out):
    """BASELINE configs[2]: 64x64 Relative / ShiftReflection / Antisymmetric, (T, L) = (60, 44), adjoint warm-up then
    Newton-GMRES(50), rtol 1e-6, tol 1e-10 (restart cycles bounded to 4 as in BASELINE.md section 2)."""
    for cid in (1, 2, 3):
        C = CLASSES[cid]
        for seed in (0, 1):
            o = C(parameters=(60.0, 44.0, 0.0)).populate(attr="state", seed=seed, discretization=(64, 64))
            t0 = time.time()
            r1 = oh.hunt(o, "adj", maxiter=2000)
            key = f"cfg2/{cid}_{seed}"
            put(out, key + "/adj", r1, o_in=o)
            r2 = oh.hunt(r1.orbit, "gmres", maxiter=3, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-6})
            put(out, key + "/gmres", r2)
            print(key, "adj", r1.status, r1.nit, r1.costs[-1], "gmres", r2.status, r2.nit, r2.costs, f"{time.time() - t0:.1f}s", flush=True)


def gen_cfg3(out):
    """BASELINE configs[3]: OrbitKS 256x256, (T, L) = (200, 200): one outer Newton step of GMRES(50), one restart cycle."""
    o = oh.OrbitKS(parameters=(200.0, 200.0, 0.0)).populate(attr="state", seed=0, discretization=(256, 256))
    t0 = time.time()
    r = oh.hunt(o, "gmres", maxiter=1, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 1, "rtol": 1e-6})
    sub = (slice(None, None, 7), slice(None, None, 5))
    put(out, "cfg3/0", r, sub=sub)
    out["cfg3/0/seed"] = np.array(0)
    out["cfg3/0/sub"] = np.array([7, 5])
    out["cfg3/0/state_norm"] = np.array(np.linalg.norm(r.orbit.state))
    print("cfg3", r.status, r.nit, r.costs, f"{time.time() - t0:.1f}s", flush=True)


def gen_square(out):
    """Every parameter constrained (continuation's case): gmres runs on the square system J dx = -F, optimize.py:1803-1850."""
    for cid in (0, 2, 1):
        C = CLASSES[cid]
        o = C(parameters=(44.0, 44.0, 0.0)).populate(attr="state", seed=3, discretization=(32, 32))
        o = C(**{**vars(o), "constraints": {"t": True, "x": True, "s": True}})
        for pc in (False, True):
            r = oh.hunt(o, "gmres", maxiter=3, preconditioning=pc, scipy_kwargs={"restart": 20, "maxiter": 2})
            key = f"square/{cid}" + ("_pc" if pc else "")
            put(out, key, r, o_in=o)
            print(key, r.status, r.nit, r.costs, flush=True)


def gen_lsmr(out):
    """lsmr (optimize.py:1319-1321) and lsqr at the same settings, 32x32, all four classes."""
    for cid, C in CLASSES.items():
        o = C(parameters=(44.0, 44.0, 0.0)).populate(attr="state", seed=3, discretization=(32, 32))
        for method, kw in (("lsmr", {"maxiter": 50}), ("lsmr", {}), ("lsqr", {"iter_lim": 300})):
            r = oh.hunt(o, method, maxiter=3, scipy_kwargs=dict(kw))
            key = f"lsq/{cid}/{method}" + ("_" + "_".join(f"{k}{v}" for k, v in kw.items()) if kw else "")
            put(out, key, r, o_in=o)
            print(key, r.status, r.nit, r.costs, flush=True)


def gen_deriv(out):
    """dt / dx (orders 1-4; SR / Anti odd orders go through spatial_modes), precondition and the preconditioner diagonal of
    the unmodified reference (ks/orbits.py:379-524, :3190-3227, :3583-3619, :1034-1141) on a 32x32 and a 16x24 seed."""
    for cid, C in CLASSES.items():
        for (n, m) in ((32, 32), (16, 24)):
            o = C(parameters=(44.0, 33.4, 1.3 if cid == 1 else 0.0)).populate(attr="state", seed=11 + cid, discretization=(n, m))
            o = C(**{**vars(o), "parameters": (44.0, 33.4, 1.3 if cid == 1 else 0.0)})
            key = f"deriv/{cid}_{n}x{m}"
            out[key + "/modes"] = o.state
            out[key + "/params"] = np.array(o.parameters, dtype=float)
            for order in (1, 2, 3, 4):
                out[key + f"/dt{order}"] = o.dt(order=order).state
                out[key + f"/dx{order}"] = o.dx(order=order).state
                out[key + f"/dx{order}_field"] = o.dx(order=order, return_basis="field").state
            out[key + "/dt1_field"] = o.dt(return_basis="field").state
            g = o.rmatvec(o)
            pc = g.precondition(pmult=o.preconditioning_parameters())
            out[key + "/rmatvec_self"] = g.state
            out[key + "/rmatvec_self_params"] = np.array(g.parameters, dtype=float)
            out[key + "/precond"] = pc.state
            out[key + "/precond_params"] = np.array(pc.parameters, dtype=float)
            P = o.preconditioner()
            out[key + "/preconditioner_diag"] = np.asarray(P.matvec(np.ones(P.shape[0]))).ravel()
    print("deriv ok", flush=True)


def _cfg0_one(seed):
    import warnings as w

    w.filterwarnings("ignore")
    o = oh.OrbitKS(parameters=(44.0, 33.4, 0.0)).populate(attr="state", seed=seed, discretization=(32, 32))
    t0 = time.time()
    r1 = oh.hunt(o, "adj", maxiter=10000, tol=1e-6)
    t1 = time.time()
    r = oh.hunt(r1.orbit, "lstsq", maxiter=200, tol=1e-10)
    t2 = time.time()
    # hunt(o, ('adj', 'lstsq'), maxiter=[10000, 200], tol=[1e-6, 1e-10]) is these two calls back to back
    status = r.status
    return [seed, r1.status, r1.nit, r1.costs[-1], status, r.nit, r.costs[-1], r.orbit.t, r.orbit.x, t1 - t0, t2 - t1]


def gen_cfg0(procs, n_seeds=256):
    """BASELINE configs[0] (hunt(('adj','lstsq'), maxiter=[10000,200], tol=[1e-6,1e-10])) on OrbitKS 32x32 seeds 0..n_seeds-1:
    per-seed table [seed, adj status, adj nit, adj cost, final status, lstsq nit, final cost, T, L, adj s, lstsq s].  Rows already
    in tests/golden/ks_cfg0_table.npz are kept (the runs are deterministic), so the table can be extended."""
    import multiprocessing as mp

    path = os.path.join(GOLD, "ks_cfg0_table.npz")
    rows = [list(r) for r in np.load(path)["table"]] if os.path.exists(path) else []
    have = {int(r[0]) for r in rows}
    todo = [s for s in range(n_seeds) if s not in have]
    with mp.get_context("fork").Pool(procs) as pool:
        for row in pool.imap_unordered(_cfg0_one, todo):
            rows.append(row)
            print("cfg0", row, flush=True)
    rows.sort(key=lambda r: r[0])
    np.savez_compressed(os.path.join(GOLD, "ks_cfg0_table.npz"), table=np.array(rows, dtype=float),
                        columns=np.array(["seed", "adj_status", "adj_nit", "adj_cost", "status", "lstsq_nit", "cost", "t", "x",
                                          "adj_seconds", "lstsq_seconds"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg2,cfg3,square,lsq,deriv")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--cfg0-seeds", type=int, default=256)
    args = ap.parse_args()
    which = set(args.only.split(","))
    path = os.path.join(GOLD, "ks_hunt_r2.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    if "cfg2" in which:
        gen_cfg2(out)
    if "cfg3" in which:
        gen_cfg3(out)
    if "square" in which:
        gen_square(out)
    if "lsq" in which:
        gen_lsmr(out)
    if "deriv" in which:
        gen_deriv(out)
    if which & {"cfg2", "cfg3", "square", "lsq", "deriv"}:
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))
    if "cfg0" in which:
        gen_cfg0(args.procs, args.cfg0_seeds)


if __name__ == "__main__":
    main()
