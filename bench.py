"""Benchmark of the KS hot path on B200:  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): a batch of 4096 random OrbitKS 32x32 seeds (reference generator,
T=44, L=33.4).  One "step" = one fused residual+adjoint evaluation (cost = 1/2|F|^2 and grad = J^T F incl. the
parameter components) of every orbit in the batch = 4096 evals.  `value` = evals/s with inputs resident in HBM;
`e2e` = the same through the public BatchedOrbitKS API from pinned HOST buffers (H2D of modes+params and D2H of
cost+gradient inside the timed region).  The timed loop rotates over several distinct input/output sets so the
working set exceeds the 126 MB L2.  Under torchrun every rank evaluates its own 4096 seeds (weak scaling, no
collective on the data path; one all_reduce(MAX) of the elapsed time and one gather of per-rank results).

The K timed steps are captured into ONE CUDA graph (K kernel nodes over the rotating sets, programmatic-dependent-launch
edges kept) and the graph launch is what the events bracket, so the number does not depend on how fast the host can issue
30-microsecond launches.  The same JSON line carries the other BASELINE workloads as side blocks: `adjoint_descent`
(configs[1]: hunt('adj', 10000 steps) on the 4096 seeds), `sweep` (configs[4]: the 65 536-seed sweep strong-scaled over the
ranks, two stages, final NCCL gather; converged orbits/s), `config2` (4096 x 3 symmetry classes at 64x64, adjoint warm-up
+ Newton-GMRES(50)) and `config3` (OrbitKS 256x256: matvec / rmatvec alone and inside GMRES(50)).

`--impl reference` times the CPU arm instead: the UNMODIFIED reference (oracle/_ref, the copy oracle/build_ref.py makes; it
is pure Python) through its own per-orbit API -- orbit.eqn(), .cost(evaleqn=False), orbit.costgrad(F), the body of its
adjoint-descent loop (optimize.py:462-476) -- one process per host core, same workload / metric / unit.  Without
oracle/_ref it falls back to the batched NumPy oracle port (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID, M_GRID, BATCH = 32, 32, 4096
T0, L0 = 44.0, 33.4
NM = (N_GRID - 1) * (M_GRID - 2)
ALG_BYTES_PER_EVAL = 8 * (2 * NM + 7)  # read modes+params, write grad modes+params+cost
METRIC = "orbit residual+adjoint evals/sec (OrbitKS 32x32)"
UNIT = "evals/s"


def make_seeds(first, count):
    """Reference-generator seeds (ks/orbits.py:1738) through the product's own populate()."""
    from orbithunter_b200 import OrbitKS

    out = np.empty((count, N_GRID - 1, M_GRID - 2))
    for i in range(count):
        out[i] = OrbitKS(parameters=(T0, L0, 0.0)).populate(attr="state", seed=first + i, discretization=(N_GRID, M_GRID)).state
    return out


class ClockSampler:
    """SM clock and throttle reasons while the GPU is under the benchmark's load, through NVML (no subprocess: forking
    nvidia-smi from the benchmark process measurably disturbed the 0.5 ms timed region -- 0.69 instead of 0.54 ms on the
    ranks it hit, profiles/r02u_n8_clock_sampler_effect.json).  The first sample is taken `delay` seconds after the start, so
    the sub-millisecond timed region itself is never touched; the same launches are then repeated (untimed) for a second so
    that the samples are taken under exactly that load.  Falls back to nvidia-smi when pynvml is missing."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, delay=0.05):
        self.index, self.samples, self.reasons, self.max_mhz, self.delay = index, [], set(), None, delay
        self.disabled = False
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except (ValueError, IndexError):
                pass
        return index

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
        self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if mask & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for nm, val in zip(names, f[2:]):
            if val.lower().startswith("active"):
                self.reasons.add(nm)

    def _run(self):
        if self._stop.wait(self.delay):
            return
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if not self.disabled:
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if not self.disabled:
            self._th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi",
                "note": "sampled every 50 ms from 50 ms after the start of the timed region on, while the timed launches are repeated "
                        "(the timed region itself lasts well under a millisecond)"}


def cpu_port_evals_per_s(modes, params, budget_s, procs):
    """Oracle port (NumPy restatement of the reference path), one process per core, for about budget_s seconds."""
    import multiprocessing as mp

    if procs == 1:  # in-process: never fork a process that holds a CUDA context
        t0 = time.perf_counter()
        n_done = _cpu_worker(modes, params, budget_s)
        return n_done / (time.perf_counter() - t0), n_done
    chunks = np.array_split(np.arange(modes.shape[0]), procs)
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        reps = pool.starmap(_cpu_worker, [(modes[c], params[c], budget_s) for c in chunks if len(c)])
        dt = time.perf_counter() - t0
    return sum(reps) / dt, sum(reps)


def _cpu_worker(modes, params, budget_s):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import ks_oracle as ko

    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        ko.costgrad(modes, params, ko.ORBIT)
        done += modes.shape[0]
    return done


def _ref_worker(first_seed, count, budget_s):
    """The unmodified reference, one orbit at a time: F = o.eqn(); cost = F.cost(evaleqn=False); grad = o.costgrad(F)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import refshim

    oh = refshim.load()
    orbits = [oh.OrbitKS(parameters=(T0, L0, 0.0)).populate(attr="state", seed=first_seed + i, discretization=(N_GRID, M_GRID))
              for i in range(count)]
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        for o in orbits:
            F = o.eqn()
            F.cost(evaleqn=False)
            o.costgrad(F)
        done += count
    return done, time.perf_counter() - t0


def reference_evals_per_s(sample, budget_s, procs):
    """Unmodified reference on `procs` host cores for about budget_s seconds -> (evals/s, evals done)."""
    import multiprocessing as mp

    per = max(1, sample // procs)
    if procs == 1:
        n, dt = _ref_worker(0, per, budget_s)
        return n / dt, n
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.starmap(_ref_worker, [(i * per, per, budget_s) for i in range(procs)])
    return sum(n / dt for n, dt in res), sum(n for n, _ in res)


def reference_available():
    try:
        from oracle import refshim

        return refshim.available()
    except Exception:
        return False


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    use_ref = reference_available()
    sample = (16 if use_ref else 64) * cores
    per_step_budget = max(1.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    rates, t_all = [], time.perf_counter()
    if use_ref:
        for _ in range(min(args.warmup, 1)):
            reference_evals_per_s(sample, min(per_step_budget, 2.0), cores)
        t_all = time.perf_counter()
        for _ in range(args.steps):
            r, _n = reference_evals_per_s(sample, per_step_budget, cores)
            rates.append(r)
    else:
        modes = make_seeds_cpu(sample)
        params = np.tile([T0, L0, 0.0], (sample, 1))
        for _ in range(args.warmup):
            cpu_port_evals_per_s(modes, params, min(per_step_budget, 2.0), cores)
        t_all = time.perf_counter()
        for _ in range(args.steps):
            r, _n = cpu_port_evals_per_s(modes, params, per_step_budget, cores)
            rates.append(r)
    value = float(np.mean(rates))
    kind = "reference" if use_ref else "port"
    how = ("unmodified reference (oracle/_ref), per-orbit eqn + cost + costgrad" if use_ref else "batched NumPy oracle")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batch of {BATCH} random OrbitKS 32x32 seeds, residual+adjoint evaluation", "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} seeds of the same workload, {how}, one process per core, {per_step_budget:.1f} s per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


def sweep_block(args, rank, world, dev):
    """BASELINE configs[4]: 65 536-seed sweep, strong scaling (seeds / world per rank), two stages, NCCL gather on rank 0."""
    import torch
    import torch.distributed as dist

    from orbithunter_b200.sweep import SWEEP_FIELDS, two_stage_sweep

    two_stage_sweep(64 * world, 16 * world, adj_maxiter=64, dense_maxiter=2)  # warm-up: allocations, cuBLAS handle, module load
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rec, tm = two_stage_sweep(args.sweep_seeds, args.sweep_refine, dense_maxiter=args.sweep_dense_maxiter)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    keys = ("seed_generation_s", "adjoint_s", "dense_newton_s", "preprocess_s", "rank_imbalance_s", "gather_s")
    sums = ("adjoint_steps", "dense_solves", "refined", "seeds")
    t = torch.tensor([wall] + [tm[k] for k in keys], dtype=torch.float64, device=dev)
    c = torch.tensor([tm[k] for k in sums], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    wall = float(t[0])
    stage = {k: float(v) for k, v in zip(keys, t[1:])}
    tot = {k: float(v) for k, v in zip(sums, c)}
    r = rec.cpu().numpy()
    assert r.shape == (args.sweep_seeds, len(SWEEP_FIELDS)) and np.array_equal(r[:, 0], np.arange(args.sweep_seeds))
    refined = r[r[:, 7] >= 0]
    st_all, st_ref = r[:, 1].astype(int), refined[:, 1].astype(int)
    conv = int((st_ref == 1).sum())
    conv_nontrivial = int(((st_ref == 1) & (refined[:, 7] == 0)).sum())
    stage2 = max(stage["dense_newton_s"] + stage["preprocess_s"], 1e-9)
    frac = conv / max(len(refined), 1)
    return {
        "workload": f"{args.sweep_seeds} OrbitKS 32x32 seeds (T=44, L=33.4), {args.sweep_seeds // world} per GPU: device seed generation -> "
                    f"hunt('adj', 10000 steps, tol 1e-6) on every seed -> hunt('lstsq', maxiter {args.sweep_dense_maxiter}, tol 1e-10) on a uniform "
                    f"sample of {int(tot['refined'])} seeds ({int(tot['refined']) // world} per GPU) -> preprocess -> gather of 8 doubles per seed (NCCL)",
        "scaling": "strong", "seeds": args.sweep_seeds, "refined_seeds": int(tot["refined"]), "wall_s": wall, "stage_s_max_over_ranks": stage,
        "seeds_per_s": args.sweep_seeds / wall, "adjoint_steps_per_s": tot["adjoint_steps"] / max(stage["adjoint_s"], 1e-9),
        "dense_solves": int(tot["dense_solves"]), "dense_solves_per_s": tot["dense_solves"] / max(stage["dense_newton_s"], 1e-9),
        "status_counts_all_seeds": {str(k): int((st_all == k).sum()) for k in (-1, 0, 1, 2, 3)},
        "status_counts_refined": {str(k): int((st_ref == k).sum()) for k in (-1, 0, 1, 2, 3)},
        "converged": conv, "converged_not_equilibrium": conv_nontrivial, "converged_fraction_of_refined": frac,
        "converged_orbits_per_s": conv / stage2,
        "converged_orbits_per_s_whole_sweep": conv / wall,
        "full_two_stage_sweep_estimate": {
            "note": "all seeds through both stages at the measured per-seed cost of each stage (stage 2 is linear in the seeds refined)",
            "wall_s": stage["seed_generation_s"] + stage["adjoint_s"] + stage2 * args.sweep_seeds / max(tot["refined"], 1) + stage["gather_s"],
            "converged": frac * args.sweep_seeds},
        "gather_ms": 1e3 * stage["gather_s"], "gather_bytes": int(r.size * 8),
        "limiting_phase": max(stage, key=stage.get),
        "reference_cpu": "tests/golden/ks_cfg0_table.npz: the unmodified reference needs 22-62 s per seed on one core for the same two "
                         "stages (columns adj_seconds, lstsq_seconds; 9 of 64 seeds converge)",
    }


def config2_block(peak_gbs):
    """BASELINE configs[2]: 4096 each of Relative / ShiftReflection / Antisymmetric 64x64, (T, L) = (60, 44): 2000 adjoint steps,
    then Newton-GMRES (restart 50, 4 cycles, rtol 1e-6, 3 outer steps, tol 1e-10) -- the settings of the reference golden
    (tests/golden/ks_hunt_r2.npz cfg2/*) at the full batch."""
    import torch

    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    out, B = [], 4096
    for cls in ("RelativeOrbitKS", "ShiftReflectionOrbitKS", "AntisymmetricOrbitKS"):
        u = BatchedOrbitKS.from_seeds(np.arange(B), (60.0, 44.0, 0.0), cls, (64, 64))
        hunt(u, ("adj", "gmres"), maxiter=[4, 1], scipy_kwargs=[{}, {"restart": 5, "maxiter": 1}])  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r1 = hunt(u, "adj", maxiter=2000)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        r2 = hunt(r1.orbit, "gmres", maxiter=3, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 4, "rtol": 1e-6})
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        N = u.state.shape[1] * u.state.shape[2] + len(u.free_indices())
        steps = int(r2.solver_steps)  # lock-step Arnoldi steps of the batch (every orbit takes part in each)
        # Arnoldi step j moves 8 N (2 (j + 1) + 1) bytes per orbit; a full restart-50 cycle 8 N 2600
        arnoldi_bytes = 8.0 * N * 2600.0 * steps / 50.0 * B
        st = r2.status.cpu().numpy()
        out.append({"class": cls, "grid": [64, 64], "batch": B, "adjoint_2000_steps_s": t1 - t0,
                    "adjoint_steps_per_s": float(r1.nit.sum().item()) / (t1 - t0),
                    "newton_gmres50_s": t2 - t1, "outer_iterations": int(r2.outer_iterations), "arnoldi_steps": steps,
                    "operator_applications": int(r2.operator_applications),
                    "arnoldi_algorithmic_GBps": arnoldi_bytes / (t2 - t1) / 1e9, "arnoldi_hbm_frac": arnoldi_bytes / (t2 - t1) / 1e9 / peak_gbs,
                    "median_cost_after_adjoint": float(r1.costs[-1].median().item()), "median_cost_after_gmres": float(r2.costs[-1].median().item()),
                    "status_counts": {str(k): int((st == k).sum()) for k in (-1, 0, 1, 2, 3)}})
    return out


def config3_block(peak_gbs):
    """BASELINE configs[3]: OrbitKS 256x256, (T, L) = (200, 200), seeds 0..63: matvec and rmatvec alone (CUDA events, SURVEY 8(d)
    bytes 8 (3 Nm + 9) per product) and one Newton step of GMRES(50) (one restart cycle, rtol 1e-6)."""
    import torch

    from orbithunter_b200 import BatchedOrbitKS
    from orbithunter_b200.optimize import hunt

    B = 64
    u = BatchedOrbitKS.from_seeds(np.arange(B), (200.0, 200.0, 0.0), "OrbitKS", (256, 256))
    v = BatchedOrbitKS.from_seeds(np.arange(B) + 1000, (200.0, 200.0, 0.0), "OrbitKS", (256, 256))
    nm = u.state.shape[1] * u.state.shape[2]
    out = {"class": "OrbitKS", "grid": [256, 256], "batch": B}
    for name, fn in (("matvec", lambda: u.matvec(v)), ("rmatvec", lambda: u.rmatvec(v)), ("costgrad", lambda: u.costgrad())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = B * 8 * ((2 * nm + 7) if name == "costgrad" else (3 * nm + 9))
        out[name] = {"per_s": B / ms * 1e3, "ms_per_batch": ms, "algorithmic_GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / peak_gbs}
    hunt(u, "gmres", maxiter=1, scipy_kwargs={"restart": 5, "rtol": 1e-2, "maxiter": 1})  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = hunt(u, "gmres", maxiter=1, tol=1e-10, scipy_kwargs={"restart": 50, "maxiter": 1, "rtol": 1e-6})
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N, steps = nm + 2, int(r.solver_steps)
    arnoldi_bytes = 8.0 * N * 2600.0 * steps / 50.0 * B
    out["newton_gmres50"] = {"wall_s": dt, "arnoldi_steps": steps, "operator_applications": int(r.operator_applications),
                             "ms_per_arnoldi_step_batch": 1e3 * dt / max(steps, 1), "arnoldi_algorithmic_GBps": arnoldi_bytes / dt / 1e9,
                             "arnoldi_hbm_frac": arnoldi_bytes / dt / 1e9 / peak_gbs,
                             "median_cost_before": float(r.costs[0].median().item()), "median_cost_after": float(r.costs[-1].median().item())}
    return out


def side_shapes(peak_gbs):
    """Fused residual+adjoint evals/s at the other BASELINE shapes; side measurements, CUDA-event timed, inputs resident."""
    import torch

    from orbithunter_b200 import BatchedOrbitKS

    out = []
    for cls, n, m, T, L, B in (("OrbitKS", 32, 32, T0, L0, 65536),  # configs[4] batch: 55 full rounds of the persistent warps
                               ("RelativeOrbitKS", 64, 64, 60.0, 44.0, 4096), ("ShiftReflectionOrbitKS", 64, 64, 60.0, 44.0, 4096),
                               ("AntisymmetricOrbitKS", 64, 64, 60.0, 44.0, 4096)):
        u = BatchedOrbitKS.from_seeds(np.arange(B), (T, L, 0.0), cls, (n, m))
        for _ in range(3):
            u.costgrad()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            u.costgrad()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        nm = u.state.shape[1] * u.state.shape[2]
        gbs = B * 8 * (2 * nm + 7) / ms / 1e6
        out.append({"class": cls, "grid": [n, m], "batch": B, "evals_per_s": B / ms * 1e3, "algorithmic_GBps": gbs,
                    "hbm_frac": gbs / peak_gbs,
                    "kernels": ("ks_warp32x1.cuh: one warp per orbit (the headline kernel at the 65 536-seed sweep's batch)"
                                if (n, m) == (32, 32) else "ks_tile64.cuh: one CTA per orbit, all passes in shared memory")})
    return out


def make_seeds_cpu(count):
    from oracle import ks_oracle as ko

    return np.stack([ko.populate_state(T0, L0, N_GRID, M_GRID, s, ko.ORBIT) for s in range(count)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--sets", type=int, default=5, help="distinct input/output sets rotated to exceed L2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hunt", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time K plain launches instead of one K-node CUDA graph")
    ap.add_argument("--e2e-streams", type=int, default=2, help="streams (each with its own pinned buffers) of the e2e loop")
    ap.add_argument("--no-side", action="store_true", help="skip the side blocks (sweep, config2, config3, other shapes)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--sweep-seeds", type=int, default=65536)
    ap.add_argument("--sweep-refine", type=int, default=4096, help="seeds (over all ranks) that also get the dense second stage")
    ap.add_argument("--no-clock-sampler", action="store_true", help="debug: do not poll nvidia-smi during the timed region")
    ap.add_argument("--sweep-dense-maxiter", type=int, default=200)
    ap.add_argument("--hunt-maxiter", type=int, default=10000)
    ap.add_argument("--variant", type=int, default=None, help="tuning: warp32 kernel variant (ks_debug_warp32_variant)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from orbithunter_b200 import BatchedOrbitKS, _lib

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if args.variant is not None:
        L.ks_debug_warp32_variant(args.variant)

    # ---- inputs: this rank's 4096 seeds, replicated into `sets` perturbed copies (distinct memory, same statistics) ----
    base = make_seeds(rank * BATCH, BATCH)
    params_h = np.tile([T0, L0, 0.0], (BATCH, 1))
    sets = []
    for s in range(args.sets):
        orb = BatchedOrbitKS(torch.as_tensor(base * (1.0 + 1e-3 * s)), torch.as_tensor(params_h), "OrbitKS", "modes")
        sets.append({"orb": orb, "cost": torch.empty(BATCH, dtype=torch.float64, device=dev),
                     "g": torch.empty_like(orb.state), "gp": torch.empty_like(orb.parameters)})
    ws_bytes = L.ks_workspace_bytes(0, N_GRID, M_GRID, BATCH)
    ws = torch.empty(max(ws_bytes, 1024), dtype=torch.uint8, device=dev)

    def step(i, stream=None):
        s = sets[i % len(sets)]
        o = s["orb"]
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(L.ks_costgrad(0, N_GRID, M_GRID, BATCH, o.state.data_ptr(), o.parameters.data_ptr(), o.cmask, 0,
                                 s["cost"].data_ptr(), s["g"].data_ptr(), s["gp"].data_ptr(), None, ws.data_ptr(), ws_bytes, st))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- W warm-up steps, then EXACTLY K timed steps.  The K launches are captured into one CUDA graph on a side stream (the
    # warm-up runs there too, so attribute opt-ins and module load are done before the capture); events bracket its launch.
    side = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(side):
        for i in range(args.warmup):
            step(i)
        side.synchronize()
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for i in range(args.steps):
                    step(args.warmup + i)
            graph.replay()  # untimed: uploads the graph
            side.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    sampler.disabled = args.no_clock_sampler
    with sampler as clocks:
        with torch.cuda.stream(side):
            ev0.record()
            if graph is not None:
                graph.replay()
            else:
                for i in range(args.steps):
                    step(args.warmup + i)
            ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        # keep the GPU busy long enough for the 10 Hz clock sampler to see it under load: repeat the graph (untimed)
        with torch.cuda.stream(side):
            t_busy = time.perf_counter()
            while time.perf_counter() - t_busy < 1.0:
                if graph is not None:
                    graph.replay()
                else:
                    step(0)
                side.synchronize()
        # the same step timed one launch at a time (best and median of >= 10 repetitions); the events between the
        # launches keep consecutive kernels from overlapping their launch latency, so these are upper bounds of the mean above
        n_single = max(10, min(args.steps, 50))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_single)]
        for i, (a0, a1) in enumerate(evs):
            a0.record()
            step(i)
            a1.record()
        torch.cuda.synchronize()
        single = sorted(a0.elapsed_time(a1) for a0, a1 in evs)
        # ---- end-to-end through the public API from pinned host buffers: every step uploads its inputs and reads its
        # results back; two streams with their own pinned buffers so step i+1's H2D overlaps step i's D2H (PCIe is duplex)
        streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.e2e_streams))]
        hbuf = []
        for _ in streams:
            hbuf.append({"m": torch.as_tensor(base).pin_memory(), "p": torch.as_tensor(params_h).pin_memory(),
                         "g": torch.empty((BATCH, N_GRID - 1, M_GRID - 2), dtype=torch.float64).pin_memory(),
                         "gp": torch.empty((BATCH, 3), dtype=torch.float64).pin_memory(),
                         "c": torch.empty(BATCH, dtype=torch.float64).pin_memory()})

        def e2e_step(i):
            sidx = i % len(streams)
            h = hbuf[sidx]
            with torch.cuda.stream(streams[sidx]):
                orb = BatchedOrbitKS(h["m"].to(dev, non_blocking=True), h["p"].to(dev, non_blocking=True), "OrbitKS", "modes")
                cost, grad = orb.costgrad()
                h["g"].copy_(grad.state, non_blocking=True)
                h["gp"].copy_(grad.parameters, non_blocking=True)
                h["c"].copy_(cost, non_blocking=True)

        e2e_steps = max(4, min(args.steps, 60))
        for i in range(4):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(e0)
        for i in range(e2e_steps):
            e2e_step(i)
        for st_ in streams[1:]:
            streams[0].wait_stream(st_)
        e1.record(streams[0])
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        assert abs(float(hbuf[0]["c"].sum()) - float(sets[0]["cost"].sum())) <= 1e-9 * abs(float(sets[0]["cost"].sum()))
        hunt_stats = None
        if not args.no_hunt:
            # BASELINE configs[1] proper: hunt('adj') on the same 4096 seeds (reference defaults tol 1e-6, ftol 1e-9, min_step 1e-6)
            from orbithunter_b200.optimize import hunt

            orb = BatchedOrbitKS(torch.as_tensor(base), torch.as_tensor(params_h), "OrbitKS", "modes")
            hunt(orb, "adj", maxiter=8)  # warm-up: first-call workspace allocation and module load stay outside the timing
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = hunt(orb, "adj", maxiter=args.hunt_maxiter)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            st = res.status.cpu().numpy()
            nit_total = int(res.nit.sum().item())
            hunt_stats = {
                "orbits": BATCH, "maxiter": args.hunt_maxiter, "wall_s": wall, "descent_steps_per_s": nit_total / wall,
                "status_counts": {str(k): int((st == k).sum()) for k in (-1, 0, 1, 2, 3)},
                "median_final_cost": float(res.costs[-1].median().item()),
                "note": "rank 0's 4096 seeds (BASELINE configs[1]); multi-pass kernel: each orbit runs up to 128 passes per launch with "
                        "its point and gradient in shared memory (wall time of the whole hunt() call: initial cost evaluation, host "
                        "polls, final state); adjoint descent alone converges none of these seeds in 10 000 steps, in the reference "
                        "as here (tests/golden/ks_cfg0_table.npz) -- converged orbits/s is the sweep block's number",
            }
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    per_rank_ms = [ms]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x[0]) for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    checksum = float(sets[0]["cost"].sum().item())
    sweep = None
    if not args.no_side and not args.no_sweep:
        sweep = sweep_block(args, rank, world, dev)  # every rank takes part (strong scaling + the NCCL gather)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_per_step = ms / args.steps
    value = world * BATCH / (ms_per_step * 1e-3)
    e2e_value = world * BATCH / (e2e_ms / e2e_steps * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = BATCH * ALG_BYTES_PER_EVAL / (ms_per_step * 1e-3) / 1e9  # per GPU: one launch = 4096 evals
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("costgrad_32x32_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"batch of {BATCH} random OrbitKS 32x32 seeds per GPU (T=44, L=33.4, reference generator), "
                               "fused residual+adjoint (cost + J^T F) evaluation; BASELINE configs[1] shape",
                   "batch_per_gpu": BATCH, "grid": [N_GRID, M_GRID], "class": "OrbitKS",
                   "l2": f"rotating {len(sets)} input/output sets ({len(sets) * BATCH * ALG_BYTES_PER_EVAL / 1e6:.0f} MB) > 126 MB L2",
                   "launch": ("the K steps are K kernel nodes of one CUDA graph; CUDA events bracket the graph launch" if graph is not None
                              else "K plain launches"),
                   "parallelism": f"orbits sharded over {world} GPU(s), no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "ks_warp32x1_kernel<OrbitKS,COSTGRAD>",
                     "algorithmic_bytes_per_launch": BATCH * ALG_BYTES_PER_EVAL,
                     "note": "not HBM-bound: the kernel executes 1704 fp64 warp-instructions per eval (ncu; 54.5 k lane-ops), so at the "
                             "MEASURED DFMA peak of 63.5 lanes/clk/SM (profiles/r02a_fp64_peak.jsonl) 100 % of the fp64 pipe is 339 M "
                             "evals/s = 0.78 of this HBM roofline; measured pipe activity 51 %, the limiter is dependent-issue latency at "
                             "two warps per scheduler (DESIGN.md section 4); 4096 orbits are 3.46 rounds of the 1184 resident warps, "
                             "B = 65536 reaches 0.45-0.46 (other_shapes[0])"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * (NM + 3) * 8,
                "d2h_bytes_per_step": BATCH * (NM + 3 + 1) * 8, "steps": e2e_steps},
        "gpu_launches": args.steps, "clocks": clocks.summary(), "checksum_cost_sum": checksum,
        "single_launch_ms": {"best": single[0], "median": single[len(single) // 2], "n": len(single)},
        "timed_region_ms_per_rank": per_rank_ms,
    }
    if hunt_stats is not None:
        line["adjoint_descent"] = hunt_stats
    if sweep is not None:
        line["sweep"] = sweep
    if not args.no_side:
        line["config2"] = config2_block(peak)
        line["config3"] = config3_block(peak)
        line["other_shapes"] = side_shapes(peak)
    if not args.no_cpu_baseline and world == 1:  # the CPU leg runs on rank 0 at N = 1 only
        if reference_available():
            rate, n_done = reference_evals_per_s(16, 10.0, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"16 seeds of the same workload, unmodified reference (oracle/_ref) per-orbit eqn + cost + "
                                              f"costgrad on one core for 10 s ({n_done} evals)"}
        else:
            cm = make_seeds_cpu(256)
            rate, n_done = cpu_port_evals_per_s(cm, np.tile([T0, L0, 0.0], (256, 1)), 10.0, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"256 seeds of the same workload, batched NumPy oracle on one core for 10 s ({n_done} evals)"}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
