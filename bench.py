#!/usr/bin/env python
"""Benchmark of the batched-iLQR hot path (BASELINE.json metric: iLQR trajectory-iterations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm (this repo)
    python bench.py --impl reference [...]                         # reference arm: CPU solver on host cores

Workload (config.workload): BASELINE.json configs[1] - ThreeLinkPlanarManipulator, 65,536 random
initial states/targets per GPU (generator of SURVEY.md section 8(d), seed 1234 + rank), T = 100, FP64,
`BasiciLQR(max_line_search=10)` defaults, solved to convergence.  One *step* = one whole-batch solve
(one persistent-kernel launch).  A trajectory-iteration = one backward pass + one line-search
forward pass of one trajectory; the number executed is data dependent and is summed from the
solver's own per-trajectory iteration counts.

value   device-timed throughput, inputs resident in HBM (CUDA events on the launch stream).
e2e     same metric through the public API (`BasiciLQR.solve_batch`) from pinned HOST buffers:
        H2D of initial states/targets and D2H of costs/iteration counts/flags/trajectories inside
        the timed region (the trajectories are written by the kernel straight into pinned host
        memory as they finish, `trajectory_to_host=True`, so that copy overlaps the solve).
roofline  achieved algorithmic HBM bytes/s of the solve kernel vs the measured copy bandwidth.
also    the other BASELINE configs measured in the SAME run (device-timed, fewer steps): ThreeLink
        1,048,576 trajectories on one GPU to convergence (north_star's configuration), the headline batch
        in fixed-work mode (is_check_stop=False, max_iter=20), RoboticArm 262,144 x T=200 FP64
        (config 3); under --gpus 8 the config-4 shard size (131,072 per GPU).  `--no-also` skips them.
cpu_baseline / --impl reference   the UNMODIFIED reference through oracle/ref_runner.py when it is installed
        on the box (baseline/_ref/, oracle/vendor_reference.sh; "kind": "reference", as shipped and numerics
        only), else the bit-checked port oracle/ilqr_numpy.py ("kind": "port"); host cores, N=1 / rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = "ThreeLinkPlanarManipulator"
METRIC = "ilqr_trajectory_iterations_per_sec"
UNIT = "traj-iter/s"
MAX_LINE_SEARCH = 10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="trajectories per GPU")
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--solver", default="basic", choices=["basic", "logbarrier"],
                    help="basic = BasiciLQR (the headline); logbarrier = LogBarrieriLQR (SURVEY 8(f) N1): a step is one "
                         "solve_batch = one inner iLQR loop per barrier parameter t in [0.5 .. 100]")
    ap.add_argument("--mode", default="converge", choices=["converge", "fixed"],
                    help="converge: reference defaults; fixed: is_check_stop=False, max_iter=20")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="trajectories per core for the CPU baseline (0 = auto)")
    ap.add_argument("--grid-blocks", type=int, default=0)
    ap.add_argument("--no-also", action="store_true", help="skip the `also` block (the other BASELINE configs)")
    ap.add_argument("--e2e-output", default="auto", choices=["auto", "direct", "copy"],
                    help="end-to-end leg: the kernel writes finished trajectories straight into pinned host memory "
                         "(direct) or the result is copied device-to-host after the solve (copy); auto = whichever "
                         "measured faster for this number of GPUs (copy on one GPU, direct on several)")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the unmodified reference (needs baseline/_ref/), the oracle port, or whichever is there")
    return ap.parse_args()


def bytes_per_traj_iter(n, m, T, s):
    """compulsory HBM traffic of one trajectory-iteration, s*T*(3d + 2m(n+1))."""
    d = n + m
    return s * T * (3 * d + 2 * m * (n + 1))


def profiled_traffic(model, dtype, batch, mode):
    """DRAM bytes (read + write) of one solve-kernel launch from the committed `ncu --set full` capture
    of this workload (profiles/*/traffic.json, written by profiles/summarize_ncu.py), or None."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json"))):
        try:
            with open(path) as fh:
                for rec in json.load(fh):
                    if (rec.get("model"), rec.get("dtype"), rec.get("batch"), rec.get("mode")) == (model, dtype, batch, mode):
                        best = rec
        except Exception:
            pass
    return best


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r.split(", ") for ts, r in self.samples if t0 <= ts <= t1 + 0.2] or [r.split(", ") for _, r in self.samples]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) >= 6 and r[2 + i].strip() == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows else None,
                "reasons": reasons, "samples": len(rows)}


def bench_inputs(model, batch, seed, T):
    """(x0 [B, n], targets [B, 2] or None): the arm scenarios draw random states and targets; the
    scenarios with a generated device model (cart-pole, vehicle: no target) perturb their own initial state."""
    from curiousrl_b200.utils.synthetic import ARM_GEOMETRY, synthetic_batch, synthetic_init_states
    if model in ARM_GEOMETRY:
        return synthetic_batch(model, batch, seed=seed)
    import curiousrl_b200.scenario.dynamic_model as models
    return synthetic_init_states(getattr(models, model)(T=T), batch, seed=seed), None


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
_WORKER = {}


def _cpu_init(model, T, barrier, solver="basic"):
    """Per-process set-up: build the oracle problem and pay the numba JIT once (untimed, as the
    reference itself excludes its compile iteration, basic_ilqr.py:82-83)."""
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "NUMBA_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import warnings
    warnings.filterwarnings("ignore")
    import curiousrl_b200.scenario.dynamic_model as models
    from oracle import ilqr_numpy as orc                 # CPU baseline / reference arm: the one place bench.py runs the oracle
    sc = getattr(models, model)(T=T)
    prob = orc.Problem(sc, use_numba=True)
    x0, tgt = bench_inputs(model, 1, 99, T)
    tgt = [None] if tgt is None else tgt
    orc.solve(prob, x0[0], tgt[0], max_line_search=MAX_LINE_SEARCH, max_iter=3)
    _WORKER["prob"], _WORKER["orc"], _WORKER["barrier"], _WORKER["prob_barrier"] = prob, orc, barrier, None
    if solver == "logbarrier":
        pb = orc.Problem(orc.BarrierScenario(sc, 0.5), use_numba=True)
        orc.solve_log_barrier(pb, prob, x0[0], tgt[0], max_line_search=MAX_LINE_SEARCH, max_iter=2)
        _WORKER["prob_barrier"] = pb


def _cpu_solve(job):
    x0, tgt = job
    orc, prob = _WORKER["orc"], _WORKER["prob"]
    t0 = time.perf_counter()
    iters = 0
    pb = _WORKER["prob_barrier"]
    for b in range(x0.shape[0]):
        if pb is None:
            iters += orc.solve(prob, x0[b], None if tgt is None else tgt[b], max_line_search=MAX_LINE_SEARCH)["iters"]
        else:
            iters += int(orc.solve_log_barrier(pb, prob, x0[b], None if tgt is None else tgt[b],
                                               max_line_search=MAX_LINE_SEARCH)["inner_iters"].sum())
    return iters, time.perf_counter() - t0


class CpuReference:
    """Pool of one worker per host core, each running the CPU oracle port single-threaded."""

    def __init__(self, model, T, cores=None, solver="basic"):
        import multiprocessing as mp
        self.model, self.T = model, T
        self.cores = cores or min(os.cpu_count() or 1, 64)
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(model, T, ctx.Barrier(self.cores), solver))
        self.pool.map(_ready, range(self.cores), chunksize=1)   # returns once every worker finished its JIT warm-up

    def run(self, per_core, seed):
        """Solve the first per_core*cores trajectories of the seeded batch; returns (iters, seconds)."""
        x0, tgt = bench_inputs(self.model, per_core * self.cores, seed, self.T)
        jobs = [(x0[c::self.cores], None if tgt is None else tgt[c::self.cores]) for c in range(self.cores)]
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_solve, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def _ready(_):
    _WORKER["barrier"].wait(timeout=600)      # one task per worker: nobody passes until all are initialised
    return 0


# ----------------------------------------------------------------------------- CPU arm: reference or port
def cpu_arm(args):
    """-> (kind, pool-like object with .cores / .sample(per_core, seed, as_shipped) / .close())."""
    from oracle import ref_shim                            # CPU baseline / reference arm only (see module docstring)
    want = args.cpu_kind
    if want != "port" and ref_shim.vendored():
        from oracle import ref_runner

        class Ref:
            kind = "reference"

            def __init__(self):
                self.pool = ref_runner.ReferencePool(ref_shim.VENDORED_ROOT, args.model, args.horizon, solver=args.solver)
                self.cores = self.pool.cores

            def sample(self, per_core, seed, as_shipped=True):
                x0, tgt = bench_inputs(args.model, per_core * self.cores, seed, args.horizon)
                return self.pool.run(x0, tgt, as_shipped)

            def close(self):
                self.pool.close()

        return Ref()
    if want == "reference":
        raise SystemExit("bench.py: --cpu-kind reference needs baseline/_ref/ (bash oracle/vendor_reference.sh)")

    class Port:
        kind = "port"

        def __init__(self):
            self.pool = CpuReference(args.model, args.horizon, solver=args.solver)
            self.cores = self.pool.cores

        def sample(self, per_core, seed, as_shipped=True):
            return self.pool.run(per_core, seed)

        def close(self):
            self.pool.close()

    return Port()


def cpu_description(arm, per_core, iters, wall):
    what = ("the UNMODIFIED reference (baseline/_ref, through the import shim oracle/ref_shim.py), BasiciLQR.solve() as shipped "
            "- including its per-iteration json.dumps of the trajectory (basic_ilqr.py:93) -" if arm.kind == "reference" else
            "CPU oracle port (oracle/ilqr_numpy.py, bit-checked against the reference) on the reference's numba+BLAS substrate,")
    return ("first %d trajectories of the same seeded batch (%d per core; %d traj-iters in %.1f s), %s one single-threaded "
            "worker process per core, numba JIT warm-up excluded" % (per_core * arm.cores, per_core, iters, wall, what))


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_core = args.cpu_sample or 64
    arm = cpu_arm(args)
    iters_tot, wall_tot = 0, 0.0
    for step in range(args.warmup + args.steps):
        iters, wall = arm.sample(per_core, seed=1234)
        if step >= args.warmup:
            iters_tot += iters
            wall_tot += wall
    value = iters_tot / wall_tot
    base = {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
            "sample": "each step = " + cpu_description(arm, per_core, iters_tot // max(1, args.steps), wall_tot / max(1, args.steps))}
    if arm.kind == "reference":                       # both flavours
        it2, w2 = arm.sample(per_core, seed=1234, as_shipped=False)
        base["numerics_only_value"] = it2 / w2
        base["note"] = "value = as shipped; numerics_only_value = the same with logger.save_to_json replaced by a no-op"
    arm.close()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_tot / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    return {"workload": "%s iLQR, batch %d random initial states per GPU, T=%d, %s, %s(max_line_search=%d), %s" % (
                args.model, args.batch, args.horizon, "FP64" if args.dtype == "float64" else "FP32",
                "BasiciLQR" if args.solver == "basic" else "LogBarrieriLQR", MAX_LINE_SEARCH,
                "to convergence (relative 1e-6, max_iter 1000)" if args.mode == "converge" else "fixed work (20 iterations)"),
            "solver": args.solver,
            "model": args.model, "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus, "horizon": args.horizon,
            "mode": args.mode, "parallelism": "dp%d (independent shards, terminal gather of J/iters/flags)" % n_gpus,
            "cache": "working set (solver workspace ~2.1 GB + 0.6 GB results) exceeds the 126 MB L2; "
                     "a 512 MB buffer is overwritten between timed steps"}


def measure_nn_jacobian(dev, trajectories=4096, horizon=100):
    """SURVEY 8(f) N4, the one tensor-core kernel of the product: value + Jacobian of the learned dynamics
    (`LargeNetwork`, random weights in eval mode) at trajectories x horizon rows in one crl_mlp_jacobian call."""
    import torch
    from curiousrl_b200.algorithm.ilqr_solver.nn_ilqr import FoldedMLP, LargeNetwork
    n, d = 8, 11
    torch.manual_seed(0)
    f = FoldedMLP(LargeNetwork(d, n).to(dev).eval(), dev)
    S = trajectories * horizon
    x = torch.randn(S, d, device=dev)
    f.jacobian(x)
    torch.cuda.synchronize()
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        f.jacobian(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = 2.0 * S * (d * f.h1 + f.h1 * f.h2 + n * f.h2 + n * f.h2 * f.h1 + n * f.h1 * d)
    peak = 1100.0                                          # TFLOP/s: dense TF32 tensor peak of a B200 (B200_PROFILING.md)
    return {"workload": "NNiLQR learned-dynamics Jacobian (nn_ilqr.py:265-274): LargeNetwork 11-800-400-8, %d trajectories x "
                        "T=%d = %d rows per call, tcgen05 kind::tf32" % (trajectories, horizon, S),
            "model": "LargeNetwork", "rows": S, "dtype": "tf32 in / f32 accumulate", "steps": steps, "warmup": 1, "ms_per_step": ms,
            "value": S / (ms * 1e-3), "unit": "rows/s",
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peak, "peak_source": "nominal dense TF32 (B200_PROFILING.md)"},
            "timing": "CUDA events on the launch stream"}


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR
    from curiousrl_b200.distributed import gather_results
    import curiousrl_b200.scenario.dynamic_model as models

    logger.set_level(45)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, T = args.batch, args.horizon
    kw = dict(max_line_search=MAX_LINE_SEARCH, dtype=args.dtype, grid_blocks=args.grid_blocks)
    if args.mode == "fixed":
        kw.update(is_check_stop=False, max_iter=20)
    if args.solver == "logbarrier":
        if args.mode != "converge" or args.dtype != "float64":
            raise SystemExit("bench.py: --solver logbarrier runs to convergence in FP64")
        algo = LogBarrieriLQR(max_line_search=MAX_LINE_SEARCH, grid_blocks=args.grid_blocks).init(getattr(models, args.model)(T=T))
    else:
        algo = BasiciLQR(**kw).init(getattr(models, args.model)(T=T))
    dm = algo._dev
    x0_h, tgt_h = bench_inputs(args.model, B, 1234 + rank, T)
    np_dt = np.float64 if args.dtype == "float64" else np.float32
    x0_pin = torch.from_numpy(x0_h.astype(np_dt)).pin_memory()
    tgt_pin = None if tgt_h is None else torch.from_numpy(tgt_h.astype(np_dt)).pin_memory()     # None: the scenario's own add_param
    x0_d, tgt_d = x0_pin.to(dev), (None if tgt_pin is None else tgt_pin.to(dev))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = None
    total_B = B * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        nonlocal out
        out = algo.solve_batch(x0_d, tgt_d, out=out) if args.solver == "basic" else algo.solve_batch(x0_d, tgt_d)
        if world > 1:
            return gather_results(out.obj_fun_value, out.iterations, out.converged, total_B)
        return out

    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    launches0 = dm.launches + (algo._real_dev.launches if args.solver == "logbarrier" else 0)
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        res = step_device()
        ends[i].record()
    launches_timed = dm.launches + (algo._real_dev.launches if args.solver == "logbarrier" else 0) - launches0
    barrier()
    t_wall1 = time.time()
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    iters_local = int(out.iterations.sum().item())                      # identical every step (deterministic)
    conv_local = int(out.converged.sum().item())
    stats = torch.tensor([ms, float(iters_local), float(conv_local)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, iters_all, conv_all = float(mx[0]), float(sm[1]), float(sm[2])
    else:
        iters_all, conv_all = float(iters_local), float(conv_local)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_per_step = ms / args.steps
    value = iters_all / (ms_per_step * 1e-3)

    # ---- end to end through the public API from pinned host memory
    J_h = torch.empty(B, dtype=torch.float64).pin_memory()
    it_h = torch.empty(B, dtype=torch.int32).pin_memory()
    cv_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    X_h = torch.empty((B, T, dm.d), dtype=dm.dtype).pin_memory()
    out_h = None
    direct = args.e2e_output == "direct" or (args.e2e_output == "auto" and world > 1)
    if args.solver == "basic" and direct:   # the kernel writes finished trajectories straight into X_h (mapped pinned host memory)
        from curiousrl_b200.algorithm.ilqr_solver.device import BatchResult
        out_h = BatchResult(X_h, torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                            torch.empty(B, dtype=torch.uint8, device=dev))

    def step_e2e():
        tgt_in = None if tgt_pin is None else tgt_pin.to(dev, non_blocking=True)
        if args.solver == "basic" and direct:
            r = algo.solve_batch(x0_pin.to(dev, non_blocking=True), tgt_in, out=out_h, trajectory_to_host=True)
        elif args.solver == "basic":
            r = algo.solve_batch(x0_pin.to(dev, non_blocking=True), tgt_in, out=out)
        else:
            r = algo.solve_batch(x0_pin.to(dev, non_blocking=True), tgt_in)
        if world > 1:
            gather_results(r.obj_fun_value, r.iterations, r.converged, total_B)
        J_h.copy_(r.obj_fun_value, non_blocking=True)
        it_h.copy_(r.iterations, non_blocking=True)
        cv_h.copy_(r.converged, non_blocking=True)
        if r.trajectory is not X_h:
            X_h.copy_(r.trajectory, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = iters_all / float(e2e_t[0])
    h2d = x0_pin.numel() * x0_pin.element_size() + (0 if tgt_pin is None else tgt_pin.numel() * tgt_pin.element_size())
    d2h = sum(t.numel() * t.element_size() for t in (J_h, it_h, cv_h, X_h))

    # ---- the other BASELINE configs, device-timed in the same run (every rank takes part; max over ranks)
    also = []
    if not args.no_also and args.solver == "basic" and args.model == MODEL and args.mode == "converge" and B == 65536 \
            and args.dtype == "float64":
        flush = None                                  # free the 512 MB flush buffer
        peak_also = measured_hbm_peak()[0]

        def measure(name, model, batch, horizon, steps, **solver_kw):
            """One more workload on this rank's GPU: 1 warm-up + `steps` timed solves, CUDA events, max over ranks."""
            algo2 = BasiciLQR(max_line_search=MAX_LINE_SEARCH, **solver_kw).init(getattr(models, model)(T=horizon))
            xh, th = bench_inputs(model, batch, 1234 + rank, horizon)
            xd, td = algo2._dev.real(xh), algo2._dev.real(th)
            o2 = algo2.solve_batch(xd, td)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tot = 0.0
            for _ in range(steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o2 = algo2.solve_batch(xd, td, out=o2)
                if world > 1:
                    gather_results(o2.obj_fun_value, o2.iterations, o2.converged, batch * world)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            it_loc = float(o2.iterations.sum().item())
            st = torch.tensor([tot / steps, it_loc], dtype=torch.float64, device=dev)
            if world > 1:
                mx, sm_ = st.clone(), st.clone()
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
                ms2, it_all = float(mx[0]), float(sm_[1])
            else:
                ms2, it_all = float(st[0]), it_loc
            d2 = algo2._dev
            bpi2 = bytes_per_traj_iter(d2.n, d2.m, horizon, 8)
            rec = {"workload": name, "model": model, "batch_per_gpu": batch, "horizon": horizon, "dtype": "f64",
                   "steps": steps, "warmup": 1, "ms_per_step": ms2, "traj_iters_per_step": it_all,
                   "value": it_all / (ms2 * 1e-3), "unit": UNIT,
                   "frac": (it_loc * bpi2) / (ms2 * 1e-3) / 1e9 / peak_also,
                   "algorithmic_bytes_per_traj_iter": bpi2, "timing": "CUDA events on the launch stream, max over ranks"}
            del algo2, xd, td, o2
            torch.cuda.empty_cache()
            return rec

        also.append(measure("fixed work: ThreeLink 65,536 per GPU, T=100, FP64, is_check_stop=False, max_iter=20 (SURVEY 8(d))",
                            MODEL, 65536, 100, 3, is_check_stop=False, max_iter=20))
        if world == 1:
            also.append(measure("north_star configuration: ThreeLink 1,048,576 trajectories on ONE GPU, T=100, FP64, to convergence",
                                MODEL, 1048576, 100, 2))
            also.append(measure("BASELINE config 3: RoboticArmTracking 262,144, T=200, FP64, to convergence",
                                "RoboticArmTracking", 262144, 200, 2))
        else:
            also.append(measure("BASELINE config 4 shard size: ThreeLink 131,072 per GPU (x %d GPUs = %d), T=100, FP64, to "
                                "convergence, terminal gather included" % (world, 131072 * world), MODEL, 131072, 100, 2))
        if world == 1:
            also.append(measure_nn_jacobian(dev))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    s = 8 if args.dtype == "float64" else 4
    bpi = bytes_per_traj_iter(dm.n, dm.m, T, s)
    peak, peak_src = measured_hbm_peak()
    achieved = (iters_local * bpi) / (ms_per_step * 1e-3) / 1e9           # rank 0's kernel: its own bytes / its own time
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "traj_iters_per_step": iters_all, "converged_fraction": conv_all / total_B,
            "mean_iters_per_traj": iters_all / total_B,
            "gpu_launches": launches_timed,   # BasiciLQR: one persistent solve_kernel per step (+ one memset node of its control
                                              # block); LogBarrieriLQR: one per barrier parameter + the real-cost kernel
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3,
                    "trajectory_output": "written by the kernel into pinned host memory" if (args.solver == "basic" and direct)
                                         else "device-to-host copy after the solve"},
            "roofline": {"bound": "hbm", "kernel": "%s<%s,%s>" % ("gen_solve_kernel" if dm.generic else "solve_kernel", args.model, args.dtype),
                         "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_traj_iter": bpi,
                         "algorithmic_bytes_per_launch": iters_local * bpi}}
    tr = profiled_traffic(args.model, args.dtype, B, args.mode) if args.solver == "basic" else None   # captured for BasiciLQR only
    if tr is not None:
        line["roofline"]["traffic"] = tr["dram_bytes"]
        line["roofline"]["traffic_source"] = tr.get("source")
    if not args.no_also and args.solver == "basic" and args.model == MODEL and args.mode == "converge" and args.batch == 65536 \
            and args.dtype == "float64":
        line["also"] = also
    if world == 1 and not args.no_cpu_baseline:
        per_core = args.cpu_sample or 256           # the first 256 x cores trajectories (~15-20 s)
        try:
            arm = cpu_arm(args)
            iters, wall = arm.sample(per_core, seed=1234)
            line["cpu_baseline"] = {"value": iters / wall, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                    "sample": cpu_description(arm, per_core, iters, wall)}
            if arm.kind == "reference":
                it2, w2 = arm.sample(max(32, per_core // 4), seed=1234, as_shipped=False)
                line["cpu_baseline"]["numerics_only_value"] = it2 / w2
            arm.close()
        except Exception as exc:      # the baseline is reporting only; never lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (exc,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
