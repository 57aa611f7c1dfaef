"""The UNMODIFIED reference as a CPU worker pool.  TEST / BENCHMARK INFRASTRUCTURE (never imported by the product).

`bench.py --impl reference` and the `cpu_baseline` leg of the default run use this when a reference tree is available
on the box (baseline/_ref/, installed by oracle/vendor_reference.sh): one worker process per host core, each driving the
reference's own public API exactly like its Example scripts do (Example/ThreeLinkPlanarManipulatorInteractive.py:8-30):

    scenario = <Scenario>(T=...);  algo = BasiciLQR(max_line_search=10).init(scenario)
    per problem:  get/set_obj_add_param(target), set_obj_fun_value(inf), set_init_state(x0), solve()

Two flavours:
  as shipped      solve() untouched - every iteration ends with logger.save_to_json(trajectory=X.tolist()), i.e. a
                  json.dumps of the whole trajectory (basic_ilqr.py:93, Logger.py:69), about half of the CPU time;
  numerics only   the same with that one call replaced by a no-op.
The per-iteration log LINES are silenced in both (logger.set_level(45), Logger.py:101-121): they are terminal output,
not work.  The first solve of a process pays numba's JIT (10-25 s, no cache=True anywhere in the reference) and is
excluded, as the reference itself excludes its first iteration from its timing (basic_ilqr.py:82-83).
Iterations are counted by wrapping the bound method `forward_pass`, which `solve()` looks up through `self`.
"""
import os
import time

import numpy as np

_W = {}


def init_worker(root, model, T, solver, barrier):
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "NUMBA_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_shim
    ref_shim.install(root)
    from CuriousRL.utils.Logger import logger as ref_logger
    import CuriousRL.scenario.dynamic_model as ref_models
    from CuriousRL.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR
    ref_logger.set_level(45)
    scenario = getattr(ref_models, model)(T=T)
    algo = (BasiciLQR if solver == "basic" else LogBarrieriLQR)(max_line_search=10)
    algo.init(scenario)
    counter = {"n": 0}
    plain_forward = algo.forward_pass

    def forward_pass(*a, **k):
        counter["n"] += 1
        return plain_forward(*a, **k)

    algo.forward_pass = forward_pass
    _W.update(algo=algo, scenario=scenario, counter=counter, logger=ref_logger, save=ref_logger.save_to_json,
              barrier=barrier, has_target=scenario.add_param is not None and model in (
                  "TwoLinkPlanarManipulator", "ThreeLinkPlanarManipulator", "RoboticArmTracking"))
    solve_one(np.asarray(scenario.init_state, dtype=np.float64).reshape(-1), None)     # JIT warm-up (untimed)


def ready(_):
    _W["barrier"].wait(timeout=1200)
    return 0


def solve_one(x0, target):
    algo = _W["algo"]
    if target is not None and _W["has_target"]:
        ap = algo.get_obj_add_param()
        ap[:, 0] = target[0]
        ap[:, 1] = target[1]
        algo.set_obj_add_param(ap)
    algo.set_init_state(np.asarray(x0, dtype=np.float64).reshape(-1, 1))
    algo.set_obj_fun_value(np.inf)
    algo.solve()


def solve_jobs(job):
    """job = (x0 [k, n], targets [k, 2] or None, as_shipped) -> (iterations, seconds)"""
    x0, tgt, as_shipped = job
    lg = _W["logger"]
    lg.save_to_json = _W["save"] if as_shipped else (lambda **kw: None)
    _W["counter"]["n"] = 0
    t0 = time.perf_counter()
    for b in range(x0.shape[0]):
        solve_one(x0[b], None if tgt is None else tgt[b])
    dt = time.perf_counter() - t0
    lg.save_to_json = _W["save"]
    return _W["counter"]["n"], dt


class ReferencePool:
    """One worker per host core, each running the unmodified reference single-threaded."""

    def __init__(self, root, model, T, cores=None, solver="basic"):
        import multiprocessing as mp
        self.cores = cores or min(os.cpu_count() or 1, 64)
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=init_worker, initargs=(root, model, T, solver, ctx.Barrier(self.cores)))
        self.pool.map(ready, range(self.cores), chunksize=1)       # returns once every worker has paid its JIT

    def run(self, x0, tgt, as_shipped=True):
        """Solve the problems (split round-robin over the cores); returns (iterations, wall seconds)."""
        jobs = [(x0[c::self.cores], None if tgt is None else tgt[c::self.cores], as_shipped) for c in range(self.cores)]
        t0 = time.perf_counter()
        out = self.pool.map(solve_jobs, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out), wall

    def close(self):
        self.pool.close()
        self.pool.join()
