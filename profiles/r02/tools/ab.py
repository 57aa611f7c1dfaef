"""A/B timing of alternative builds of the solve kernel (development tool; B200 only).

    python profiles/r02/tools/ab.py LIB[,LIB...] CASE [CASE ...]

Every LIB is loaded in a fresh interpreter (CURIOUSRL_B200_LIB), every CASE is `name:batch:mode[:variant[:coop]]` with
mode = conv (to convergence) | fixN (N iterations each, no stopping test); prints the best of 3 kernel times, the
trajectory-iterations per second and a SHA-256 of (iterations, J, X) so that builds can be compared bit for bit.
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
MODELS = {"3l": ("ThreeLinkPlanarManipulator", 100), "2l": ("TwoLinkPlanarManipulator", 100),
          "ra": ("RoboticArmTracking", 200), "ra100": ("RoboticArmTracking", 100),
          "veh": ("VehicleTracking", 150), "cp1": ("CartPoleSwingUp1", 150), "cp2": ("CartPoleSwingUp2", 150),
          "park": ("CarParking", 500)}


def worker(cases):
    sys.path.insert(0, ROOT)
    import torch
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    import curiousrl_b200.scenario.dynamic_model as models
    from curiousrl_b200.utils.synthetic import synthetic_batch
    logger.set_level(45)
    lib = os.path.basename(os.environ.get("CURIOUSRL_B200_LIB", "default"))
    for case in cases:
        f = case.split(":")
        name, B, mode = f[0], int(f[1]), f[2]
        variant = int(f[3]) if len(f) > 3 else 0
        coop = int(f[4]) if len(f) > 4 else 0
        model, T = MODELS[name]
        kw = dict(max_line_search=10, verify=False, occupancy=variant, coop_threshold=coop)
        if mode.startswith("fix"):
            kw.update(is_check_stop=False, max_iter=int(mode[3:]))
        scen = getattr(models, model)(T=T)
        algo = BasiciLQR(**kw).init(scen)
        if algo._dev.generic:                       # generated models: the scenario's own initial state +- 0.05, no target
            from curiousrl_b200.utils.synthetic import synthetic_init_states
            x0d, tgd = algo._dev.real(synthetic_init_states(scen, B, seed=1234)), None
        else:
            x0, tgt = synthetic_batch(model, B, seed=1234)
            x0d, tgd = algo._dev.real(x0), algo._dev.real(tgt)
        best = 1e30
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            res = algo.solve_batch(x0d, tgd)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        h = hashlib.sha256()
        h.update(res.iterations.cpu().numpy().tobytes())
        h.update(res.obj_fun_value.cpu().numpy().tobytes())
        h.update(res.trajectory.cpu().numpy().tobytes())
        it = int(res.iterations.sum())
        print("%-12s %-22s %9.2f ms %8.3f M traj-iter/s  iters %d  sha %s" % (lib, case, best, it / best / 1e3, it,
                                                                            h.hexdigest()[:16]), flush=True)
        # -DCRL_PHASE_TIMERS builds: SM cycles per phase, summed over the warps (control-block words 8..16) of the LAST solve
        ph = algo._dev._workspace[:256].view(torch.int64)[8:17].cpu().numpy()
        if ph.any():
            names = ["refill", "backward", "linesearch", "retire", "coop_wait", "coop_gather", "coop_backward",
                     "coop_linesearch", "coop_output"]
            print("    phases (M cycles): " + "  ".join("%s %.2f" % (n, v / 1e6) for n, v in zip(names, ph)), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2:])
    else:
        for lib in sys.argv[1].split(","):
            env = dict(os.environ)
            if lib != "default":
                env["CURIOUSRL_B200_LIB"] = os.path.abspath(lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"] + sys.argv[2:], env=env)
