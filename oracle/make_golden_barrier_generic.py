"""Generate tests/golden/LogBarrierGeneric_*.npz from the UNMODIFIED reference's LogBarrieriLQR on the scenarios with a
generated device model (CartPoleSwingUp1/2, VehicleTracking, CarParking).  TEST INFRASTRUCTURE; needs /root/reference.

    python oracle/make_golden_barrier_generic.py [tag ...]

`LogBarrieriLQR(max_line_search=10, max_iter=100)`: here the barrier also sits on the finite STATE bounds
(log_barrier_ilqr.py:91-95 runs over every row of `constr`).  Per scenario: barrier cost / gradient / Hessian diagonal at
the initial trajectory, the default problem (per-iteration barrier and real cost, accepted step, stop flag, iterations of
every inner loop, final trajectory) and a seeded batch of perturbed initial states solved one by one.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden_barrier as mb  # noqa: E402  (import shim, Recorder)
from CuriousRL.algorithm.ilqr_solver import LogBarrieriLQR  # noqa: E402
import CuriousRL.scenario.dynamic_model as ref_models  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_init_states  # noqa: E402

CONFIGS = {
    "LogBarrierGeneric_CartPole1_T150": ("CartPoleSwingUp1", dict(T=150), 4),
    "LogBarrierGeneric_CartPole2_T150": ("CartPoleSwingUp2", dict(T=150), 3),
    "LogBarrierGeneric_Vehicle_T150": ("VehicleTracking", dict(T=150), 4),
    "LogBarrierGeneric_CarParking_T500": ("CarParking", dict(T=500), 3),
}
MAX_ITER = 100


def run(tag):
    cls, kw, nb = CONFIGS[tag]
    sc = getattr(ref_models, cls)(**kw)
    algo = LogBarrieriLQR(max_line_search=mb.MAX_LINE_SEARCH, max_iter=MAX_ITER).init(sc)
    rec = mb.Recorder(algo)
    out = {"model": np.str_(cls), "T": np.int64(kw["T"]), "t_list": np.asarray(algo._t, dtype=np.float64),
           "max_line_search": np.int64(mb.MAX_LINE_SEARCH), "max_iter": np.int64(MAX_ITER)}
    X0 = algo.dynamic_model.eval_traj()
    out["init_X"] = X0[:, :, 0].copy()
    out["init_J"] = np.float64(algo.obj_fun.eval_obj_fun(X0))
    out["init_c"] = algo.obj_fun.eval_grad_obj_fun(X0)[:, :, 0].copy()
    out["init_C"] = algo.obj_fun.eval_hessian_obj_fun(X0).copy()
    algo.solve()
    out["def_J"] = np.asarray(rec.J)
    out["def_Jreal"] = np.asarray(rec.Jreal)
    out["def_alpha"] = np.asarray(rec.alpha, dtype=np.int64)
    out["def_stop"] = np.asarray(rec.stop)
    out["def_inner_iters"] = np.asarray(inner_iters(rec, len(algo._t)), dtype=np.int64)
    out["def_X"] = algo._trajectory[:, :, 0].copy()
    print("%s default: inner iterations %s (total %d), real J = %.17g" % (tag, out["def_inner_iters"].tolist(), len(rec.J), rec.Jreal[-1]))
    x0 = synthetic_init_states(sc, nb, seed=0)
    b_inner, b_J, b_X = [], [], []
    for b in range(nb):
        algo_b = LogBarrieriLQR(max_line_search=mb.MAX_LINE_SEARCH, max_iter=MAX_ITER).init(getattr(ref_models, cls)(**kw))
        rec_b = mb.Recorder(algo_b)
        algo_b.set_init_state(x0[b].reshape(-1, 1))
        algo_b.solve()
        b_inner.append(inner_iters(rec_b, len(algo_b._t)))
        b_J.append(rec_b.Jreal[-1])
        b_X.append(algo_b._trajectory[:, :, 0].copy())
        print("  batch %d: inner %s real J %.17g" % (b, b_inner[-1], rec_b.Jreal[-1]))
    out["batch_x0"] = x0
    out["batch_inner_iters"] = np.asarray(b_inner, dtype=np.int64)
    out["batch_Jreal"] = np.asarray(b_J)
    out["batch_X"] = np.stack(b_X)
    path = os.path.join(ROOT, "tests", "golden", tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


def inner_iters(rec, n_t):
    """Iterations of each inner loop: a loop ends with its first stop flag or after MAX_ITER iterations."""
    counts, n = [], 0
    for s in rec.stop:
        n += 1
        if s or n == MAX_ITER:
            counts.append(n)
            n = 0
    if n:
        counts.append(n)
    return counts + [0] * (n_t - len(counts))


if __name__ == "__main__":
    for tag in (sys.argv[1:] or list(CONFIGS)):
        run(tag)
