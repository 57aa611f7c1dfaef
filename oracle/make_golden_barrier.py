"""Generate tests/golden/LogBarrier_*.npz from the UNMODIFIED reference's LogBarrieriLQR.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):

    python oracle/make_golden_barrier.py

For every config, with `LogBarrieriLQR(max_line_search=10)` (the setting of
Example/RoboticArmTrackingIteractive.py:11) it records

* the scenario's default problem: per iteration the barrier cost, the real (barrier-free) cost, the
  accepted step index and the stop flag; the iteration count of every inner loop (one per barrier
  parameter t); the final trajectory; plus stage tensors of the barrier cost at the initial
  trajectory for t = t[0] (l, grad l, diag hess l per time step);
* a small seeded synthetic batch (curiousrl_b200.utils.synthetic, seed 0), each problem solved on its
  own: inner-loop iteration counts, final real cost, final trajectory.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from CuriousRL.utils.Logger import logger as ref_logger  # noqa: E402
from CuriousRL.algorithm.ilqr_solver import LogBarrieriLQR  # noqa: E402
import CuriousRL.scenario.dynamic_model as ref_models  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

ref_logger.set_level(45)

CONFIGS = {
    "LogBarrier_TwoLink_T100": ("TwoLinkPlanarManipulator", dict(T=100), 4),
    "LogBarrier_ThreeLink_T100": ("ThreeLinkPlanarManipulator", dict(T=100), 4),
    "LogBarrier_RoboticArm_T100": ("RoboticArmTracking", dict(T=100), 4),
}
MAX_LINE_SEARCH = 10


class Recorder:
    def __init__(self, algo):
        self.algo = algo
        self._forward = algo.forward_pass
        self._update = algo.dynamic_model.update_traj
        algo.forward_pass = self.forward_pass
        algo.dynamic_model.update_traj = self.update_traj
        self.reset()

    def reset(self):
        self.J, self.Jreal, self.alpha, self.stop = [], [], [], []
        self._n = 0

    def update_traj(self, old, K, k, alpha):
        self._n += 1
        return self._update(old, K, k, alpha)

    def forward_pass(self, traj, K, k):
        self._n = 0
        out = self._forward(traj, K, k)
        accepted = out[0] is not traj
        self.J.append(float(out[4]))
        self.Jreal.append(float(self.algo._real_obj_fun.eval_obj_fun(out[0])))
        self.alpha.append(self._n - 1 if accepted else -1)
        self.stop.append(bool(out[5]))
        return out

    def inner_iters(self):
        """Iterations of each inner loop: an inner loop ends with the first stop flag (log_barrier_ilqr.py:140-146)."""
        counts, n = [], 0
        for s in self.stop:
            n += 1
            if s:
                counts.append(n)
                n = 0
        if n:
            counts.append(n)
        return counts


def run(tag):
    cls, kw, nb = CONFIGS[tag]
    sc = getattr(ref_models, cls)(**kw)
    algo = LogBarrieriLQR(max_line_search=MAX_LINE_SEARCH).init(sc)
    rec = Recorder(algo)
    out = {"t_list": np.asarray(algo._t, dtype=np.float64), "max_line_search": np.int64(MAX_LINE_SEARCH)}
    # stage tensors at the initial trajectory, t = t[0]
    X0 = algo.dynamic_model.eval_traj()
    out["init_X"] = X0[:, :, 0].copy()
    out["init_J"] = np.float64(algo.obj_fun.eval_obj_fun(X0))
    out["init_c"] = algo.obj_fun.eval_grad_obj_fun(X0)[:, :, 0].copy()
    out["init_Cdiag"] = np.stack([np.diag(C) for C in algo.obj_fun.eval_hessian_obj_fun(X0)])
    algo.solve()
    out["def_J"] = np.asarray(rec.J)
    out["def_Jreal"] = np.asarray(rec.Jreal)
    out["def_alpha"] = np.asarray(rec.alpha, dtype=np.int64)
    out["def_stop"] = np.asarray(rec.stop)
    out["def_inner_iters"] = np.asarray(rec.inner_iters(), dtype=np.int64)
    out["def_X"] = algo._trajectory[:, :, 0].copy()
    print("%s default: inner iterations %s (total %d), real J = %.17g" % (tag, rec.inner_iters(), len(rec.J), rec.Jreal[-1]))
    # seeded batch, one problem at a time
    x0, tgt = synthetic_batch(cls, nb, seed=0)
    T = kw["T"]
    b_inner, b_J, b_X = [], [], []
    for b in range(nb):
        sc_b = getattr(ref_models, cls)(**kw)
        algo_b = LogBarrieriLQR(max_line_search=MAX_LINE_SEARCH).init(sc_b)
        rec_b = Recorder(algo_b)
        ap = algo_b.get_obj_add_param()
        ap[:, 0] = tgt[b, 0]
        ap[:, 1] = tgt[b, 1]
        algo_b.set_obj_add_param(ap)
        algo_b.set_init_state(x0[b].reshape(-1, 1))
        algo_b.solve()
        inner = rec_b.inner_iters()
        inner = inner + [0] * (len(algo_b._t) - len(inner))
        b_inner.append(inner)
        b_J.append(rec_b.Jreal[-1])
        b_X.append(algo_b._trajectory[:, :, 0].copy())
        print("  batch %d: inner %s real J %.17g" % (b, inner, rec_b.Jreal[-1]))
    out["batch_x0"] = x0
    out["batch_target"] = tgt
    out["batch_inner_iters"] = np.asarray(b_inner, dtype=np.int64)
    out["batch_Jreal"] = np.asarray(b_J)
    out["batch_X"] = np.stack(b_X)
    path = os.path.join(ROOT, "tests", "golden", tag + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    for tag in (sys.argv[1:] or list(CONFIGS)):
        run(tag)
