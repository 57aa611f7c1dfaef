"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # all configs
    python oracle/make_golden.py ThreeLink  # one config

For every config it records, with `BasiciLQR(max_line_search=10)`:

* the scenario's default problem: per-iteration cost / accepted step index /
  rollout count, final trajectory, and full stage tensors (X, F, c, C, K, k,
  the candidate trajectory and cost of every alpha tried) at iterations
  0, 1 and the last one;
* a seeded synthetic batch (curiousrl_b200.utils.synthetic, seed 0): x0,
  target, iterations, final cost, final trajectory, step-index and cost traces;
* the same batch re-solved with K, k perturbed by 1e-15 relative noise
   so tests can tell noise-stable problems from
  knife-edge ones, and by 6e-8 relative noise (float32 round-off) to
  delimit what an FP32 solver can reproduce.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from CuriousRL.utils.Logger import logger as ref_logger  # noqa: E402
from CuriousRL.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
import CuriousRL.scenario.dynamic_model as ref_models  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

ref_logger.set_level(45)

CONFIGS = {
    # tag: (class name, ctor kwargs, batch size of the seeded set)
    "TwoLink_T100": ("TwoLinkPlanarManipulator", dict(T=100), 16),
    "ThreeLink_T100": ("ThreeLinkPlanarManipulator", dict(T=100), 16),
    "RoboticArm_T100": ("RoboticArmTracking", dict(T=100), 16),
    "RoboticArm_T200": ("RoboticArmTracking", dict(T=200), 8),
    "ThreeLink_T30": ("ThreeLinkPlanarManipulator", dict(T=30), 8),   # short horizon, fast CPU checks
}
MAX_LINE_SEARCH = 10
NOISE_RUNS = 4
NOISE_EPS = 1e-15
NOISE_EPS_F32 = 6e-8      # float32 unit round-off: the perturbation scale an FP32 solver injects


class Recorder:
    """Wraps the bound methods the reference looks up through `self`."""

    def __init__(self, algo, stage_iters=()):
        self.algo = algo
        self.stage_iters = set(stage_iters)
        self.reset()
        dm = algo.dynamic_model
        self._update = dm.update_traj
        self._forward = algo.forward_pass
        self._backward = algo.backward_pass
        dm.update_traj = self.update_traj
        algo.forward_pass = self.forward_pass
        algo.backward_pass = self.backward_pass
        self.noise_rng = None
        self.noise_eps = NOISE_EPS

    def reset(self):
        self.Js, self.alphas, self.rollouts, self.stages = [], [], [], {}
        self._n_update = 0
        self._cands = []
        self._bw = None
        self.it = 0

    def update_traj(self, old, K, k, alpha):
        out = self._update(old, K, k, alpha)
        self._n_update += 1
        if self.it in self.stage_iters:
            self._cands.append((alpha, out.copy(), float(self.algo.obj_fun.eval_obj_fun(out))))
        return out

    def backward_pass(self, C, c, F):
        K, k = self._backward(C, c, F)
        if self.noise_rng is not None:
            K = K * (1.0 + self.noise_eps * self.noise_rng.standard_normal(K.shape))
            k = k * (1.0 + self.noise_eps * self.noise_rng.standard_normal(k.shape))
        if self.it in self.stage_iters:
            self._bw = dict(C=C.copy(), c=c[:, :, 0].copy(), F=F.copy(), K=K.copy(), k=k[:, :, 0].copy())
        return K, k

    def forward_pass(self, traj, K, k):
        self._n_update = 0
        self._cands = []
        J_last = self.algo.get_obj_fun_value()
        out = self._forward(traj, K, k)
        new_traj, J = out[0], out[4]
        accepted = new_traj is not traj
        self.Js.append(float(J))
        self.rollouts.append(self._n_update)
        self.alphas.append(self._n_update - 1 if accepted else -1)
        if self.it in self.stage_iters:
            st = dict(self._bw)
            st["X"] = traj[:, :, 0].copy()
            st["J_last"] = np.float64(J_last)
            st["cand_alpha"] = np.asarray([a for a, _, _ in self._cands])
            st["cand_X"] = np.stack([x[:, :, 0] for _, x, _ in self._cands])
            st["cand_J"] = np.asarray([j for _, _, j in self._cands])
            st["X_next"] = new_traj[:, :, 0].copy()
            st["J_next"] = np.float64(J)
            st["stop"] = np.bool_(out[5])
            self.stages[self.it] = st
        self.it += 1
        return out


def run(algo, rec, x0=None, target=None, noise_seed=None, noise_eps=NOISE_EPS):
    rec.reset()
    rec.noise_rng = None if noise_seed is None else np.random.default_rng(noise_seed)
    rec.noise_eps = noise_eps
    if target is not None:
        ap = algo.get_obj_add_param()
        ap[:, 0] = target[0]
        ap[:, 1] = target[1]
        algo.set_obj_add_param(ap)
    if x0 is not None:
        algo.set_init_state(np.asarray(x0, dtype=np.float64).reshape(-1, 1))
    algo.set_obj_fun_value(np.inf)
    algo.solve()
    return dict(X=algo._trajectory[:, :, 0].copy(), J=float(algo.get_obj_fun_value()), iters=len(rec.Js),
                Js=np.asarray(rec.Js), alphas=np.asarray(rec.alphas, dtype=np.int32),
                rollouts=np.asarray(rec.rollouts, dtype=np.int32))


def pad(rows, fill, dtype):
    width = max(len(r) for r in rows)
    out = np.full((len(rows), width), fill, dtype=dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def make(tag):
    cls_name, kwargs, batch = CONFIGS[tag]
    t0 = time.time()
    scenario = getattr(ref_models, cls_name)(**kwargs)
    algo = BasiciLQR(max_line_search=MAX_LINE_SEARCH)
    algo.init(scenario)
    x0_default = scenario.init_state[:, 0].copy()
    target_default = scenario.add_param[0].copy()
    out = dict(model=np.str_(cls_name), T=np.int64(scenario.T), n=np.int64(scenario.n), m=np.int64(scenario.m),
               max_line_search=np.int64(MAX_LINE_SEARCH))

    # -- default problem: first find the iteration count, then record stages 0, 1, last
    rec = Recorder(algo)
    first = run(algo, rec)
    last_it = first["iters"] - 1
    rec.stage_iters = {0, 1, last_it}
    res = run(algo, rec)
    assert res["iters"] == first["iters"] and res["J"] == first["J"]
    out["default_x0"] = x0_default
    out["default_target"] = target_default
    for key in ("X", "J", "iters", "Js", "alphas", "rollouts"):
        out["default_" + key] = np.asarray(res[key])
    out["default_X_init"] = scenario_initial_rollout(algo)
    out["stage_iters"] = np.asarray(sorted(rec.stages), dtype=np.int64)
    for it, st in rec.stages.items():
        for key, val in st.items():
            out["stage%d_%s" % (it, key)] = val
    rec.stage_iters = set()

    # -- seeded synthetic batch
    x0s, targets = synthetic_batch(cls_name, batch, seed=0)
    finals = [run(algo, rec, x0s[b], targets[b]) for b in range(batch)]
    out["batch_x0"] = x0s
    out["batch_target"] = targets
    out["batch_iters"] = np.asarray([r["iters"] for r in finals], dtype=np.int32)
    out["batch_J"] = np.asarray([r["J"] for r in finals])
    out["batch_X"] = np.stack([r["X"] for r in finals])
    out["batch_Js"] = pad([r["Js"] for r in finals], np.nan, np.float64)
    out["batch_alphas"] = pad([r["alphas"] for r in finals], -2, np.int32)

    # -- noise envelope of the reference itself (K, k perturbed by 1e-15 relative)
    n_iters = np.zeros((batch, NOISE_RUNS), dtype=np.int32)
    n_J = np.zeros((batch, NOISE_RUNS))
    n_Xerr = np.zeros((batch, NOISE_RUNS))
    for b in range(batch):
        for s in range(NOISE_RUNS):
            r = run(algo, rec, x0s[b], targets[b], noise_seed=1000 * b + s)
            n_iters[b, s] = r["iters"]
            n_J[b, s] = r["J"]
            ref = finals[b]["X"]
            n_Xerr[b, s] = np.max(np.abs(r["X"] - ref)) / max(1.0, np.max(np.abs(ref)))
    out["noise_iters"], out["noise_J"], out["noise_Xerr"] = n_iters, n_J, n_Xerr
    # the same with float32-sized noise: which problems can an FP32 solver be expected to reproduce?
    f_iters = np.zeros((batch, NOISE_RUNS), dtype=np.int32)
    f_J = np.zeros((batch, NOISE_RUNS))
    for b in range(batch):
        for s in range(NOISE_RUNS):
            r = run(algo, rec, x0s[b], targets[b], noise_seed=5000 + 1000 * b + s, noise_eps=NOISE_EPS_F32)
            f_iters[b, s], f_J[b, s] = r["iters"], r["J"]
    out["noise32_iters"], out["noise32_J"] = f_iters, f_J
    d_iters = np.zeros(NOISE_RUNS, dtype=np.int32)
    d_J = np.zeros(NOISE_RUNS)
    for s in range(NOISE_RUNS):
        r = run(algo, rec, x0_default, target_default, noise_seed=77 + s)
        d_iters[s], d_J[s] = r["iters"], r["J"]
    out["default_noise_iters"], out["default_noise_J"] = d_iters, d_J

    path = os.path.join(ROOT, "tests", "golden", tag + ".npz")
    np.savez_compressed(path, **out)
    print("%-16s default: %3d iters J=%.16g | batch iters %s | %.1f s -> %s (%.0f KB)" % (
        tag, res["iters"], res["J"], out["batch_iters"].tolist(), time.time() - t0, path,
        os.path.getsize(path) / 1024))


def scenario_initial_rollout(algo):
    return algo.dynamic_model.eval_traj()[:, :, 0].copy()


if __name__ == "__main__":
    tags = [t for t in CONFIGS if not sys.argv[1:] or any(a in t for a in sys.argv[1:])]
    for tag in tags:
        make(tag)
