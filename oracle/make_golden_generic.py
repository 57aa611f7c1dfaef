"""Generate tests/golden/Generic_*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE.

The remaining model-based scenarios: CartPoleSwingUp1/2, VehicleTracking,
CarParking under `BasiciLQR(max_line_search=10, max_iter=100)` (the survey's observation setting).  Needs
/root/reference; run in the build container:

    python oracle/make_golden_generic.py [tag-substring ...]

Per scenario: the default problem (per-iteration cost / accepted step / rollouts, final trajectory, stage tensors
X, F, c, C, K, k and every candidate at iterations 0, 1, last), a seeded batch of perturbed initial states
(curiousrl_b200.utils.synthetic.synthetic_init_states, seed 0) and the 1e-15 noise envelope of both
(SURVEY.md section 4.3).  Also asserts that curiousrl_b200's scenario classes carry the same symbolic
expressions, bounds and parameters as the reference's.
"""
import os
import sys
import time

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden as mg  # noqa: E402  (installs the import shim, provides Recorder / run / pad)
from CuriousRL.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
import CuriousRL.scenario.dynamic_model as ref_models  # noqa: E402
import curiousrl_b200.scenario.dynamic_model as our_models  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_init_states  # noqa: E402

CONFIGS = {
    # tag: (class name, ctor kwargs, batch size)
    "Generic_CartPole1_T150": ("CartPoleSwingUp1", dict(T=150), 6),
    "Generic_CartPole2_T150": ("CartPoleSwingUp2", dict(T=150), 6),
    "Generic_Vehicle_T150": ("VehicleTracking", dict(T=150), 6),
    "Generic_CarParking_T500": ("CarParking", dict(T=500), 4),
    "Generic_Vehicle_T40": ("VehicleTracking", dict(T=40), 6),          # short horizon, fast CPU checks
    # the reference's default max_line_search = 50 (CarParking 17 iterations, J = 22258.98642621647)
    "Generic_CarParking_T500_ls50": ("CarParking", dict(T=500), 3, 50),
}
MAX_ITER = 100


def same_scenario(ref, ours):
    assert sp.srepr(ref.dynamic_function) == sp.srepr(ours.dynamic_function)
    assert sp.srepr(ref.obj_fun) == sp.srepr(ours.obj_fun)
    assert np.array_equal(ref.constr, ours.constr) and np.array_equal(ref.init_state, ours.init_state)
    assert np.array_equal(ref.init_action, ours.init_action)
    assert (ref.add_param is None) == (ours.add_param is None)
    assert ref.add_param is None or np.array_equal(ref.add_param, ours.add_param)
    assert str(ref.add_param_var) == str(ours.add_param_var) and ref.name == ours.name


def make(tag):
    cls_name, kwargs, batch = CONFIGS[tag][:3]
    max_ls = CONFIGS[tag][3] if len(CONFIGS[tag]) > 3 else mg.MAX_LINE_SEARCH
    t0 = time.time()
    scenario = getattr(ref_models, cls_name)(**kwargs)
    same_scenario(scenario, getattr(our_models, cls_name)(**kwargs))
    algo = BasiciLQR(max_line_search=max_ls, max_iter=MAX_ITER)
    algo.init(scenario)
    x0_default = scenario.init_state[:, 0].copy()
    out = dict(model=np.str_(cls_name), T=np.int64(scenario.T), n=np.int64(scenario.n), m=np.int64(scenario.m),
               max_line_search=np.int64(max_ls), max_iter=np.int64(MAX_ITER))
    rec = mg.Recorder(algo)
    first = mg.run(algo, rec, x0_default)
    rec.stage_iters = {0, 1, first["iters"] - 1}
    res = mg.run(algo, rec, x0_default)
    assert res["iters"] == first["iters"] and res["J"] == first["J"]
    out["default_x0"] = x0_default
    for key in ("X", "J", "iters", "Js", "alphas", "rollouts"):
        out["default_" + key] = np.asarray(res[key])
    algo.set_init_state(x0_default.reshape(-1, 1))
    out["default_X_init"] = mg.scenario_initial_rollout(algo)
    out["stage_iters"] = np.asarray(sorted(rec.stages), dtype=np.int64)
    for it, st in rec.stages.items():
        for key, val in st.items():
            out["stage%d_%s" % (it, key)] = val
    rec.stage_iters = set()

    x0s = synthetic_init_states(scenario, batch, seed=0)
    finals = [mg.run(algo, rec, x0s[b]) for b in range(batch)]
    out["batch_x0"] = x0s
    out["batch_iters"] = np.asarray([r["iters"] for r in finals], dtype=np.int32)
    out["batch_J"] = np.asarray([r["J"] for r in finals])
    out["batch_X"] = np.stack([r["X"] for r in finals])
    out["batch_Js"] = mg.pad([r["Js"] for r in finals], np.nan, np.float64)
    out["batch_alphas"] = mg.pad([r["alphas"] for r in finals], -2, np.int32)
    n_iters = np.zeros((batch, mg.NOISE_RUNS), dtype=np.int32)
    n_J = np.zeros((batch, mg.NOISE_RUNS))
    n_Xerr = np.zeros((batch, mg.NOISE_RUNS))
    for b in range(batch):
        for s in range(mg.NOISE_RUNS):
            r = mg.run(algo, rec, x0s[b], noise_seed=1000 * b + s)
            n_iters[b, s], n_J[b, s] = r["iters"], r["J"]
            n_Xerr[b, s] = np.max(np.abs(r["X"] - finals[b]["X"])) / max(1.0, np.max(np.abs(finals[b]["X"])))
    out["noise_iters"], out["noise_J"], out["noise_Xerr"] = n_iters, n_J, n_Xerr
    d_iters = np.zeros(mg.NOISE_RUNS, dtype=np.int32)
    d_J = np.zeros(mg.NOISE_RUNS)
    d_Xerr = np.zeros(mg.NOISE_RUNS)
    for s in range(mg.NOISE_RUNS):
        r = mg.run(algo, rec, x0_default, noise_seed=77 + s)
        d_iters[s], d_J[s] = r["iters"], r["J"]
        d_Xerr[s] = np.max(np.abs(r["X"] - res["X"])) / max(1.0, np.max(np.abs(res["X"])))
    out["default_noise_iters"], out["default_noise_J"], out["default_noise_Xerr"] = d_iters, d_J, d_Xerr
    path = os.path.join(ROOT, "tests", "golden", tag + ".npz")
    np.savez_compressed(path, **out)
    print("%-24s default: %3d iters J=%.16g (noise: %s, Xerr %.1e) | batch iters %s noise-stable %s | %.1f s (%.0f KB)" % (
        tag, res["iters"], res["J"], d_iters.tolist(), d_Xerr.max(), out["batch_iters"].tolist(),
        np.all(n_iters == out["batch_iters"][:, None], axis=1).astype(int).tolist(), time.time() - t0,
        os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for tag in [t for t in CONFIGS if not sys.argv[1:] or any(a in t for a in sys.argv[1:])]:
        make(tag)
