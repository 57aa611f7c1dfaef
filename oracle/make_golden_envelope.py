"""Noise envelope of the UNMODIFIED reference on the seeded RoboticArm batches.  TEST INFRASTRUCTURE.

    python oracle/make_golden_envelope.py          # needs /root/reference (build container only)

SURVEY.md section 4.3: RoboticArm is reference-unstable - perturbing the gains K, k by 1e-15 relative changes the
reference's OWN iteration count and final cost.  For every problem of the seeded batches of
tests/golden/RoboticArm_T{100,200}.npz this script re-solves with K, k multiplied by (1 + eps N(0, 1)) for
eps in {1e-15, 1e-12} (the two levels of SURVEY appendix A) and, as a third family, with the INPUTS F, c of the
backward pass perturbed by 1e-15 relative, RUNS seeds each, and records per run the iteration
count, the final cost, the accepted-step trace and the distance of the final trajectory from the unperturbed one.
tests/test_gpu_solve_parity.py::test_robotic_arm_envelope_rule judges the CUDA path per problem against these.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden as mg  # noqa: E402  (installs the import shim, imports the reference)

RUNS = 8
LEVELS = (1e-15, 1e-12)          # relative noise on the OUTPUTS of backward_pass (K, k): SURVEY appendix A
INPUT_EPS = 1e-15                # relative noise on its INPUTS (F, c): what 1-ulp differences in sin / cos amount to.
# The backward recursion itself amplifies rounding (|V| grows to 1e21 at T = 200), so this third family is the one
# that tells how far two correct implementations of the recursion can be apart; noise on K, k enters after it.


def input_noise(rec):
    """Route the recorder's backward_pass through a perturbation of F and c when rec.input_rng is set."""
    plain = rec._backward

    def backward(C, c, F):
        if rec.input_rng is not None:
            F = F * (1.0 + INPUT_EPS * rec.input_rng.standard_normal(F.shape))
            c = c * (1.0 + INPUT_EPS * rec.input_rng.standard_normal(c.shape))
        return plain(C, c, F)

    rec.input_rng = None
    rec._backward = backward


def make(tag):
    cls_name, kwargs, batch = mg.CONFIGS[tag]
    base = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    scenario = getattr(mg.ref_models, cls_name)(**kwargs)
    algo = mg.BasiciLQR(max_line_search=mg.MAX_LINE_SEARCH)
    algo.init(scenario)
    rec = mg.Recorder(algo)
    input_noise(rec)
    x0s, targets = base["batch_x0"], base["batch_target"]
    out = dict(model=np.str_(cls_name), T=np.int64(scenario.T), levels=np.asarray(LEVELS), runs=np.int64(RUNS))
    NL = len(LEVELS) + 1                         # families: K,k at LEVELS[0], K,k at LEVELS[1], F,c at INPUT_EPS
    iters = np.zeros((batch, NL, RUNS), dtype=np.int32)
    J = np.zeros((batch, NL, RUNS))
    Xerr = np.zeros((batch, NL, RUNS))
    traces = []
    for b in range(batch):
        plain = mg.run(algo, rec, x0s[b], targets[b])
        assert plain["iters"] == int(base["batch_iters"][b]) and plain["J"] == float(base["batch_J"][b]), \
            "the reference no longer reproduces the committed golden batch"
        for li in range(NL):
            for s in range(RUNS):
                if li < len(LEVELS):
                    r = mg.run(algo, rec, x0s[b], targets[b], noise_seed=90000 + 1000 * b + 100 * li + s, noise_eps=LEVELS[li])
                else:
                    rec.input_rng = np.random.default_rng(70000 + 1000 * b + s)
                    r = mg.run(algo, rec, x0s[b], targets[b])
                    rec.input_rng = None
                iters[b, li, s], J[b, li, s] = r["iters"], r["J"]
                Xerr[b, li, s] = np.max(np.abs(r["X"] - plain["X"])) / max(1.0, np.max(np.abs(plain["X"])))
                traces.append(r["alphas"])
        print("%s b=%d: reference %d iters J=%.10g | K,k 1e-15: iters %s | K,k 1e-12: %s | F,c 1e-15: %s" % (
            tag, b, plain["iters"], plain["J"], iters[b, 0].tolist(), iters[b, 1].tolist(), iters[b, 2].tolist()), flush=True)
    out["iters"], out["J"], out["Xerr"] = iters, J, Xerr
    out["alphas"] = mg.pad(traces, -2, np.int8).reshape(batch, NL, RUNS, -1)
    out["input_eps"] = np.float64(INPUT_EPS)
    path = os.path.join(ROOT, "tests", "golden", "Envelope_" + tag + ".npz")
    np.savez_compressed(path, **out)
    print("->", path, "%.0f KB" % (os.path.getsize(path) / 1024))


F32_EPS = 6e-8                   # float32 unit round-off


def make_f32(tag):
    """What an FP32 solver can be expected to reproduce (north_star: FP32 mode <= 1e-4 relative on the final cost): the
    reference re-solved with float32-sized relative noise on the outputs (K, k) and on the inputs (F, c) of its backward
    pass, RUNS seeds each.  A problem whose own cost moves by more than 1e-4 under either is outside that bar by the
    reference's own standard.  -> tests/golden/Envelope32_<tag>.npz"""
    global INPUT_EPS
    cls_name, kwargs, batch = mg.CONFIGS[tag]
    base = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    scenario = getattr(mg.ref_models, cls_name)(**kwargs)
    algo = mg.BasiciLQR(max_line_search=mg.MAX_LINE_SEARCH)
    algo.init(scenario)
    rec = mg.Recorder(algo)
    input_noise(rec)
    saved, INPUT_EPS = INPUT_EPS, F32_EPS
    x0s, targets = base["batch_x0"], base["batch_target"]
    iters = np.zeros((batch, 2, RUNS), dtype=np.int32)
    J = np.zeros((batch, 2, RUNS))
    for b in range(batch):
        for s in range(RUNS):
            r = mg.run(algo, rec, x0s[b], targets[b], noise_seed=30000 + 1000 * b + s, noise_eps=F32_EPS)
            iters[b, 0, s], J[b, 0, s] = r["iters"], r["J"]
            rec.input_rng = np.random.default_rng(40000 + 1000 * b + s)
            r = mg.run(algo, rec, x0s[b], targets[b])
            rec.input_rng = None
            iters[b, 1, s], J[b, 1, s] = r["iters"], r["J"]
        Jref = float(base["batch_J"][b])
        print("%s b=%d: reference %d iters | K,k 6e-8: iters %s dJ %.1e | F,c 6e-8: iters %s dJ %.1e" % (
            tag, b, int(base["batch_iters"][b]), iters[b, 0].tolist(), np.abs(J[b, 0] - Jref).max() / abs(Jref),
            iters[b, 1].tolist(), np.abs(J[b, 1] - Jref).max() / abs(Jref)), flush=True)
    INPUT_EPS = saved
    path = os.path.join(ROOT, "tests", "golden", "Envelope32_" + tag + ".npz")
    np.savez_compressed(path, model=np.str_(cls_name), T=np.int64(scenario.T), eps=np.float64(F32_EPS), runs=np.int64(RUNS),
                        iters=iters, J=J)
    print("->", path)


if __name__ == "__main__":
    which = sys.argv[1:] or ["f64", "f32"]
    if "f64" in which:
        for tag in ("RoboticArm_T100", "RoboticArm_T200"):
            make(tag)
    if "f32" in which:
        for tag in ("TwoLink_T100", "ThreeLink_T100", "RoboticArm_T100"):
            make_f32(tag)
