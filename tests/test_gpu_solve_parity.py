"""End-to-end parity of the persistent solve kernel (through the C ABI) with the reference.

* default problems and seeded batches of Two/ThreeLink: same iteration count, same accepted
  step sequence, final cost and trajectory <= 1e-9 relative (north_star FP64 bar) on problems
  the reference itself solves reproducibly under 1e-15 noise;
* RoboticArm: noise-envelope protocol;
* FP32: final cost <= 1e-4 relative on noise-stable problems;
* size-independent properties at the benchmark's batch size.
"""
import numpy as np
import pytest
import torch

from conftest import make_scenario

pytestmark = pytest.mark.gpu


def make_algo(model, T, **kw):
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    from curiousrl_b200 import logger
    logger.set_level(45)
    kw.setdefault("max_line_search", 10)
    return BasiciLQR(**kw).init(make_scenario(model, T))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def stable_mask(g, tol=1e-9):
    """Problems whose reference solution is reproducible under 1e-15 perturbation of K, k."""
    it_ok = np.all(g["noise_iters"] == g["batch_iters"][:, None], axis=1)
    J_ok = np.all(np.abs(g["noise_J"] - g["batch_J"][:, None]) <= tol * np.abs(g["batch_J"][:, None]), axis=1)
    X_ok = np.all(g["noise_Xerr"] <= tol, axis=1)
    return it_ok & J_ok & X_ok


@pytest.mark.parametrize("tag", ["TwoLink_T100", "ThreeLink_T100", "ThreeLink_T30"])
def test_default_problem(golden, tag):
    g = golden(tag)
    algo = make_algo(str(g["model"]), int(g["T"]))
    res = algo.solve_batch(g["default_x0"][None], g["default_target"][None], trace=True).to_host()
    n = int(g["default_iters"])
    assert int(res.iterations[0]) == n and int(res.converged[0]) == 1
    assert np.array_equal(res.alpha_trace[0, :n], g["default_alphas"])
    assert np.allclose(res.obj_fun_trace[0, :n], g["default_Js"], rtol=1e-9, atol=0)
    assert abs(res.obj_fun_value[0] - float(g["default_J"])) <= 1e-9 * float(g["default_J"])
    assert rel(res.trajectory[0], g["default_X"]) <= 1e-9


@pytest.mark.parametrize("tag", ["TwoLink_T100", "ThreeLink_T100", "ThreeLink_T30"])
def test_seeded_batch(golden, tag):
    g = golden(tag)
    algo = make_algo(str(g["model"]), int(g["T"]))
    res = algo.solve_batch(g["batch_x0"], g["batch_target"], trace=True).to_host()
    stable = stable_mask(g)
    assert stable.sum() >= 0.7 * len(stable)
    matched = 0
    for b in range(len(stable)):
        n = int(g["batch_iters"][b])
        same = (int(res.iterations[b]) == n and np.array_equal(res.alpha_trace[b, :n], g["batch_alphas"][b, :n]))
        if stable[b]:
            assert same, "problem %d: iteration trace differs on a noise-stable problem" % b
            assert abs(res.obj_fun_value[b] - g["batch_J"][b]) <= 1e-9 * abs(g["batch_J"][b])
            assert rel(res.trajectory[b], g["batch_X"][b]) <= 1e-9
        else:   # knife-edge problem (the reference itself moves under 1e-15 noise): the solve must
            # terminate normally at a cost no worse than the reference's own worst outcome
            hi = max(g["noise_J"][b].max(), g["batch_J"][b])
            assert int(res.converged[b]) == 1 and res.obj_fun_value[b] <= hi * (1 + 1e-3)
        matched += same
    assert matched >= stable.sum()


@pytest.mark.parametrize("tag", ["RoboticArm_T100", "RoboticArm_T200"])
def test_robotic_arm_default_problem(golden, tag):
    """RoboticArm is reference-unstable.  T = 100 default: the reference's 6 iterations and accepted
    steps, cost within 4x the reference's own 1e-15-noise spread; T = 200 default (the reference's own iteration count is
    5 ... 11 under that noise): an iteration count the reference shows, or a cost inside its spread."""
    g = golden(tag)
    algo = make_algo(str(g["model"]), int(g["T"]))
    res = algo.solve_batch(g["default_x0"][None], g["default_target"][None], trace=True).to_host()
    Jref = float(g["default_J"])
    spread = max(np.abs(g["default_noise_J"] - Jref).max() / Jref, 1e-9)
    iters_seen = set(g["default_noise_iters"].tolist()) | {int(g["default_iters"])}
    assert int(res.iterations[0]) in iters_seen or abs(res.obj_fun_value[0] - Jref) <= 4 * spread * Jref
    if int(g["T"]) == 100:
        n = int(g["default_iters"])
        assert int(res.iterations[0]) == n and np.array_equal(res.alpha_trace[0, :n], g["default_alphas"])
        assert abs(res.obj_fun_value[0] - Jref) <= max(4 * spread, 2e-6) * Jref


@pytest.mark.parametrize("tag,min_stable", [("RoboticArm_T100", 11), ("RoboticArm_T200", 0)])
def test_robotic_arm_envelope_rule(golden, tag, min_stable):
    """SURVEY.md section 4.3, PER PROBLEM, on the seeded RoboticArm batches (BASELINE config 3 is RoboticArm at T = 200):
    every problem whose reference trace survives 24 rounding-sized perturbations must be reproduced with the same
    iteration count and accepted steps and a cost within max(1e-9, 2 x the reference's own radius); every other problem
    must show one of the reference's 25 traces or a cost inside the reference's envelope (tests/envelope_rule.py;
    fixtures from the unmodified reference: oracle/make_golden_envelope.py)."""
    import envelope_rule as er
    g, e = golden(tag), golden("Envelope_" + tag)
    env = er.Envelope(g, e)
    algo = make_algo(str(g["model"]), int(g["T"]))
    res = algo.solve_batch(g["batch_x0"], g["batch_target"], trace=True).to_host()
    assert np.all(np.isfinite(res.obj_fun_value)) and np.all(res.converged == 1)
    J0 = algo._dev.rollout(g["batch_x0"], g["batch_target"])[1].cpu().numpy()
    assert np.all(res.obj_fun_value < J0)
    verdicts = [env.judge(b, res.iterations[b], res.alpha_trace[b], res.obj_fun_value[b]) for b in range(env.B)]
    s = er.summarise(verdicts)
    print(tag, s)
    assert s["trace_stable"] >= min_stable
    assert s["passed"] == s["problems"], (s, [v for v in verdicts if not v["ok"]])


@pytest.mark.parametrize("tag,min_subset", [("TwoLink_T100", 8), ("ThreeLink_T100", 6), ("RoboticArm_T100", 2)])
def test_float32_final_cost(golden, tag, min_subset):
    """north_star FP32 bar: final cost <= 1e-4 relative to the FP64 reference on every problem the reference itself
    reproduces to 1e-4 under float32-sized noise on the inputs / outputs of its backward pass and decides clear of the
    stopping criterion (tests/envelope_rule.py::f32_reproducible; fixtures from the unmodified reference); >= 90 % of
    all noise-stable problems; everything terminates."""
    import envelope_rule as er
    g, e = golden(tag), golden("Envelope32_" + tag)
    stable, subset = er.f32_reproducible(g, e)
    assert subset.sum() >= min_subset
    algo = make_algo(str(g["model"]), int(g["T"]), dtype="float32")
    res = algo.solve_batch(g["batch_x0"], g["batch_target"]).to_host()
    err = np.abs(res.obj_fun_value - g["batch_J"]) / np.abs(g["batch_J"])
    assert np.all(err[subset] <= 1e-4), err[subset]
    assert (err[stable] <= 1e-4).sum() >= 0.9 * stable.sum(), err[stable]
    assert np.all(res.converged == 1)


def test_float32_is_refused_where_no_claim_can_be_made():
    """RoboticArm beyond T = 100 is chaotic for the FP64 reference itself (0 of 8 seeded problems reproduce under 1e-15
    noise at T = 200): no FP32 accuracy statement is possible there, so the mode is not offered."""
    with pytest.raises(Exception, match="float32"):
        make_algo("RoboticArmTracking", 200, dtype="float32")


def test_reference_call_sequence(golden):
    """Example/ThreeLinkPlanarManipulatorInteractive.py:8-30: solve, retarget, reset, re-solve; and quirk Q7."""
    from curiousrl_b200 import logger
    g = golden("ThreeLink_T100")
    from curiousrl_b200.scenario.dynamic_model import ThreeLinkPlanarManipulator
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    logger.set_level(45)
    scenario = ThreeLinkPlanarManipulator()
    algo = BasiciLQR(max_line_search=10)
    assert algo.init(scenario) is algo
    assert algo.solve() is None
    assert abs(algo.get_obj_fun_value() - float(g["default_J"])) <= 1e-9 * float(g["default_J"])
    assert algo._trajectory.shape == (100, 11, 1)
    traj = np.asarray(logger.read_from_json()["trajectory"])
    assert traj.shape == (100, 11, 1) and rel(traj[:, :, 0], g["default_X"]) <= 1e-9
    # Q7: solving again without resetting the stored cost stops after one rejected iteration
    J_before = algo.get_obj_fun_value()
    algo.solve()
    assert algo.get_obj_fun_value() == J_before and int(algo.last_result.iterations.item()) == 1
    assert rel(algo._trajectory[:, :, 0], g["default_X_init"]) <= 1e-13
    # retarget exactly like onclick()
    b = 3
    ap = algo.get_obj_add_param()
    assert ap is scenario.add_param        # aliasing quirk Q8
    ap[:, 0], ap[:, 1] = g["batch_target"][b]
    algo.set_obj_add_param(ap)
    algo.set_obj_fun_value(np.inf)
    algo.set_init_state(g["batch_x0"][b].reshape(-1, 1))
    algo.solve()
    assert int(algo.last_result.iterations.item()) == int(g["batch_iters"][b])
    assert abs(algo.get_obj_fun_value() - g["batch_J"][b]) <= 1e-9 * g["batch_J"][b]


def test_forward_backward_methods(golden):
    """The reference's per-iteration entry points on NumPy arrays of the reference's shapes."""
    g = golden("TwoLink_T100")
    algo = make_algo("TwoLinkPlanarManipulator", 100)
    X = algo.dynamic_model.eval_traj()
    assert X.shape == (100, 8, 1) and rel(X[:, :, 0], g["default_X_init"]) <= 1e-14
    C = algo.obj_fun.eval_hessian_obj_fun(X)
    c = algo.obj_fun.eval_grad_obj_fun(X)
    F = algo.dynamic_model.eval_grad_dynamic_model(X)
    assert C.shape == (100, 8, 8) and c.shape == (100, 8, 1) and F.shape == (100, 6, 8)
    algo.set_obj_fun_value(np.inf)
    Js = []
    for i in range(int(g["default_iters"])):
        K, k = algo.backward_pass(C, c, F)
        assert K.shape == (100, 2, 6) and k.shape == (100, 2, 1)
        X, C, c, F, J, stop = algo.forward_pass(X, K, k)
        Js.append(J)
    assert stop and np.allclose(Js, g["default_Js"], rtol=1e-9, atol=0)
    assert rel(X[:, :, 0], g["default_X"]) <= 1e-9


def test_fixed_work_mode_and_options(golden):
    g = golden("ThreeLink_T30")
    algo = make_algo("ThreeLinkPlanarManipulator", 30, is_check_stop=False, max_iter=12)
    res = algo.solve_batch(g["batch_x0"], g["batch_target"], trace=True).to_host()
    assert np.all(res.iterations == 12) and np.all(res.converged == 0)
    for b in range(len(g["batch_iters"])):       # cost never increases; after convergence every step is rejected
        Js = res.obj_fun_trace[b]
        assert np.all(np.diff(Js) <= 0)
    none = make_algo("ThreeLinkPlanarManipulator", 30, line_search_method="none", max_iter=3, is_check_stop=False)
    r2 = none.solve_batch(g["batch_x0"], g["batch_target"], trace=True).to_host()
    assert np.all(r2.alpha_trace[:, :3] == 0)


def test_full_size_properties():
    """BASELINE config 2 (ThreeLink, B = 65,536, T = 100, FP64): size-independent properties."""
    from curiousrl_b200.utils.synthetic import synthetic_batch
    B = 65536
    x0, tgt = synthetic_batch("ThreeLinkPlanarManipulator", B, seed=1234)
    algo = make_algo("ThreeLinkPlanarManipulator", 100)
    res = algo.solve_batch(x0, tgt)
    J0 = algo._dev.rollout(x0, tgt)[1]
    # a few in 10^4 problems run into max_iter (observed 0.05 %); everything else stops by the criterion
    assert float(res.converged.double().mean()) >= 0.995 and torch.all(res.iterations >= 1)
    assert torch.all(res.iterations[res.converged == 0] == 1000)
    # (iteration 0 is accepted unconditionally, so a final cost above the initial one is possible but rare)
    assert torch.all(torch.isfinite(res.obj_fun_value)) and float((res.obj_fun_value < J0).double().mean()) >= 0.99
    # returned trajectories are dynamically consistent: re-rolling their actions reproduces them,
    # and their cost is the reported cost
    X = res.trajectory
    Xr, Jr = algo._dev.rollout(x0, tgt, X[:, :, 8:].contiguous())
    assert torch.allclose(Xr, X, rtol=0, atol=1e-9)
    assert torch.allclose(algo._dev.cost(X, tgt), res.obj_fun_value, rtol=1e-12, atol=0)
    assert torch.all(X[:, -1, 8:] == 0) and float(X[:, :, 8:].abs().max()) <= 100.0
    # shard invariance: any split of the batch gives bit-identical per-trajectory results
    lo, hi = 1000, 1000 + 4096
    part = algo.solve_batch(x0[lo:hi], tgt[lo:hi])
    assert torch.equal(part.obj_fun_value, res.obj_fun_value[lo:hi])
    assert torch.equal(part.iterations, res.iterations[lo:hi])
    assert torch.equal(part.trajectory, res.trajectory[lo:hi])
    # matches the seeded reference statistics (median 14 iterations)
    med = float(res.iterations.double().median())
    assert 10 <= med <= 20


@pytest.mark.parametrize("model,B,T", [("ThreeLinkPlanarManipulator", 4096, 100), ("TwoLinkPlanarManipulator", 2048, 60),
                                       ("ThreeLinkPlanarManipulator", 5, 100), ("RoboticArmTracking", 2048, 120),
                                       ("RoboticArmTracking", 3, 200)])
def test_every_scheduling_path_gives_the_same_bits(model, B, T):
    """Results must not depend on which path finished a trajectory: lane-per-trajectory kernel only,
    lane kernel with the cooperative tail (default and 'everything after the queue drains'),
    free-running CTAs, and the cooperative engine as the whole solve - all bit-identical."""
    from curiousrl_b200.utils.synthetic import synthetic_batch
    x0, tgt = synthetic_batch(model, B, seed=3)
    ref = make_algo(model, T, verify=False, coop_threshold=-1).solve_batch(x0, tgt, trace=True)
    assert int((ref.iterations > 50).sum()) > 0 or B < 100          # the batch does have a tail
    # occupancy=0 picks the configuration by batch size (all-engine below the resident lane count), so the hybrid
    # paths are asked for explicitly (12 = phase-locked lane kernel + cooperative tail)
    for kw in (dict(), dict(occupancy=12), dict(occupancy=12, coop_threshold=32), dict(occupancy=12, coop_threshold=2),
               dict(occupancy=2), dict(occupancy=2, coop_threshold=6), dict(occupancy=20), dict(occupancy=21)):
        res = make_algo(model, T, verify=False, **kw).solve_batch(x0, tgt, trace=True)
        for f in ("iterations", "converged", "obj_fun_value", "trajectory", "alpha_trace"):
            assert torch.equal(getattr(res, f), getattr(ref, f)), "%s differs with %s" % (f, kw)
        a, b = res.obj_fun_trace, ref.obj_fun_trace
        assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


def test_cooperative_paths_fixed_work_and_line_search_modes():
    """Cooperative engine with is_check_stop=False (carry of rejected iterations), more step sizes than
    lanes per trajectory (several passes), line_search_method none / feasibility, vanilla stopping."""
    from curiousrl_b200.utils.synthetic import synthetic_batch
    model, T = "ThreeLinkPlanarManipulator", 40
    x0, tgt = synthetic_batch(model, 96, seed=5)
    for kw in (dict(is_check_stop=False, max_iter=25), dict(max_line_search=40, max_iter=60),
               dict(line_search_method="none", max_iter=15), dict(line_search_method="feasibility"),
               dict(stopping_method="vanilla", stopping_criterion=1e-3), dict(max_line_search=0, max_iter=5)):
        ref = make_algo(model, T, verify=False, coop_threshold=-1, **kw).solve_batch(x0, tgt, trace=True)
        for extra in (dict(), dict(occupancy=12, coop_threshold=32), dict(occupancy=21)):
            res = make_algo(model, T, verify=False, **kw, **extra).solve_batch(x0, tgt, trace=True)
            for f in ("iterations", "converged", "obj_fun_value", "trajectory", "alpha_trace"):
                assert torch.equal(getattr(res, f), getattr(ref, f)), "%s differs with %s %s" % (f, kw, extra)


@pytest.mark.parametrize("model", ["ThreeLinkPlanarManipulator", "TwoLinkPlanarManipulator", "RoboticArmTracking"])
@pytest.mark.parametrize("T", [1, 2, 3, 33])
def test_scheduling_paths_at_tiny_horizons(model, T):
    """Degenerate horizons (T = 1: no action is ever applied) through every path."""
    from curiousrl_b200.utils.synthetic import synthetic_batch
    x0, tgt = synthetic_batch(model, 70, seed=11)
    ref = make_algo(model, T, verify=False, coop_threshold=-1, max_iter=30).solve_batch(x0, tgt, trace=True)
    for kw in (dict(), dict(occupancy=12), dict(occupancy=12, coop_threshold=32), dict(occupancy=20)):
        res = make_algo(model, T, verify=False, max_iter=30, **kw).solve_batch(x0, tgt, trace=True)
        for f in ("iterations", "converged", "obj_fun_value", "trajectory", "alpha_trace"):
            assert torch.equal(getattr(res, f), getattr(ref, f)), "%s differs with %s at T=%d" % (f, kw, T)


def test_recorded_solve_equals_the_fused_one(golden):
    """`logger.set_is_save_json(True)`: one record per iteration like the reference (basic_ilqr.py:93), replayable with
    `read_from_json(folder, no_iter)` (Logger.py:82-99) - and the same final numbers, bit for bit, as the fused solve."""
    from curiousrl_b200 import logger
    g = golden("ThreeLink_T100")
    fused = make_algo("ThreeLinkPlanarManipulator", 100)
    fused.solve()
    iters = int(fused.last_result.iterations.item())
    assert iters == int(g["default_iters"])
    logger.set_folder_name("test_recorded_solve", remove_existing_folder=True).set_is_use_logger(True).set_is_save_json(True)
    try:
        rec = make_algo("ThreeLinkPlanarManipulator", 100)
        rec.solve()
        assert rec.get_obj_fun_value() == fused.get_obj_fun_value()
        assert np.array_equal(rec._trajectory, fused._trajectory)
        last = np.asarray(logger.read_from_json("test_recorded_solve", -1)["trajectory"])
        assert np.array_equal(last, fused._trajectory)
        first = np.asarray(logger.read_from_json("test_recorded_solve", 0)["trajectory"])
        assert first.shape == (100, 11, 1) and not np.array_equal(first, last)
        assert np.array_equal(np.asarray(logger.read_from_json("test_recorded_solve", iters - 1)["trajectory"]), last)
        with pytest.raises(Exception):
            logger.read_from_json("test_recorded_solve", iters)
    finally:
        logger.set_is_save_json(False).set_is_use_logger(False)
