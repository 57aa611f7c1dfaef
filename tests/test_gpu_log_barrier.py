"""LogBarrieriLQR on the B200 against the unmodified reference.

Fixtures: tests/golden/LogBarrier_*.npz, written by oracle/make_golden_barrier.py in the build container
from /root/reference (LogBarrieriLQR(max_line_search=10), the setting of
Example/RoboticArmTrackingIteractive.py:11).  The CUDA path is compared through the reference's own
Python surface: `init(scenario)`, `obj_fun.eval_*`, `solve()`, plus the batched extension `solve_batch`.

Tolerances (FP64): cost / gradient / Hessian of the barrier cost <= 1e-13 relative (CUDA's log is within
1 ulp of glibc's); end to end the SAME iteration count in every inner loop (one per barrier parameter t),
final real cost <= 1e-9 relative and final trajectory <= 1e-8 of its largest entry.
"""
import numpy as np
import pytest
import torch

from conftest import make_scenario

pytestmark = pytest.mark.gpu

TAGS = [("LogBarrier_TwoLink_T100", "TwoLinkPlanarManipulator"), ("LogBarrier_ThreeLink_T100", "ThreeLinkPlanarManipulator"),
        ("LogBarrier_RoboticArm_T100", "RoboticArmTracking")]


def make_algo(model, **kw):
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR
    logger.set_level(45)
    return LogBarrieriLQR(max_line_search=10, **kw).init(make_scenario(model, 100))


@pytest.mark.parametrize("tag,model", TAGS)
def test_barrier_cost_and_derivatives(golden, tag, model):
    g = golden(tag)
    algo = make_algo(model)
    X0 = algo.dynamic_model.eval_traj()
    assert np.allclose(X0[:, :, 0], g["init_X"], rtol=1e-13, atol=1e-13)
    assert algo.obj_fun.eval_obj_fun(X0) == pytest.approx(float(g["init_J"]), rel=1e-13)
    c = algo.obj_fun.eval_grad_obj_fun(X0)[:, :, 0]
    assert np.allclose(c, g["init_c"], rtol=1e-13, atol=1e-13 * np.abs(g["init_c"]).max())
    C = algo.obj_fun.eval_hessian_obj_fun(X0)
    assert np.allclose(np.stack([np.diag(Ct) for Ct in C]), g["init_Cdiag"], rtol=1e-13)
    off = C - np.stack([np.diag(np.diag(Ct)) for Ct in C])
    assert not off.any()                                        # the barrier keeps the Hessian diagonal
    assert algo.get_obj_add_param().shape == (100, 3)           # (target_x, target_y, t), log_barrier_ilqr.py:100-101


@pytest.mark.parametrize("tag,model", TAGS)
def test_default_problem(golden, tag, model):
    """solve(): one inner loop per t; iterations per loop, real cost and trajectory as the reference's."""
    g = golden(tag)
    algo = make_algo(model)
    assert algo.solve() is None
    assert algo.inner_iterations == [int(v) for v in g["def_inner_iters"]]
    Jreal = algo._real_obj_fun.eval_obj_fun(algo._trajectory)
    assert Jreal == pytest.approx(float(g["def_Jreal"][-1]), rel=1e-9)
    X = algo._trajectory[:, :, 0]
    assert algo._trajectory.shape == (100, X.shape[1], 1)
    assert np.abs(X - g["def_X"]).max() <= 1e-8 * np.abs(g["def_X"]).max()
    assert np.all(X[-1, algo.dynamic_model.n:] == 0.0)          # last action stays 0
    # the stored cost was reset after the last converged inner loop (log_barrier_ilqr.py:141)
    assert algo.get_obj_fun_value() == np.inf
    # the last column of add_param holds the last barrier parameter
    assert np.all(algo.get_obj_add_param()[:, -1] == float(g["t_list"][-1]))


@pytest.mark.parametrize("tag,model", TAGS)
def test_seeded_batch(golden, tag, model):
    """solve_batch: every problem of the seeded batch takes the reference's number of iterations in every inner
    loop and ends on the reference's trajectory.  (The reference's own 'real' cost of a re-targeted problem is
    evaluated with the scenario's default target - its _real_obj_fun never sees set_obj_add_param - so the
    trajectories are compared, and the real cost is checked against the fixture with that same default target.)"""
    g = golden(tag)
    algo = make_algo(model)
    res = algo.solve_batch(g["batch_x0"], g["batch_target"])
    assert np.array_equal(res.inner_iterations.cpu().numpy(), g["batch_inner_iters"])
    assert np.array_equal(res.iterations.cpu().numpy(), g["batch_inner_iters"].sum(axis=1))
    X = res.trajectory.cpu().numpy()
    scale = np.abs(g["batch_X"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(X - g["batch_X"]) / scale).max() <= 1e-7
    default_target = np.asarray(make_scenario(model, 100).add_param, dtype=np.float64)[0]
    J_default = algo._real_dev.cost(res.trajectory, algo._real_dev.real(default_target)).cpu().numpy()
    assert np.allclose(J_default, g["batch_Jreal"], rtol=1e-8)
    # the real cost the batched API reports uses each problem's own target and is barrier-free
    own = algo._real_dev.cost(res.trajectory, algo._real_dev.real(g["batch_target"])).cpu().numpy()
    assert np.array_equal(res.real_obj_fun_value.cpu().numpy(), own)


def test_resolve_after_retarget(golden):
    """Example/RoboticArmTrackingIteractive.py:19-31: retarget, reset the stored cost, move the initial state,
    solve again.  The second solve starts with the barrier parameter the first one ended on (the first inner
    loop never writes t, log_barrier_ilqr.py:122) - so its first loop is short."""
    g = golden("LogBarrier_RoboticArm_T100")
    algo = make_algo("RoboticArmTracking")
    algo.solve()
    first = list(algo.inner_iterations)
    ap = algo.get_obj_add_param()
    ap[:, 0], ap[:, 1] = 1.0, -1.5
    algo.set_obj_add_param(ap)
    algo.set_obj_fun_value(np.inf)
    algo.set_init_state(algo._trajectory[5, :algo.dynamic_model.n].reshape(-1, 1))
    algo.solve()
    assert len(algo.inner_iterations) == len(first) and sum(algo.inner_iterations) > 0
    assert np.isfinite(algo._real_obj_fun.eval_obj_fun(algo._trajectory))
    X = algo._trajectory[:, :, 0]
    assert np.all(np.abs(X[:, algo.dynamic_model.n:]) < 500.0)   # strictly inside the action box


def test_unsupported_inputs_raise():
    from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR
    import curiousrl_b200.scenario.dynamic_model as models
    with pytest.raises(Exception):
        LogBarrieriLQR().init(models.RoboticArmTracking(is_with_constraints=False))     # no action box: no barrier model


@pytest.mark.parametrize("model,B", [("ThreeLinkPlanarManipulator", 192), ("RoboticArmTracking", 160), ("TwoLinkPlanarManipulator", 96)])
def test_every_scheduling_path_gives_the_same_bits(model, B):
    """Barrier costs through the lane-per-trajectory kernel only, the lane kernel with the cooperative tail, and the
    cooperative engine as the whole solve (the default at this batch size): bit-identical results."""
    from curiousrl_b200.utils.synthetic import synthetic_batch
    x0, tgt = synthetic_batch(model, B, seed=7)
    ref = make_algo(model, verify=False, occupancy=12, coop_threshold=-1, max_iter=150).solve_batch(x0, tgt, fused=False)
    assert int(ref.inner_iterations[:, 0].max()) > 40
    for kw in (dict(), dict(occupancy=12), dict(occupancy=12, coop_threshold=32), dict(occupancy=20), dict(occupancy=21),
               dict(occupancy=12, coop_threshold=-1)):
        for fused in (True, False):
            res = make_algo(model, verify=False, max_iter=150, **kw).solve_batch(x0, tgt, fused=fused)
            for f in ("inner_iterations", "iterations", "converged", "obj_fun_value", "trajectory", "real_obj_fun_value"):
                assert torch.equal(getattr(res, f), getattr(ref, f)), "%s differs with %s fused=%s" % (f, kw, fused)


@pytest.mark.parametrize("tag,model", TAGS)
def test_fused_schedule_equals_one_launch_per_t(golden, tag, model):
    """crl_solve_stages (the whole barrier schedule inside one persistent kernel) against one crl_solve launch per t:
    identical bits, also when inner loops end by max_iter (no reset of the stored cost) and without the stopping test."""
    g = golden(tag)
    for kw in (dict(), dict(max_iter=3), dict(max_iter=4, is_check_stop=False)):
        a = make_algo(model, verify=False, **kw).solve_batch(g["batch_x0"], g["batch_target"], fused=True)
        b = make_algo(model, verify=False, **kw).solve_batch(g["batch_x0"], g["batch_target"], fused=False)
        for f in ("inner_iterations", "iterations", "converged", "obj_fun_value", "trajectory", "real_obj_fun_value"):
            assert torch.equal(getattr(a, f), getattr(b, f)), "%s differs with %s" % (f, kw)
