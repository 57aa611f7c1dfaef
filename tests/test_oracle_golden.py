"""Pin the CPU oracle (oracle/ilqr_numpy.py) to vectors produced by the unmodified reference.

Fixtures: tests/golden/*.npz, written by oracle/make_golden.py in the build
container from /root/reference (see that script's docstring).  The oracle runs
the same sympy -> lambdify -> NumPy/LAPACK substrate as the reference, so the
expectation here is *bit equality* stage by stage, and identical iteration
traces end to end.
"""
import numpy as np
import pytest

from conftest import make_problem
from oracle import ilqr_numpy as orc

TAGS = ["TwoLink_T100", "ThreeLink_T100", "RoboticArm_T100", "RoboticArm_T200", "ThreeLink_T30"]


def _prob(g):
    return make_problem(str(g["model"]), int(g["T"]))


@pytest.mark.parametrize("tag", TAGS)
def test_initial_rollout_and_cost(golden, tag):
    g = golden(tag)
    prob = _prob(g)
    X = orc.rollout(prob, g["default_x0"])
    assert np.array_equal(X, g["default_X_init"])
    # J of the candidate trajectories recorded by the reference
    P = np.broadcast_to(g["default_target"], (prob.T, 2))
    for it in g["stage_iters"]:
        for Xc, Jc in zip(g["stage%d_cand_X" % it], g["stage%d_cand_J" % it]):
            assert orc.cost(prob, Xc, P) == Jc


@pytest.mark.parametrize("tag", TAGS)
def test_stage_derivs_backward_update(golden, tag):
    g = golden(tag)
    prob = _prob(g)
    P = np.broadcast_to(g["default_target"], (prob.T, 2))
    for it in g["stage_iters"]:
        pre = "stage%d_" % it
        X = g[pre + "X"]
        F, c, C = orc.derivs(prob, X, P)
        assert np.array_equal(F, g[pre + "F"])
        assert np.array_equal(c, g[pre + "c"])
        assert np.array_equal(C, g[pre + "C"])
        K, k = orc.backward(prob, F, c, C)
        assert np.array_equal(K, g[pre + "K"])
        assert np.array_equal(k, g[pre + "k"])
        for alpha, Xc in zip(g[pre + "cand_alpha"], g[pre + "cand_X"]):
            assert np.array_equal(orc.update_traj(prob, X, K, k, alpha), Xc)
        Xn, Jn, a_idx, n_roll, _ = orc.line_search(prob, X, K, k, float(g[pre + "J_last"]), P)
        assert np.array_equal(Xn, g[pre + "X_next"]) and Jn == float(g[pre + "J_next"])
        assert n_roll == len(g[pre + "cand_alpha"])
        assert orc.is_stop(Jn, float(g[pre + "J_last"])) == bool(g[pre + "stop"])


def test_structural_known_answers(golden):
    """K[T-1] = 0, k[T-1] = -0, last action 0, diagonal of C."""
    g = golden("ThreeLink_T100")
    assert np.all(g["stage1_K"][-1] == 0) and np.all(g["stage1_k"][-1] == 0)
    assert np.all(g["default_X"][-1, 8:] == 0)
    diag = np.array([1e-100, 20, 1e-100, 20, 1e-100, 20, 2e4, 2e4, 2, 2, 2])
    assert np.array_equal(np.diagonal(g["stage0_C"][0]), diag)
    assert float(g["default_J"]) == 1610256.517185301 and int(g["default_iters"]) == 11
    assert float(golden("TwoLink_T100")["default_J"]) == 501004.8775924033
    assert float(golden("RoboticArm_T100")["default_J"]) == 426981.6859563185
    assert float(golden("RoboticArm_T200")["default_J"]) == 786029.087571384


@pytest.mark.parametrize("tag", TAGS)
def test_default_problem_end_to_end(golden, tag):
    g = golden(tag)
    prob = _prob(g)
    r = orc.solve(prob, g["default_x0"], g["default_target"], max_line_search=int(g["max_line_search"]))
    assert r["iters"] == int(g["default_iters"])
    assert np.array_equal(r["Js"], g["default_Js"])
    assert np.array_equal(r["alphas"], g["default_alphas"])
    assert np.array_equal(r["X"], g["default_X"])
    assert r["rollouts"] == int(g["default_rollouts"].sum())


@pytest.mark.parametrize("tag", ["ThreeLink_T30", "RoboticArm_T200", "TwoLink_T100"])
def test_seeded_batch_end_to_end(golden, tag):
    g = golden(tag)
    prob = _prob(g)
    # keep the CPU suite short: the cheapest few problems of each batch
    order = np.argsort(g["batch_iters"])[:4]
    for b in order:
        r = orc.solve(prob, g["batch_x0"][b], g["batch_target"][b], max_line_search=int(g["max_line_search"]))
        assert r["iters"] == int(g["batch_iters"][b])
        assert r["J"] == float(g["batch_J"][b])
        assert np.array_equal(r["X"], g["batch_X"][b])
        assert np.array_equal(r["alphas"], g["batch_alphas"][b][:r["iters"]])


# ----------------------------------------------------------------------------- LogBarrieriLQR
BARRIER_TAGS = [("LogBarrier_TwoLink_T100", "TwoLinkPlanarManipulator"), ("LogBarrier_ThreeLink_T100", "ThreeLinkPlanarManipulator"),
                ("LogBarrier_RoboticArm_T100", "RoboticArmTracking")]
_BARRIER_PROBLEMS = {}


def barrier_problems(model, T, t0=0.5):
    """(Problem of the barrier cost, Problem of the plain cost), both on the reference's substrate."""
    from conftest import make_scenario
    key = (model, T)
    if key not in _BARRIER_PROBLEMS:
        numba = model == "RoboticArmTracking"
        sc = make_scenario(model, T)
        _BARRIER_PROBLEMS[key] = (orc.Problem(orc.BarrierScenario(sc, t0), use_numba=numba), make_problem(model, T))
    return _BARRIER_PROBLEMS[key]


@pytest.mark.parametrize("tag,model", BARRIER_TAGS)
def test_log_barrier_cost_derivatives(golden, tag, model):
    """Barrier cost, gradient and Hessian diagonal at the initial trajectory, t = t[0]: bit-equal to the reference."""
    g = golden(tag)
    pb, _ = barrier_problems(model, 100)
    X = orc.rollout(pb)
    assert np.array_equal(X, g["init_X"])
    assert orc.cost(pb, X) == float(g["init_J"])
    _F, c, C = orc.derivs(pb, X)
    assert np.array_equal(c, g["init_c"])
    assert np.array_equal(np.stack([np.diag(Ct) for Ct in C]), g["init_Cdiag"])


@pytest.mark.parametrize("tag,model", BARRIER_TAGS)
def test_log_barrier_default_solve(golden, tag, model):
    """LogBarrieriLQR.solve on the scenario's default problem: iterations of every inner loop, step indices,
    barrier and real cost per iteration and the final trajectory - all as the reference produced them."""
    g = golden(tag)
    pb, pr = barrier_problems(model, 100)
    res = orc.solve_log_barrier(pb, pr, t_list=tuple(g["t_list"]), max_line_search=int(g["max_line_search"]))
    assert list(res["inner_iters"]) == list(g["def_inner_iters"])
    assert np.array_equal(res["alphas"], g["def_alpha"])
    assert np.array_equal(res["Js"], g["def_J"])
    assert np.array_equal(res["Jreals"], g["def_Jreal"])
    assert np.array_equal(res["X"], g["def_X"])
