"""Run-time compiled device models (curiousrl_b200/jit.py): a scenario the package has never seen gets a library of its own
at `init` - generated header + nvcc, same C ABI.  CPU part: the library builds, loads, exports the iLQR part of the ABI
and reports the scenario's sizes and box; GPU part: BasiciLQR / LogBarrieriLQR on it against the CPU oracle, which
lambdifies the same sympy expressions the way the reference does (basic_ilqr.py:50-66 takes any DynamicModelWrapper)."""
import ctypes
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_codegen import ToyPendulumCart  # noqa: E402

from curiousrl_b200 import _native, build  # noqa: E402

needs_nvcc = pytest.mark.skipif(build.find_nvcc() is None, reason="nvcc not available")


@needs_nvcc
def test_library_of_an_unseen_scenario_builds_and_reports_it():
    from curiousrl_b200 import jit
    scenario = ToyPendulumCart(T=12)
    lib, path = jit.compile_scenario(scenario)
    assert os.path.exists(path) and os.path.dirname(os.path.dirname(path)) == jit.JIT_DIR
    assert lib.crl_abi_version() == _native.ABI_VERSION
    for name in _native.EXPORTS:
        if not name.startswith(("crl_mlp", "crl_smooth")):
            assert hasattr(lib, name), name
    n, m, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.crl_model_dims(jit.USER_MODEL_ID, n, m, p) == 0
    assert (n.value, m.value) == (scenario.n, scenario.m) and p.value == np.asarray(scenario.add_param).shape[1]
    assert lib.crl_model_dims(jit.USER_BARRIER_MODEL_ID, n, m, p) == 0 and p.value == np.asarray(scenario.add_param).shape[1] + 2
    assert lib.crl_model_dims(_native.GENERIC_MODEL_IDS["VehicleTracking"], n, m, p) == -2          # not in THIS library
    d = scenario.n + scenario.m
    lo, hi = (ctypes.c_double * d)(), (ctypes.c_double * d)()
    assert lib.crl_model_bounds(jit.USER_MODEL_ID, lo, hi) == 0
    assert np.array_equal(np.array([lo[:], hi[:]]).T, np.asarray(scenario.constr, dtype=np.float64))
    # a second request is served from the cache (same object)
    assert jit.compile_scenario(ToyPendulumCart(T=12))[0] is lib
    # no compute without a GPU: a host pointer is refused before any launch
    prob = _native.Problem(jit.USER_MODEL_ID, 0, 1, 12, 0, 0, 1, 0.0, 0.0, None)
    buf = (ctypes.c_double * 64)()
    assert lib.crl_cost(ctypes.byref(prob), buf, buf, None) != 0


@needs_nvcc
def test_edited_scenario_is_a_different_library():
    from curiousrl_b200 import jit

    class ToyPendulumCart2(ToyPendulumCart):
        pass

    a = jit.compile_scenario(ToyPendulumCart(T=12))[1]
    longer = ToyPendulumCart(T=20)                      # same expressions, other horizon: same library (T is a run-time argument)
    assert jit.compile_scenario(longer)[1] == a


@pytest.mark.gpu
def test_basic_ilqr_on_an_unseen_scenario_against_the_oracle():
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    from oracle import ilqr_numpy as orc
    logger.set_level(45)
    scenario = ToyPendulumCart(T=30)
    algo = BasiciLQR(max_line_search=10, max_iter=60).init(scenario)
    assert algo._dev.jit and algo._dev.model_id == 100
    algo.solve()
    prob = orc.Problem(scenario)
    ref = orc.solve(prob, max_line_search=10, max_iter=60)
    res = algo.last_result.to_host()
    assert int(res.iterations[0]) == ref["iters"]
    assert abs(algo.get_obj_fun_value() - ref["J"]) <= 1e-9 * abs(ref["J"])
    assert np.abs(algo._trajectory[:, :, 0] - ref["X"]).max() <= 1e-8
    # stage surface of the same session
    X = algo.dynamic_model.eval_traj()
    F = algo.dynamic_model.eval_grad_dynamic_model(X)
    Fr, cr, Cr = orc.derivs(prob, X[:, :, 0])
    assert np.abs(F - Fr).max() <= 1e-12 * max(1.0, np.abs(Fr).max())
    K, k = algo.backward_pass(algo.obj_fun.eval_hessian_obj_fun(X), algo.obj_fun.eval_grad_obj_fun(X), F)
    Kr, kr = orc.backward(prob, Fr, cr, Cr)
    assert np.abs(K - Kr).max() <= 1e-9 * max(1.0, np.abs(Kr).max())
    # a batch of perturbed initial states: every problem equals its single-problem oracle solve
    rng = np.random.default_rng(0)
    x0 = np.asarray(scenario.init_state, dtype=np.float64).reshape(1, -1) + 0.05 * rng.normal(size=(8, scenario.n))
    out = algo.solve_batch(x0).to_host()
    for b in range(8):
        r = orc.solve(prob, x0[b], None, max_line_search=10, max_iter=60)
        assert int(out.iterations[b]) == r["iters"], b
        assert abs(out.obj_fun_value[b] - r["J"]) <= 1e-9 * abs(r["J"])


@pytest.mark.gpu
def test_log_barrier_on_an_unseen_scenario_against_the_oracle():
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR
    from oracle import ilqr_numpy as orc
    logger.set_level(45)
    scenario = ToyPendulumCart(T=30)
    algo = LogBarrieriLQR(max_line_search=10, max_iter=40).init(scenario)
    algo.solve()
    ref = orc.solve_log_barrier(orc.Problem(orc.BarrierScenario(scenario, 0.5)), orc.Problem(scenario), max_line_search=10,
                                max_iter=40)
    # (the stored cost is +inf after a converged inner loop, log_barrier_ilqr.py:140-141: report the barrier-free cost)
    J_real = algo._real_obj_fun.eval_obj_fun(algo._trajectory)
    assert abs(J_real - ref["Jreal"]) <= 1e-7 * abs(ref["Jreal"])
    assert np.array_equal(np.asarray(algo.inner_iterations), ref["inner_iters"])
    assert np.abs(algo._trajectory[:, :, 0] - ref["X"]).max() <= 1e-6
