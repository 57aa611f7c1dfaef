"""Scenarios whose device models are generated from their sympy expressions (state bounds, time-varying cost weights)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR  # noqa: E402
from curiousrl_b200.scenario.dynamic_model import CartPoleSwingUp1, VehicleTracking  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_init_states  # noqa: E402

logger.set_level(logger.WARN)
for cls in (VehicleTracking, CartPoleSwingUp1):
    scenario = cls()
    algo = BasiciLQR(max_line_search=10, max_iter=100).init(scenario)
    algo.solve()
    print("%s: %d iterations, J = %.10f" % (scenario.name, int(algo.last_result.iterations[0]), algo.get_obj_fun_value()))
    B = int(os.environ.get("EXAMPLE_BATCH", "65536"))
    res = algo.solve_batch(synthetic_init_states(scenario, B, seed=1))
    print("  batch of %d perturbed initial states: %d trajectory-iterations, median cost %.4f" % (
        B, int(res.iterations.sum()), float(res.obj_fun_value.median())))

scenario = VehicleTracking()
barrier = LogBarrieriLQR(max_line_search=10, max_iter=100).init(scenario)
barrier.solve()
print("VehicleTracking under the log barrier: iterations per barrier parameter %s, real cost %.10f" % (
    barrier.inner_iterations, barrier._real_obj_fun.eval_obj_fun(barrier._trajectory)))
