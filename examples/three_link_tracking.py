"""Example/ThreeLinkPlanarManipulatorInteractive.py without the GUI: solve, re-target, re-solve - one problem, then a batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
from curiousrl_b200.scenario.dynamic_model import ThreeLinkPlanarManipulator  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

logger.set_level(logger.WARN)
scenario = ThreeLinkPlanarManipulator()
algo = BasiciLQR(max_line_search=10)
algo.init(scenario)
algo.solve()
print("default problem: J = %.10f, end effector at (%.4f, %.4f)" % (
    algo.get_obj_fun_value(), algo._trajectory[-1, 6, 0], algo._trajectory[-1, 7, 0]))

# what onclick() does in the reference example: new target, reset the stored cost, start from the current state
add_param = algo.get_obj_add_param()
add_param[:, 0], add_param[:, 1] = -1.5, 2.5
algo.set_obj_add_param(add_param)
algo.set_obj_fun_value(np.inf)
algo.set_init_state(algo._trajectory[40, :scenario.n].reshape(-1, 1))
algo.solve()
print("re-targeted to (-1.5, 2.5) from step 40: J = %.10f, end effector at (%.4f, %.4f)" % (
    algo.get_obj_fun_value(), algo._trajectory[-1, 6, 0], algo._trajectory[-1, 7, 0]))

# the same for 65,536 independent problems in one persistent kernel
B = int(os.environ.get("EXAMPLE_BATCH", "65536"))
x0, targets = synthetic_batch(scenario.name, B, seed=1)
res = algo.solve_batch(x0, targets)
print("batch of %d: %d trajectory-iterations, %.2f %% converged, median cost %.4f" % (
    B, int(res.iterations.sum()), 100.0 * float(res.converged.float().mean()), float(res.obj_fun_value.median())))
again = algo.resolve_batch(res, at_step=40, add_param=targets[::-1].copy(), warm_start=True)
print("re-solved from step 40 with swapped targets (warm start): %d trajectory-iterations" % int(again.iterations.sum()))
