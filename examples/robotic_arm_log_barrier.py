"""Example/RoboticArmTrackingIteractive.py without the GUI: LogBarrieriLQR on one problem and on a batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR  # noqa: E402
from curiousrl_b200.scenario.dynamic_model import RoboticArmTracking  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

logger.set_level(logger.WARN)
scenario = RoboticArmTracking()
algo = LogBarrieriLQR(max_line_search=10).init(scenario)
algo.solve()
print("default problem: iterations per barrier parameter %s, real cost %.10f" % (
    algo.inner_iterations, algo._real_obj_fun.eval_obj_fun(algo._trajectory)))

B = int(os.environ.get("EXAMPLE_BATCH", "16384"))
x0, targets = synthetic_batch(scenario.name, B, seed=1)
res = algo.solve_batch(x0, targets)          # the whole barrier schedule inside one persistent kernel
print("batch of %d: mean iterations per barrier parameter %s, median real cost %.4f" % (
    B, res.inner_iterations.float().mean(dim=0).cpu().numpy().round(1).tolist(), float(res.real_obj_fun_value.median())))
