"""The batched, binary counterpart of logger.set_is_save_json(True) / scenario.play(folder, no_iter)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
from curiousrl_b200.scenario.dynamic_model import TwoLinkPlanarManipulator  # noqa: E402
from curiousrl_b200.utils.result_stream import ResultStream  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

logger.set_level(logger.WARN)
logger.set_folder_name("example_stream").set_is_save_json(True)
scenario = TwoLinkPlanarManipulator()
algo = BasiciLQR(max_line_search=10).init(scenario)
x0, targets = synthetic_batch(scenario.name, 4096, seed=1)
first = algo.solve_batch(x0, targets)                                   # record 0
algo.resolve_batch(first, at_step=25, add_param=targets, warm_start=True)   # record 1
stream = ResultStream("example_stream")
print("%d records under %s" % (len(stream), stream.path))
rec = logger.read_batch(no_iter=0)                                      # memory-mapped arrays
print("record 0: trajectory %s, iterations %d..%d" % (rec["trajectory"].shape, rec["iterations"].min(), rec["iterations"].max()))
one = stream.trajectory_as_json(no_iter=-1, index=7)                    # what the reference's read_from_json returns
print("record 1, problem 7 in the reference's encoding: [T][d][1] = %s" % (np.asarray(one["trajectory"]).shape,))
