"""One small invocation of the persistent solve kernel, for ncu captures.

    python profiles/run_solve.py [--model M] [--batch B] [--horizon T] [--iters I] [--dtype D] [--converge]

Default: ThreeLink, B = 65,536, T = 100, FP64, fixed work (is_check_stop=False, max_iter = I).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
import curiousrl_b200.scenario.dynamic_model as models  # noqa: E402
from curiousrl_b200.utils.synthetic import ARM_GEOMETRY, synthetic_batch, synthetic_init_states  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ThreeLinkPlanarManipulator")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=100)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--dtype", default="float64")
ap.add_argument("--converge", action="store_true")
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--occupancy", type=int, default=0)
a = ap.parse_args()

logger.set_level(45)
kw = dict(max_line_search=10, dtype=a.dtype, verify=False, occupancy=a.occupancy)
if not a.converge:
    kw.update(is_check_stop=False, max_iter=a.iters)
algo = BasiciLQR(**kw).init(getattr(models, a.model)(T=a.horizon))
if a.model in ARM_GEOMETRY:
    x0, tgt = synthetic_batch(a.model, a.batch, seed=1234)
    tgt = algo._dev.real(tgt)
else:                                  # generated-model scenarios: perturbed initial states, the scenario's own add_param
    x0, tgt = synthetic_init_states(getattr(models, a.model)(T=a.horizon), a.batch, seed=1234), None
x0 = algo._dev.real(x0)
for _ in range(a.repeat):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    res = algo.solve_batch(x0, tgt)
    e.record()
    torch.cuda.synchronize()
    iters = int(res.iterations.sum())
    print("solve: %.3f ms, %d traj-iters, %.3f M traj-iter/s" % (s.elapsed_time(e), iters, iters / s.elapsed_time(e) / 1e3))
