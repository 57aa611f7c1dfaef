import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from curiousrl_b200 import logger
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR
import curiousrl_b200.scenario.dynamic_model as models
from curiousrl_b200.utils.synthetic import synthetic_batch
logger.set_level(45)
for model in ("ThreeLinkPlanarManipulator", "RoboticArmTracking"):
    for B in (1, 4, 148 * 8 * 4):
        x0, tgt = synthetic_batch(model, B, seed=1234)
        for name, algo in (("basic", BasiciLQR(max_line_search=10, verify=False, is_check_stop=False, max_iter=300)),
                           ("barrier", LogBarrieriLQR(max_line_search=10, verify=False, is_check_stop=False, max_iter=300, t=[0.5]))):
            algo.init(getattr(models, model)(T=100))
            x0d, tgd = algo._dev.real(x0), algo._dev.real(tgt)
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.time()
                res = algo.solve_batch(x0d, tgd)
                torch.cuda.synchronize(); dt = time.time() - t0
            print("%s %s B=%d: %.1f ms for 300 iterations -> %.3f ms/iteration" % (model, name, B, dt * 1e3, dt * 1e3 / 300))
