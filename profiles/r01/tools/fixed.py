import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from curiousrl_b200 import logger
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
import curiousrl_b200.scenario.dynamic_model as models
from curiousrl_b200.utils.synthetic import synthetic_batch
logger.set_level(45)
MODEL = "ThreeLinkPlanarManipulator"
NAMES = ["refill", "backward", "linesearch", "retire", "wait", "gather", "c_backward", "c_linesearch", "c_output"]
B = int(os.environ.get("B", 65536))
x0, tgt = synthetic_batch(MODEL, B, seed=1234)
for cfg in sys.argv[1:]:
    coop, occ, iters = (int(v) for v in cfg.split(":"))
    algo = BasiciLQR(max_line_search=10, verify=False, coop_threshold=coop, occupancy=occ, is_check_stop=False, max_iter=iters).init(getattr(models, MODEL)(T=100))
    x0d, tgd = algo._dev.real(x0), algo._dev.real(tgt)
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); res = algo.solve_batch(x0d, tgd, return_trajectory=os.environ.get('RT','1')=='1'); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    tm = algo._dev._workspace[:256].view(torch.int64)[8:8 + len(NAMES)].cpu().numpy().astype(np.float64)
    it = int(res.iterations.sum())
    print("%s B=%d coop=%d variant=%d fixed %d: %.1f ms, %.2f M traj-iter/s | %s" % (os.environ.get("CURIOUSRL_B200_LIB", "")[-14:], B, coop, occ, iters, best, it / best / 1e3,
          ", ".join("%s %.1f" % (n, v / 1184 / 1.965e6) for n, v in zip(NAMES, tm))), flush=True)
