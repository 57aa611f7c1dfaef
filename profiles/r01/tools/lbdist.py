import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from curiousrl_b200 import logger
from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR
import curiousrl_b200.scenario.dynamic_model as models
from curiousrl_b200.utils.synthetic import synthetic_batch
logger.set_level(45)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
KW = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
for model in sys.argv[1].split(","):
    sc = getattr(models, model)(T=100)
    algo = LogBarrieriLQR(max_line_search=10, verify=False, **KW).init(sc)
    x0, tgt = synthetic_batch(model, B, seed=1234)
    x0 = torch.as_tensor(x0).cuda(); tgt = torch.as_tensor(tgt).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        res = algo.solve_batch(x0, tgt)
        torch.cuda.synchronize(); dt = time.time() - t0
    inner = res.inner_iterations.cpu().numpy()
    tot = inner.sum(1)
    print(model, KW, "B", B, "time %.3f s" % dt, "total traj-iters", tot.sum())
    print("  total iters pct 50/90/99/99.9/max:", np.percentile(tot, [50, 90, 99, 99.9]).tolist(), tot.max())
    print("  per-stage mean:", inner.mean(0).round(1).tolist())
    print("  per-stage max :", inner.max(0).tolist())
    print("  per-stage #(>=1000):", (inner >= 1000).sum(0).tolist())
    print("  trajectories with total > 1000/2000/4000:", (tot > 1000).sum(), (tot > 2000).sum(), (tot > 4000).sum())
