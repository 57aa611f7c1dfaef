"""Time line of one persistent solve (development tool; B200 only).

    python profiles/r02/tools/timeline.py [--batch B] [--occupancy V] [--coop C] [--out FILE.csv]

While the solve kernel runs, a side stream copies the kernel's 256-byte control block (work counter, straggler-queue
tail / head, warps that left lane mode, queue trajectories finished) and the per-trajectory iteration counts (written
once, when a trajectory retires) to pinned host memory every ~millisecond.  Prints, per sample: time since launch,
batch trajectories handed out, warps in engine mode, stragglers published / claimed / finished, trajectories finished
in total and the trajectory-iterations those account for.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from curiousrl_b200 import logger  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR  # noqa: E402
from curiousrl_b200.algorithm.ilqr_solver.device import BatchResult  # noqa: E402
import curiousrl_b200.scenario.dynamic_model as models  # noqa: E402
from curiousrl_b200.utils.synthetic import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ThreeLinkPlanarManipulator")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=100)
ap.add_argument("--occupancy", type=int, default=0)
ap.add_argument("--coop", type=int, default=0)
ap.add_argument("--every-ms", type=float, default=2.0)
ap.add_argument("--out", default="")
a = ap.parse_args()

logger.set_level(45)
algo = BasiciLQR(max_line_search=10, verify=False, occupancy=a.occupancy, coop_threshold=a.coop).init(
    getattr(models, a.model)(T=a.horizon))
dev = algo._dev
x0, tgt = synthetic_batch(a.model, a.batch, seed=1234)
x0d, tgd = dev.real(x0), dev.real(tgt)
B = a.batch
algo.solve_batch(x0d, tgd)                                   # warm-up (allocates the workspace)
torch.cuda.synchronize()

out = BatchResult(dev.empty(B, a.horizon, dev.d), dev.empty(B, dtype=torch.float64),
                  torch.zeros(B, dtype=torch.int32, device=dev.device), dev.empty(B, dtype=torch.uint8))
ws = dev._workspace
side = torch.cuda.Stream()
ctl_h = torch.zeros(32, dtype=torch.int64).pin_memory()
it_h = torch.zeros(B, dtype=torch.int32).pin_memory()
ctl_view = ws[:256].view(torch.int64)
done = torch.cuda.Event()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = algo.solve_batch(x0d, tgd, out=out)         # the public call bench.py times, with the pre-zeroed result object
done.record()
rows = []
while True:
    finished = done.query()
    with torch.cuda.stream(side):
        ctl_h.copy_(ctl_view, non_blocking=True)
        it_h.copy_(out.iterations, non_blocking=True)
    side.synchronize()
    t = (time.perf_counter() - t0) * 1e3
    its = it_h.numpy()
    nz = its > 0
    c = ctl_h.numpy()
    rows.append((t, int(c[0]), int(c[3]), int(c[1]), int(c[2]), int(c[4]), int(nz.sum()), int(its[nz].sum())))
    if finished:
        break
    time.sleep(a.every_ms / 1e3)
torch.cuda.synchronize()
total = int(out.iterations.sum())
print("# %s B=%d occupancy=%d coop=%d: %.1f ms, %d traj-iters" % (a.model, B, a.occupancy, a.coop, rows[-1][0], total))
print("# t_ms next warps_engine q_tail q_head q_fin finished finished_iters")
last = -1e9
for r in rows:
    if r[0] - last >= a.every_ms * 0.999 or r is rows[-1]:
        print("%8.2f %8d %5d %6d %6d %6d %7d %9d" % r)
        last = r[0]
if a.out:
    np.savetxt(a.out, np.array(rows), fmt="%.3f", header="t_ms next warps_engine q_tail q_head q_fin finished finished_iters")
