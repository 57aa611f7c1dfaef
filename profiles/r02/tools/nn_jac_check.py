"""Check and time crl_mlp_jacobian (development tool; B200 only).

    python profiles/r02/tools/nn_jac_check.py [--net large|small] [--points N ...]

Random network in eval mode (BatchNorm with non-trivial running statistics), random rows: the FP32 CUDA-core path
(impl 0) against torch autograd done the way the reference does it (n copies of the row, backward with the identity,
nn_ilqr.py:265-274), the tensor-core path (impl 1) against impl 0, and the time of both.
"""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from curiousrl_b200.algorithm.ilqr_solver.nn_ilqr import FoldedMLP, LargeNetwork, SmallNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="large")
ap.add_argument("--points", type=int, nargs="+", default=[100, 1000, 16384, 409600])
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--m", type=int, default=3)
a = ap.parse_args()

torch.manual_seed(0)
dev = torch.device("cuda", 0)
n, d = a.n, a.n + a.m
net = (LargeNetwork if a.net == "large" else SmallNetwork)(d, n).to(dev)
net.train()
with torch.no_grad():
    for _ in range(3):
        net(torch.randn(512, d, device=dev) * 2.0)          # move the BatchNorm running statistics off (0, 1)
net.eval()
folded = FoldedMLP(net, dev)


def reference_jacobian(x):
    """nn_ilqr.py:265-274, row by row"""
    eye = torch.eye(n, device=dev)
    out = torch.zeros((x.shape[0], n, d), device=dev)
    for t in range(x.shape[0]):
        r = x[t].repeat(n, 1).requires_grad_(True)
        net(r).backward(eye)
        out[t] = r.grad
    return out


for S in a.points:
    x = torch.randn(S, d, device=dev) * 2.0
    y0, j0 = folded.jacobian(x, impl=0) if S <= 20000 else (None, None)
    torch.cuda.synchronize()
    y1, j1 = folded.jacobian(x, impl=1)
    torch.cuda.synchronize()
    line = "S=%d" % S
    if S <= 1000:
        jr = reference_jacobian(x)
        with torch.no_grad():
            yr = net(x)
        line += "  impl0 vs autograd: jac %.2e  y %.2e" % (float((j0 - jr).norm() / jr.norm()), float((y0 - yr).norm() / yr.norm()))
    if j0 is not None:
        line += "  impl1 vs impl0: jac %.2e (max row %.2e)  y %.2e" % (
            float((j1 - j0).norm() / j0.norm()),
            float(((j1 - j0).flatten(1).norm(dim=1) / j0.flatten(1).norm(dim=1).clamp_min(1e-20)).max()),
            float((y1 - y0).norm() / y0.norm()))
    line += "  sha(jac1) %s" % hashlib.sha256(j1.cpu().numpy().tobytes()).hexdigest()[:12]
    for impl in ((0, 1) if S <= 20000 else (1,)):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        folded.jacobian(x, impl=impl)
        t0.record()
        for _ in range(3):
            folded.jacobian(x, impl=impl)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        flops = 2.0 * S * (d * folded.h1 + folded.h1 * folded.h2 + n * folded.h2 + n * folded.h2 * folded.h1 + n * folded.h1 * d)
        line += "  impl%d %.3f ms (%.1f TFLOP/s)" % (impl, ms, flops / ms / 1e9)
    print(line, flush=True)
