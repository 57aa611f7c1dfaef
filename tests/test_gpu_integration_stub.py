"""The reference-side binding shown in INTEGRATION.md is executable: the two ctypes stubs are extracted from the document,
pointed at the in-tree library and compared with the package's own solvers (bit for bit - same kernels)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, make_scenario

pytestmark = pytest.mark.gpu


def _stub_namespace():
    from curiousrl_b200 import build
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = [b for b in blocks if "def solve_on_b200" in b or "def log_barrier_on_b200" in b or "def jacobians_on_b200" in b]
    assert len(code) == 3
    ns = {}
    exec(code[0].replace('ctypes.CDLL("libcuriousrl_b200.so")', "ctypes.CDLL(%r)" % build.LIB_PATH), ns)
    exec(code[1], ns)
    exec(code[2], ns)
    return ns


def test_jacobian_stub_equals_the_package():
    """`jacobians_on_b200`: the override of NNiLQRDynamicModel.eval_grad_dynamic_model shown in the guide, on the
    reference-shaped network, against the package's own `NNiLQRDynamicModel` - the same kernels, the same bits."""
    from curiousrl_b200.algorithm.ilqr_solver import LargeNetwork, NNiLQRDynamicModel
    ns = _stub_namespace()
    torch.manual_seed(1)
    net = LargeNetwork(11, 8).cuda()
    net.train()
    with torch.no_grad():
        net(torch.randn(256, 11, device="cuda"))
    net.eval()
    rng = np.random.default_rng(0)
    traj = rng.normal(size=(100, 11, 1))
    F = ns["jacobians_on_b200"](net, traj)
    model = NNiLQRDynamicModel(net, np.zeros((8, 1)), np.zeros((100, 3, 1)))
    assert F.shape == (100, 8, 11) and np.array_equal(F, model.eval_grad_dynamic_model(traj))


def test_stubs_of_the_integration_guide_run_and_agree_with_the_package():
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR
    from curiousrl_b200.utils.synthetic import synthetic_batch
    logger.set_level(45)
    ns = _stub_namespace()
    sc = make_scenario("ThreeLinkPlanarManipulator", 30)
    x0, tgt = synthetic_batch(sc.name, 48, seed=2)
    x0d, tgd = torch.as_tensor(x0).cuda(), torch.as_tensor(tgt).cuda()
    opts = ns["SolverOpts"](1000, 1, 1e-6, 10, 0.5, 0, 0, 0, 0, 0, 0)
    X, J, it, ok = ns["solve_on_b200"](sc, x0d, tgd, opts)
    ref = BasiciLQR(max_line_search=10, verify=False).init(sc).solve_batch(x0d, tgd)
    torch.cuda.synchronize()
    assert torch.equal(X, ref.trajectory) and torch.equal(J, ref.obj_fun_value) and torch.equal(it, ref.iterations)
    assert torch.equal(ok, ref.converged)
    # LogBarrieriLQR: the whole schedule through crl_solve_stages
    t_list = [0.5, 1., 2., 5., 10., 20., 50., 100.]
    bopts = ns["SolverOpts"](1000, 1, 1e-6, 10, 0.5, 1, 0, 0, 0, 0, 0)          # line_search_method = feasibility
    Xb, Jb, inner = ns["log_barrier_on_b200"](sc, x0d, tgd, bopts, t_list)
    bref = LogBarrieriLQR(max_line_search=10, verify=False).init(sc).solve_batch(x0d, tgd)
    torch.cuda.synchronize()
    assert torch.equal(inner, bref.inner_iterations) and torch.equal(Xb, bref.trajectory) and torch.equal(Jb, bref.obj_fun_value)
