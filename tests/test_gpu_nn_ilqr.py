"""NNiLQR on the GPU: the learned-dynamics Jacobian kernels, the time filter, one iteration against the
CPU restatement (oracle/nn_ilqr_numpy.py) and the solver end to end.

Tolerances.  impl 0 (FP32 CUDA cores) against float32 autograd: 1e-5 relative.  impl 1 (tcgen05, TF32 inputs, FP32
accumulation): the value y to 3e-3; the Jacobian is piecewise constant in x, and a hidden unit whose pre-activation lies
within TF32 rounding of zero may fall on the other side of its kink than in FP32 (measured: 0.04 units per point for a
random LargeNetwork, each moving that point's Jacobian by O(10 %)) - so: at least 95 % of the points within 5e-3 relative,
all of them finite.  With the masks agreeing, the TF32 products alone are within 1e-3 (profiles/r02/nn_jacobian.md)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(cls, d, n, seed, dev):
    torch.manual_seed(seed)
    net = cls(d, n).to(dev)
    net.train()
    with torch.no_grad():
        for _ in range(3):
            net(torch.randn(512, d, device=dev) * 2.0)
    return net.eval()


def _autograd(net, x, n):
    eye = torch.eye(n, device=x.device)
    out = torch.zeros((x.shape[0], n, x.shape[1]), device=x.device)
    for t in range(x.shape[0]):
        r = x[t].repeat(n, 1).requires_grad_(True)
        net(r).backward(eye)
        out[t] = r.grad
    return out


@pytest.mark.parametrize("net_name,d,n,points", [("LargeNetwork", 11, 8, 300), ("SmallNetwork", 8, 6, 200),
                                                 ("LargeNetwork", 6, 4, 37), ("SmallNetwork", 11, 8, 1)])
def test_jacobian_kernels(net_name, d, n, points):
    from curiousrl_b200.algorithm.ilqr_solver import nn_ilqr
    dev = torch.device("cuda", 0)
    net = _net(getattr(nn_ilqr, net_name), d, n, seed=points, dev=dev)
    f = nn_ilqr.FoldedMLP(net, dev)
    x = torch.randn(points, d, device=dev) * 2.0
    with torch.no_grad():
        y_ref = net(x)
    j_ref = _autograd(net, x, n)
    y0, j0 = f.jacobian(x, impl=0)
    assert float((y0 - y_ref).norm() / y_ref.norm()) <= 1e-5
    assert float((j0 - j_ref).norm() / j_ref.norm()) <= 1e-5
    y1, j1 = f.jacobian(x, impl=1)
    assert torch.isfinite(j1).all() and torch.isfinite(y1).all()
    assert float((y1 - y_ref).norm() / y_ref.norm()) <= 3e-3
    row_err = (j1 - j_ref).flatten(1).norm(dim=1) / j_ref.flatten(1).norm(dim=1).clamp_min(1e-20)
    assert float((row_err <= 5e-3).float().mean()) >= 0.95, row_err.sort().values[-5:]


def test_jacobian_batch_sizes_and_tiles():
    """ragged tile counts (points not a multiple of the 128-row tiles; more tiles than SMs) give the same rows"""
    from curiousrl_b200.algorithm.ilqr_solver import nn_ilqr
    dev = torch.device("cuda", 0)
    net = _net(nn_ilqr.LargeNetwork, 11, 8, seed=3, dev=dev)
    f = nn_ilqr.FoldedMLP(net, dev)
    x = torch.randn(5000, 11, device=dev)
    _, full = f.jacobian(x, impl=1)
    for S in (1, 15, 16, 17, 127, 129, 2500):
        _, part = f.jacobian(x[:S].contiguous(), impl=1)
        assert torch.equal(part, full[:S]), S


def test_time_filter_is_scipys():
    from scipy.ndimage import gaussian_filter1d
    from curiousrl_b200.algorithm.ilqr_solver.nn_ilqr import gaussian_filter_matrix, smooth_time
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    F = rng.normal(size=(3, 100, 8, 11))
    G = torch.from_numpy(gaussian_filter_matrix(100, 5)).to(dev)
    out = smooth_time(torch.from_numpy(F).to(dev), G).cpu().numpy()
    ref = np.stack([gaussian_filter1d(F[b], sigma=5, axis=0) for b in range(3)])
    assert np.abs(out - ref).max() <= 1e-13


def _solver(scenario, **kw):
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import NNiLQR
    logger.set_level(45)
    logger.set_folder_name("test_nn_ilqr", remove_existing_folder=True).set_is_use_logger(True)
    return NNiLQR(seed=0, **kw).init(scenario)


def test_one_iteration_against_the_cpu_restatement():
    """F (network Jacobians, filtered), K, k and the accepted trajectory of one loop body (nn_ilqr.py:376-384) for an
    untrained-but-fixed network: device (FP32 CUDA-core Jacobian, so the masks are float32's) against oracle."""
    from curiousrl_b200.algorithm.ilqr_solver import SmallNetwork
    from curiousrl_b200.scenario.dynamic_model import TwoLinkPlanarManipulator
    from oracle import ilqr_numpy as orc, nn_ilqr_numpy as nn_orc
    scenario = TwoLinkPlanarManipulator(T=40)
    algo = _solver(scenario, network_class=SmallNetwork, trial_no=10, training_stopping_criterion=1e9, jacobian_impl=0)
    dev = algo._dev
    prob = orc.Problem(scenario)
    X = orc.rollout(prob)
    J0 = orc.cost(prob, X)
    import copy
    ref = nn_orc.iteration(prob, copy.deepcopy(algo.nn_dynamic_model._model), X, J0)
    Xd = dev.real(X.reshape(1, dev.T, dev.d))
    F = algo.learned_F(Xd)
    assert np.abs(F.cpu().numpy()[0] - ref["F"]).max() <= 1e-5 * np.abs(ref["F"]).max()
    _, c, C = dev.derivs(Xd, algo._params_for(1), want=("c", "C"))
    K, k = dev.backward_dense(F, c, C)
    scale = np.abs(ref["K"]).max()
    assert np.abs(K.cpu().numpy()[0] - ref["K"]).max() <= 1e-3 * scale      # float32 F in, float64 recursion out
    algo.set_obj_fun_value(J0)
    traj, _C, _c, _F, J, stop = algo.forward_pass(X[:, :, None], K.cpu().numpy()[0], k.cpu().numpy()[0][:, :, None])
    ref_ls = orc.line_search(prob, X, K.cpu().numpy()[0], k.cpu().numpy()[0], J0, None, 10, 0.5)
    assert abs(J - ref_ls[1]) <= 1e-9 * abs(ref_ls[1]) and J <= J0
    assert np.abs(traj[:, :, 0] - ref_ls[0]).max() <= 1e-8


def test_solve_and_solve_batch_reduce_the_cost():
    from curiousrl_b200.algorithm.ilqr_solver import SmallNetwork
    from curiousrl_b200.scenario.dynamic_model import TwoLinkPlanarManipulator
    from curiousrl_b200.utils.synthetic import synthetic_batch
    scenario = TwoLinkPlanarManipulator(T=40)
    algo = _solver(scenario, network_class=SmallNetwork, trial_no=20, training_stopping_criterion=5e-2, iLQR_max_iter=6,
                   pretrain_max_epoch=20000)
    J0 = algo.obj_fun.eval_obj_fun(algo.dynamic_model.eval_traj())
    algo.solve()
    assert algo.get_obj_fun_value() < J0
    assert algo._trajectory.shape == (40, 8, 1)
    x0, tgt = synthetic_batch(scenario.name, 64, seed=0)
    X, J, iters, stopped = algo.solve_batch(x0, tgt, max_iter=5)
    Xi, Ji = algo._dev.rollout(x0, algo._dev.real(tgt), None)
    # the model was fitted around the scenario's own initial state; on random initial states its Jacobians are rough, but
    # the line search runs on the real system: no cost may rise, and a good part of the problems must improve
    assert X.shape == (64, 40, 8) and bool((J <= Ji).all()) and float((J < Ji).float().mean()) > 0.25
    assert int(iters.max()) <= 5
