"""Stage ("teacher-forced") parity of the CUDA path, called through the C ABI, against the
reference's golden stage tensors and the CPU oracle.

Tolerances (FP64): Two/ThreeLink <= 1e-11 relative per stage (the reference's BLAS summation
order and libm sin/cos are not bit-reproducible on a GPU); RoboticArm gains are compared on the
well-conditioned early iterations only, with cond * eps head-room, as the recursion amplifies
rounding noise by up to 1e6 there.
"""
import numpy as np
import pytest
import torch

from conftest import make_problem, make_scenario
from oracle import ilqr_numpy as orc

pytestmark = pytest.mark.gpu

TAGS = ["TwoLink_T100", "ThreeLink_T100", "RoboticArm_T100", "RoboticArm_T200", "ThreeLink_T30"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def device_model(g, dtype="float64"):
    from curiousrl_b200.algorithm.ilqr_solver import DeviceModel
    from curiousrl_b200 import _native
    model = str(g["model"])
    lo, hi = _native.model_default_bounds(_native.MODEL_IDS[model])
    return DeviceModel(model, int(g["T"]), lo, hi, dtype=dtype)


@pytest.mark.parametrize("tag", TAGS)
def test_rollout_cost_derivs(golden, tag):
    g = golden(tag)
    dev = device_model(g)
    tgt = g["default_target"]
    X, J = dev.rollout(g["default_x0"][None], tgt)
    Xh = X.cpu().numpy()[0]
    assert rel(Xh, g["default_X_init"]) <= 1e-14
    prob = make_problem(str(g["model"]), int(g["T"]))
    assert abs(J.item() - orc.cost(prob, g["default_X_init"], tgt)) <= 1e-13 * abs(J.item())
    for it in g["stage_iters"]:
        pre = "stage%d_" % it
        F, c, C = dev.derivs(g[pre + "X"][None], tgt)
        assert rel(F.cpu().numpy()[0], g[pre + "F"]) <= 1e-13
        assert rel(c.cpu().numpy()[0], g[pre + "c"]) <= 1e-14
        assert np.array_equal(C.cpu().numpy()[0], g[pre + "C"])
        Jc = dev.cost(g[pre + "cand_X"], tgt).cpu().numpy()
        assert np.allclose(Jc, g[pre + "cand_J"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("tag", TAGS)
def test_backward_and_update(golden, tag):
    g = golden(tag)
    dev = device_model(g)
    tgt = g["default_target"]
    arm = str(g["model"]) == "RoboticArmTracking"
    for it in g["stage_iters"]:
        pre = "stage%d_" % it
        X = g[pre + "X"]
        K, k = dev.backward(X[None], tgt)
        Kd, kd = dev.backward_dense(g[pre + "F"][None], g[pre + "c"][None], g[pre + "C"][None])
        if not arm:
            tol = 1e-10
            assert rel(K.cpu().numpy()[0], g[pre + "K"]) <= tol and rel(k.cpu().numpy()[0], g[pre + "k"]) <= tol
            assert rel(Kd.cpu().numpy()[0], g[pre + "K"]) <= tol and rel(kd.cpu().numpy()[0], g[pre + "k"]) <= tol
        elif int(g["T"]) == 100 and it == 0:
            assert rel(K.cpu().numpy()[0], g[pre + "K"]) <= 1e-6 and rel(kd.cpu().numpy()[0], g[pre + "k"]) <= 1e-9
        # structural known answers
        assert np.all(K.cpu().numpy()[0][-1] == 0)
        # closed-loop rollouts from the reference's own gains: well conditioned for every model
        for alpha, Xc, Jc in zip(g[pre + "cand_alpha"], g[pre + "cand_X"], g[pre + "cand_J"]):
            Xn, Jn = dev.update_traj(X[None], g[pre + "K"][None], g[pre + "k"][None], float(alpha), tgt)
            assert rel(Xn.cpu().numpy()[0], Xc) <= 1e-12
            assert abs(Jn.item() - Jc) <= 1e-12 * abs(Jc)
            assert np.all(Xn.cpu().numpy()[0][-1, dev.n:] == 0)


@pytest.mark.parametrize("tag", ["TwoLink_T100", "ThreeLink_T100", "RoboticArm_T100"])
def test_forward_pass_decisions(golden, tag):
    """Accepted step index, returned trajectory and stop flag given the reference's X, K, k, J_last."""
    from curiousrl_b200 import _native
    g = golden(tag)
    dev = device_model(g)
    tgt = g["default_target"]
    opts = _native.SolverOpts(1, 1, 1e-6, int(g["max_line_search"]), 0.5, 0, 0, 0, 0, 0, 0)
    for it in g["stage_iters"]:
        pre = "stage%d_" % it
        Xn, Jn, aidx, stop = dev.forward_pass(opts, g[pre + "X"][None], g[pre + "K"][None], g[pre + "k"][None],
                                              np.asarray([float(g[pre + "J_last"])]), tgt)
        assert int(aidx.item()) == int(g["default_alphas"][it])
        assert rel(Xn.cpu().numpy()[0], g[pre + "X_next"]) <= 1e-12
        assert abs(Jn.item() - float(g[pre + "J_next"])) <= 1e-12 * abs(Jn.item())
        assert bool(stop.item()) == bool(g[pre + "stop"])


def test_stage_kernels_batched_equal_single(golden):
    """A batch through the stage kernels equals the same problems one at a time (bit for bit)."""
    g = golden("ThreeLink_T100")
    dev = device_model(g)
    x0, tgt = g["batch_x0"], g["batch_target"]
    X, J = dev.rollout(x0, tgt)
    K, k = dev.backward(X, tgt)
    for b in (0, 5, 15):
        Xb, Jb = dev.rollout(x0[b:b + 1], tgt[b:b + 1])
        Kb, kb = dev.backward(Xb, tgt[b:b + 1])
        assert torch.equal(Xb[0], X[b]) and torch.equal(Jb[0], J[b])
        assert torch.equal(Kb[0], K[b]) and torch.equal(kb[0], k[b])


def test_float32_stage_accuracy(golden):
    g = golden("ThreeLink_T100")
    dev = device_model(g, "float32")
    tgt = g["default_target"]
    X, J = dev.rollout(g["default_x0"][None], tgt)
    assert rel(X.cpu().numpy()[0], g["default_X_init"]) <= 1e-6
    K, k = dev.backward(g["stage1_X"][None], tgt)
    assert rel(K.cpu().numpy()[0], g["stage1_K"]) <= 1e-3
