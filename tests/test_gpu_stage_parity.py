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


# ------------------------------------------------------------------------------------------------------------------
# Pivoted LU and teacher-forced Riccati steps through the C ABI.  Same inputs and bounds as the
# host-build tests in tests/test_riccati_teacher_forced.py; reference: ilqr_wrapper.py:240-255.

def dense_backward(n, m, F, c, C):
    """crl_backward_dense (the reference's backward_pass(C, c, F) signature) for any compiled (n, m); numpy in / out."""
    import ctypes
    from curiousrl_b200 import _native
    lib = _native.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    Ft = torch.as_tensor(np.ascontiguousarray(F, dtype=np.float64)).to(dev)
    B, T = Ft.shape[0], Ft.shape[1]
    ct = torch.as_tensor(np.ascontiguousarray(c, dtype=np.float64)).to(dev)
    Ct = torch.as_tensor(np.ascontiguousarray(C, dtype=np.float64)).to(dev)
    K = torch.empty((B, T, m, n), dtype=torch.float64, device=dev)
    k = torch.empty((B, T, m), dtype=torch.float64, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _native.check(lib.crl_backward_dense(n, m, 0, B, T, p(Ft), p(ct), p(Ct), p(K), p(k),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "crl_backward_dense")
    return K.cpu().numpy(), k.cpu().numpy()


@pytest.mark.parametrize("n,m", [(6, 2), (8, 3), (4, 1), (5, 1), (4, 2)])
def test_pivoted_lu_against_numpy(n, m):
    """crl_backward_dense with T = 1 is K = -solve(Quu, Qux), k = -solve(Quu, qu) for a caller-chosen Quu: non-symmetric,
    indefinite, negative definite, with row exchanges (one and two), zero / tied / tiny pivots, cond up to 1e10."""
    import riccati_cases as rc
    rng = np.random.default_rng(100 * n + m)
    names, Fs, cs, Cs, Ks, ks, conds = [], [], [], [], [], [], []
    for rep in range(20):
        for name, A in rc.lu_blocks(m, rng).items():
            F, c, C, Kref, kref = rc.lu_problem(n, m, A, rng)
            names.append(name); Fs.append(F); cs.append(c); Cs.append(C); Ks.append(Kref); ks.append(kref)
            conds.append(np.linalg.cond(A))
    K, k = dense_backward(n, m, np.stack(Fs), np.stack(cs), np.stack(Cs))
    for i, name in enumerate(names):
        tol = 4 * conds[i] * rc.EPS
        assert np.abs(K[i, 0] - Ks[i]).max() <= tol * np.abs(Ks[i]).max(), (name, conds[i])
        assert np.abs(k[i, 0] - ks[i]).max() <= tol * np.abs(ks[i]).max(), (name, conds[i])


@pytest.mark.parametrize("it", [0, 1, 5])
def test_dense_riccati_teacher_forced_robotic_arm_t200(golden, it):
    """RoboticArm T = 200 (BASELINE config 3): every step of the reference's backward pass from the reference's own F, c, C
    and value function - 28 / 9 / 15 of the 200 steps have a non-PD Quu, 29 / 43 / 55 exchange rows - within
    4 cond(Quu) eps; and windows of 4 and 8 consecutive steps wherever the recursion's own amplification keeps the bound
    below 1e-8."""
    import riccati_cases as rc
    g = golden("RoboticArm_T200")
    pre = "stage%d_" % it
    F, c, C, K, k = (g[pre + s] for s in ("F", "c", "C", "K", "k"))
    ref = rc.reference_value_functions(F, c, C, 6, 2)
    assert np.array_equal(ref["K"], K) and np.array_equal(ref["k"], k)
    assert (ref["lam_min"] < 0).sum() >= 9 and ref["swap"].sum() >= 29
    starts, Fw, cw, Cw = rc.teacher_forced_windows(F, c, C, ref["V"], ref["v"], 1)
    Kd, kd = dense_backward(6, 2, Fw, cw, Cw)
    for b, t in enumerate(starts):
        tol = 4 * ref["cond"][t] * rc.EPS
        assert np.abs(Kd[b, 0] - K[t]).max() <= tol * ref["K_scale"][t], (t, ref["lam_min"][t], ref["swap"][t])
        assert np.abs(kd[b, 0] - k[t]).max() <= tol * ref["k_scale"][t], t
    Vn = np.abs(ref["V"]).reshape(len(ref["V"]), -1).max(axis=1)
    for L in (4, 8):
        starts, Fw, cw, Cw = rc.teacher_forced_windows(F, c, C, ref["V"], ref["v"], L)
        Kd, kd = dense_backward(6, 2, Fw, cw, Cw)
        checked, checked_bad = 0, 0
        for b, t0 in enumerate(starts):
            win = slice(t0, t0 + L)
            growth = max(1.0, Vn[t0 + 1:t0 + L + 1].max() / max(Vn[t0 + 1:t0 + L + 1].min(), 1e-300))
            tol = 64.0 * L * ref["cond"][win].max() * min(growth, 1e30) * rc.EPS
            if tol > 1e-8:
                continue
            assert rc.rel_rows(Kd[b, :L], K[t0:t0 + L]).max() <= tol, (L, t0)
            checked += 1
            checked_bad += int((ref["lam_min"][win] < 0).any())
        assert checked >= 100 and checked_bad > 0


@pytest.mark.parametrize("tag", ["RoboticArm_T200", "RoboticArm_T100", "ThreeLink_T100"])
def test_fused_backward_equals_host_build(golden, tag):
    """The structured Riccati pass (crl_backward: F, c, C recomputed from X, what the solve kernel runs) on the GPU is the
    SAME arithmetic as the g++ build of the same headers, so the host build's teacher-forced parity at every step
    (tests/test_riccati_teacher_forced.py::test_structured_step_teacher_forced, non-PD steps included) carries over:
    gains bit-equal on the reference's stage trajectories, also where the recursion has blown up (|V| ~ 1e21)."""
    from hostsim import HostSim
    g = golden(tag)
    dev = device_model(g)
    hs = HostSim(str(g["model"]))
    tgt = g["default_target"]
    for it in g["stage_iters"]:
        X = g["stage%d_X" % it]
        K, k = dev.backward(X[None], tgt)
        Kh, kh = hs.backward(X, tgt)
        assert np.array_equal(K.cpu().numpy()[0], Kh) and np.array_equal(k.cpu().numpy()[0], kh)
