"""The C-ABI library: loads without a GPU, exports every symbol include/curiousrl_b200.h declares,
validates its arguments before touching CUDA, and the Python layer fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "curiousrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from curiousrl_b200 import _native
    lib = _native.load()
    names = header_functions()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), "symbol %s declared in the header is not exported" % name
    assert sorted(_native.EXPORTS) == names
    assert lib.crl_abi_version() == _native.ABI_VERSION == 8


def test_model_tables_and_status_strings():
    from curiousrl_b200 import _native
    assert _native.model_dims(0) == (6, 2, 2) and _native.model_dims(1) == (8, 3, 2) and _native.model_dims(2) == (6, 2, 2)
    assert _native.model_default_bounds(1) == (-100.0, 100.0) and _native.model_default_bounds(2) == (-500.0, 500.0)
    # LogBarrieriLQR's variants of the three models: same sizes, parameter row (target_x, target_y, t, t0)
    assert _native.model_dims(3) == (6, 2, 4) and _native.model_dims(4) == (8, 3, 4) and _native.model_dims(5) == (6, 2, 4)
    assert _native.model_default_bounds(5) == (-500.0, 500.0)
    lib = _native.load()
    assert lib.crl_model_dims(14, None, None, None) == -2
    assert _native.model_dims(12) == (4, 2, 2) and _native.model_dims(10) == (4, 1, 7)    # generated barrier rows: (columns, t, t0)
    assert b"unsupported model" in lib.crl_status_string(-2)
    assert lib.crl_status_string(0) == b"ok"
    with pytest.raises(RuntimeError, match="unsupported model"):
        _native.model_dims(99)
    # generated device models (CRL_GEN_*): sizes and the compiled-in box equal the scenario classes'
    import curiousrl_b200.scenario.dynamic_model as models
    for name, mid in _native.GENERIC_MODEL_IDS.items():
        sc = getattr(models, name)(T=5)
        p = 1 if sc.add_param is None else sc.add_param.shape[1]
        assert _native.model_dims(mid) == (sc.n, sc.m, p)
        assert np.array_equal(np.asarray(_native.model_bounds(mid)), np.asarray(sc.constr, dtype=np.float64))
    assert _native.model_bounds(1) == [[-np.inf, np.inf]] * 8 + [[-100.0, 100.0]] * 3


def test_argument_validation_happens_before_any_cuda_call():
    """Bad descriptors are rejected with status codes, not crashes - also on a box without a GPU."""
    from curiousrl_b200 import _native
    lib = _native.load()
    good = dict(model=1, dtype=0, batch=4, horizon=10, params_stride_t=0, params_stride_b=2, params=0x1000,
                u_lo=-100.0, u_hi=100.0)

    def prob(**kw):
        d = dict(good)
        d.update(kw)
        return _native.Problem(d["model"], d["dtype"], d["batch"], d["horizon"], d["params_stride_t"],
                               d["params_stride_b"], d["params"], d["u_lo"], d["u_hi"])

    opts = _native.SolverOpts(10, 1, 1e-6, 10, 0.5, 0, 0, 0, 0, 0, 0)
    dummy = ctypes.c_void_p(0x1000)
    assert lib.crl_rollout(None, dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(prob(model=14), dummy, None, 0, dummy, None, None) == -2
    assert lib.crl_rollout(prob(model=6, dtype=1), dummy, None, 0, dummy, None, None) == -2   # generated models are FP64 only
    assert lib.crl_rollout(prob(model=6, params_stride_t=2), dummy, None, 0, dummy, None, None) == -1   # its rows are 5 wide
    assert lib.crl_rollout(prob(model=5, dtype=1), dummy, None, 0, dummy, None, None) == -2   # barrier models are FP64 only
    assert lib.crl_rollout(prob(dtype=3), dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(prob(horizon=0), dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(prob(batch=-1), dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(prob(params_stride_t=3), dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(prob(u_lo=1.0, u_hi=-1.0), dummy, None, 0, dummy, None, None) == -1
    # generated models: a caller-supplied box (host array [d][2]) is validated; the barrier variants take the default only
    box = np.array([[-np.inf, np.inf]] * 4 + [[-2.0, 2.0]], dtype=np.float64)
    bad = np.array([[-np.inf, np.inf]] * 4 + [[2.0, -2.0]], dtype=np.float64)
    gen = dict(model=6, params_stride_t=5, params_stride_b=0)
    g_ok = _native.Problem(6, 0, 4, 10, 5, 0, 0x1000, 0.0, 0.0, box.ctypes.data)
    g_bad = _native.Problem(6, 0, 4, 10, 5, 0, 0x1000, 0.0, 0.0, bad.ctypes.data)
    g_bar = _native.Problem(10, 0, 4, 10, 7, 0, 0x1000, 0.0, 0.0, box.ctypes.data)
    assert lib.crl_solve_workspace_bytes(ctypes.byref(g_ok), ctypes.byref(opts)) > 0
    assert lib.crl_rollout(g_bad, dummy, None, 0, dummy, None, None) == -1
    assert lib.crl_rollout(g_bar, dummy, None, 0, dummy, None, None) == -2
    assert lib.crl_rollout(prob(batch=0), None, None, 0, None, None, None) == 0          # empty batch: nothing to do
    assert lib.crl_rollout(prob(), None, None, 0, dummy, None, None) == -1               # missing x0
    bad_opts = _native.SolverOpts(10, 1, 1e-6, 10, 0.5, 7, 0, 0, 0, 0, 0)
    assert lib.crl_forward_pass(prob(), ctypes.byref(bad_opts), dummy, dummy, dummy, dummy, dummy, dummy, dummy, dummy,
                                None) == -1
    assert lib.crl_backward_dense(5, 2, 0, 1, 10, dummy, dummy, dummy, dummy, dummy, None) == -2
    assert lib.crl_backward_dense(4, 1, 1, 1, 10, dummy, dummy, dummy, dummy, dummy, None) == -2   # FP64 only
    assert lib.crl_solve_workspace_bytes(ctypes.byref(prob(model=14)), ctypes.byref(opts)) == 0
    assert lib.crl_solve_workspace_bytes(ctypes.byref(prob(model=8, batch=33, horizon=10)), ctypes.byref(opts)) == 2 * 10 * 32 * (12 + 8 + 2) * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, DeviceModel
    from conftest import make_scenario
    logger.set_level(45)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BasiciLQR(max_line_search=10).init(make_scenario("TwoLinkPlanarManipulator", 100))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceModel("ThreeLinkPlanarManipulator", 100, -100.0, 100.0)
    # a host pointer (or no device at all) never reaches a kernel
    from curiousrl_b200 import _native
    lib = _native.load()
    x = np.zeros(8)
    prob = _native.Problem(1, 0, 1, 10, 0, 0, x.ctypes.data, -100.0, 100.0)
    rc = lib.crl_rollout(prob, x.ctypes.data_as(ctypes.c_void_p), None, 0, x.ctypes.data_as(ctypes.c_void_p), None, None)
    assert rc < 0


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under curiousrl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "curiousrl_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hostsim" not in text or f == "ilqr_common.cuh", f


def test_learned_dynamics_entry_points_validate_before_any_cuda_call():
    """crl_mlp_jacobian / crl_smooth_time (NNiLQR): sizes outside what the kernels are written for, missing
    buffers, a workspace that is too small or misaligned and host pointers are all refused with a status - no launch."""
    from curiousrl_b200 import _native
    lib = _native.load()
    vp = ctypes.c_void_p
    raw = np.zeros(4096 + 64, dtype=np.float32)
    off = (-raw.ctypes.data % 256) // 4                                   # a 256-byte aligned view
    buf = raw[off:off + 4096]
    p = buf.ctypes.data_as(vp)
    need = lib.crl_mlp_jacobian_workspace_bytes(11, 800, 416, 8, 100, 1)
    assert need > 0 and lib.crl_mlp_jacobian_workspace_bytes(11, 800, 416, 8, 100, 0) > need        # impl 0 keeps its activations
    assert lib.crl_mlp_jacobian_workspace_bytes(11, 0, 416, 8, 100, 1) == 0

    def call(d=11, h1=800, h2=416, n=8, S=100, ws=p, ws_bytes=1 << 30, impl=1, x=p):
        return lib.crl_mlp_jacobian(d, h1, h2, n, S, x, p, p, p, p, p, p, p, p, p, ws, ws_bytes, impl, None)

    assert call(S=0) == 0                                                 # nothing to do
    for bad in (dict(d=17), dict(d=0), dict(n=17), dict(h1=801), dict(h2=400), dict(h2=544), dict(h1=16), dict(S=-1), dict(impl=2)):
        assert call(**bad) == -1, bad                                     # CRL_ERR_BAD_ARGUMENT
    assert call(x=None) == -1 and call(ws=None) == -1
    assert call(ws_bytes=16) == -3                                        # CRL_ERR_WORKSPACE_TOO_SMALL
    assert call(ws=vp(buf.ctypes.data + 4)) == -1                         # workspace must be 256-byte aligned
    assert call() < 0                                                     # host pointers never reach a kernel
    g = np.zeros((4, 4))
    gp = g.ctypes.data_as(vp)
    assert lib.crl_smooth_time(0, 4, 3, gp, gp, gp, None) == 0
    assert lib.crl_smooth_time(1, 0, 3, gp, gp, gp, None) == -1 and lib.crl_smooth_time(1, 4, 3, gp, gp, gp, None) == -1   # in == out
    out = np.zeros((4, 3))
    assert lib.crl_smooth_time(1, 4, 3, gp, g.ctypes.data_as(vp), out.ctypes.data_as(vp), None) < 0          # host pointers
