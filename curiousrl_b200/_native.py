"""ctypes binding of the C ABI declared in include/curiousrl_b200.h.

There is no CPU path behind these calls: if the library cannot be loaded, or no CUDA device
is present when a compute entry point is called, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

ABI_VERSION = 8
MODEL_IDS = {"TwoLinkPlanarManipulator": 0, "ThreeLinkPlanarManipulator": 1, "RoboticArmTracking": 2}
# the same scenarios under LogBarrieriLQR's cost (include/curiousrl_b200.h: CRL_*_BARRIER); FP64 only
BARRIER_MODEL_IDS = {name: mid + 3 for name, mid in MODEL_IDS.items()}
# scenarios with a GENERATED device model (CRL_GEN_*; curiousrl_b200/codegen.py): FP64, bounds compiled in
GENERIC_MODEL_IDS = {"CartPoleSwingUp1": 6, "CartPoleSwingUp2": 7, "VehicleTracking": 8, "CarParking": 9}
GENERIC_BARRIER_MODEL_IDS = {name: mid + 4 for name, mid in GENERIC_MODEL_IDS.items()}      # CRL_GEN_*_BARRIER
DTYPE_IDS = {"float64": 0, "float32": 1}
LINE_SEARCH_IDS = {"vanilla": 0, "feasibility": 1, "none": 2}
STOPPING_IDS = {"relative": 0, "vanilla": 1}

# every symbol include/curiousrl_b200.h declares (tests check the library exports all of them)
EXPORTS = ("crl_abi_version", "crl_status_string", "crl_model_dims", "crl_model_default_bounds", "crl_model_bounds", "crl_rollout",
           "crl_cost", "crl_derivs", "crl_backward", "crl_backward_dense", "crl_update_traj", "crl_forward_pass",
           "crl_solve_workspace_bytes", "crl_solve", "crl_solve_stages", "crl_mlp_jacobian_workspace_bytes", "crl_mlp_jacobian",
           "crl_smooth_time")


class Problem(Structure):
    _fields_ = [("model", c_int32), ("dtype", c_int32), ("batch", c_int64), ("horizon", c_int32),
                ("params_stride_t", c_int32), ("params_stride_b", c_int64), ("params", c_void_p),
                ("u_lo", c_double), ("u_hi", c_double), ("bounds_host", c_void_p)]


class SolverOpts(Structure):
    _fields_ = [("max_iter", c_int32), ("is_check_stop", c_int32), ("stopping_criterion", c_double),
                ("max_line_search", c_int32), ("gamma", c_double), ("line_search_method", c_int32),
                ("stopping_method", c_int32), ("grid_blocks", c_int32), ("reserved", c_int32),
                ("coop_threshold", c_int32), ("reserved2", c_int32)]


_lib = None


def _declare(lib):
    """argtypes / restype of every entry point the library has (a run-time compiled scenario library, jit.py, exports the
    iLQR part of the ABI only)."""
    P, O = POINTER(Problem), POINTER(SolverOpts)
    vp = c_void_p
    table = {
        "crl_abi_version": (c_int, []),
        "crl_status_string": (c_char_p, [c_int]),
        "crl_model_dims": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
        "crl_model_default_bounds": (c_int, [c_int, POINTER(c_double), POINTER(c_double)]),
        "crl_model_bounds": (c_int, [c_int, POINTER(c_double), POINTER(c_double)]),
        "crl_rollout": (c_int, [P, vp, vp, c_int64, vp, vp, vp]),
        "crl_cost": (c_int, [P, vp, vp, vp]),
        "crl_derivs": (c_int, [P, vp, vp, vp, vp, vp]),
        "crl_backward": (c_int, [P, vp, vp, vp, vp]),
        "crl_backward_dense": (c_int, [c_int32, c_int32, c_int32, c_int64, c_int32, vp, vp, vp, vp, vp, vp]),
        "crl_update_traj": (c_int, [P, vp, vp, vp, c_double, vp, vp, vp]),
        "crl_forward_pass": (c_int, [P, O, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "crl_solve_workspace_bytes": (c_size_t, [P, O]),
        "crl_solve": (c_int, [P, O, vp, vp, c_int64, vp, vp, vp, vp, vp, vp, vp, vp, c_size_t, vp]),
        "crl_solve_stages": (c_int, [P, O, c_int32, vp, vp, c_int64, vp, vp, vp, vp, vp, vp, vp, c_size_t, vp]),
        "crl_mlp_jacobian_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32, c_int64, c_int32]),
        "crl_mlp_jacobian": (c_int, [c_int32, c_int32, c_int32, c_int32, c_int64] + [vp] * 11 + [c_size_t, c_int32, vp]),
        "crl_smooth_time": (c_int, [c_int64, c_int32, c_int32, vp, vp, vp, vp]),
    }
    assert set(table) == set(EXPORTS)
    for name, (restype, argtypes) in table.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = restype
            fn.argtypes = argtypes


def load():
    """Load (building first if the in-tree library is missing or stale and nvcc is present)."""
    global _lib
    if _lib is None:
        path = os.environ.get("CURIOUSRL_B200_LIB")          # tuning experiments: load an alternative build
        if path:
            lib = ctypes.CDLL(path)
            _declare(lib)
            _lib = lib
            return _lib
        path = _build.LIB_PATH
        if _build.is_stale() and _build.find_nvcc() is not None:
            path = _build.build_native()
        if not os.path.exists(path):
            raise RuntimeError("curiousrl_b200: %s not found and cannot be built (no nvcc). "
                               "There is no CPU fallback for the iLQR path." % path)
        lib = ctypes.CDLL(path)
        _declare(lib)
        if lib.crl_abi_version() != ABI_VERSION:
            raise RuntimeError("curiousrl_b200: ABI version mismatch in %s" % path)
        _lib = lib
    return _lib


def check(status: int, what: str = "call"):
    if status != 0:
        msg = load().crl_status_string(status).decode()
        raise RuntimeError("curiousrl_b200 %s failed: %s (status %d)" % (what, msg, status))


def model_dims(model_id: int):
    n, m, p = c_int(), c_int(), c_int()
    check(load().crl_model_dims(model_id, n, m, p), "crl_model_dims")
    return n.value, m.value, p.value


def model_default_bounds(model_id: int):
    lo, hi = c_double(), c_double()
    check(load().crl_model_default_bounds(model_id, lo, hi), "crl_model_default_bounds")
    return lo.value, hi.value


def model_bounds(model_id: int):
    """The scenario's box `constr` as a [d, 2] list of (lo, hi) rows."""
    n, m, _ = model_dims(model_id)
    lo, hi = (c_double * (n + m))(), (c_double * (n + m))()
    check(load().crl_model_bounds(model_id, lo, hi), "crl_model_bounds")
    return [[lo[i], hi[i]] for i in range(n + m)]
