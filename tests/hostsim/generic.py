"""ctypes front end of the TEST-ONLY host build of the generic (generated-model) device math."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_hostsim_generic.so")
SRC = os.path.join(HERE, "hostsim_generic.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "curiousrl_b200", "csrc", f)
                for f in ("ilqr_common.cuh", "ilqr_models.cuh", "ilqr_core.cuh", "ilqr_generic.cuh", "ilqr_generated.cuh")]
MODEL_ID = {"CartPoleSwingUp1": 6, "CartPoleSwingUp2": 7, "VehicleTracking": 8, "CarParking": 9}
BARRIER_ID = {name: mid + 4 for name, mid in MODEL_ID.items()}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO],
                           check=True)
        _lib = ctypes.CDLL(SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GenericHostSim:
    def __init__(self, model, barrier=False):
        self.mid = (BARRIER_ID if barrier else MODEL_ID)[model]
        n, m, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib().hg_dims(self.mid, ctypes.byref(n), ctypes.byref(m), ctypes.byref(p)) == 0
        self.n, self.m, self.p = n.value, m.value, p.value
        self.d = self.n + self.m

    def set_box(self, constr=None):
        """Box of the following calls ([d, 2] rows of (lo, hi)); None = the model's default."""
        if constr is None:
            lib().hg_set_box(None, 0)
        else:
            rows = _f64(constr).reshape(self.d, 2)
            lib().hg_set_box(_p(rows), self.d)

    def bounds(self):
        lo, hi = np.zeros(self.d), np.zeros(self.d)
        assert lib().hg_bounds(self.mid, _p(lo), _p(hi)) == 0
        return np.stack([lo, hi], axis=1)

    def _params(self, P, T):
        P = _f64(np.zeros((T, 1)) if P is None else P)
        return (P, 0) if P.ndim == 1 else (P, P.shape[1])

    def rollout(self, x0, P, T, u0=None):
        x0 = _f64(x0).reshape(-1)
        u0 = None if u0 is None else _f64(u0).reshape(T, self.m)
        P, st = self._params(P, T)
        X, J = np.zeros((T, self.d)), ctypes.c_double()
        assert lib().hg_rollout(self.mid, _p(x0), _p(u0), _p(P), st, T, _p(X), ctypes.byref(J)) == 0
        return X, J.value

    def derivs(self, X, P):
        X = _f64(X)
        T = X.shape[0]
        P, st = self._params(P, T)
        F, c, C = np.zeros((T, self.n, self.d)), np.zeros((T, self.d)), np.zeros((T, self.d, self.d))
        assert lib().hg_derivs(self.mid, _p(X), _p(P), st, T, _p(F), _p(c), _p(C)) == 0
        return F, c, C

    def backward(self, X, P):
        X = _f64(X)
        T = X.shape[0]
        P, st = self._params(P, T)
        K, k = np.zeros((T, self.m, self.n)), np.zeros((T, self.m))
        assert lib().hg_backward(self.mid, _p(X), _p(P), st, T, _p(K), _p(k)) == 0
        return K, k

    def update(self, X, K, k, alpha, P):
        X, K, k = _f64(X), _f64(K), _f64(k)
        T = X.shape[0]
        P, st = self._params(P, T)
        Xn, J = np.zeros_like(X), ctypes.c_double()
        assert lib().hg_update(self.mid, _p(X), _p(K), _p(k), ctypes.c_double(alpha), _p(P), st, T, _p(Xn),
                               ctypes.byref(J)) == 0
        return Xn, J.value

    def solve(self, x0, P, T, max_iter=1000, is_check_stop=True, stopping_criterion=1e-6, max_line_search=10, gamma=0.5,
              line_search_method=0, stopping_method=0, J_init=np.inf, u0=None, stages=None):
        """stages = None: one solve with parameter rows P.  stages = [S, T, p] (or [S, p]): LogBarrieriLQR's outer loop,
        one inner loop per parameter set; the result carries `inner_iters`."""
        x0 = _f64(x0).reshape(-1)
        u0 = None if u0 is None else _f64(u0).reshape(T, self.m)
        n_stage, stage_stride, sit = 1, 0, None
        if stages is not None:
            stages = _f64(stages)
            n_stage = stages.shape[0]
            P, st = (stages, stages.shape[2]) if stages.ndim == 3 else (stages, 0)
            stage_stride = T * stages.shape[2] if stages.ndim == 3 else stages.shape[1]
            sit = np.zeros(n_stage, dtype=np.int32)
        else:
            P, st = self._params(P, T)
        X, J, iters, conv = np.zeros((T, self.d)), ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
        trace = stages is None
        Jt, at = np.full(max_iter, np.nan), np.full(max_iter, -2, dtype=np.int8)
        assert lib().hg_solve(self.mid, _p(x0), _p(u0), _p(P), st, T, max_iter, int(is_check_stop),
                              ctypes.c_double(stopping_criterion), max_line_search, ctypes.c_double(gamma),
                              line_search_method, stopping_method, ctypes.c_double(J_init), _p(X), ctypes.byref(J),
                              ctypes.byref(iters), ctypes.byref(conv), _p(Jt) if trace else None, _p(at) if trace else None,
                              n_stage, ctypes.c_longlong(stage_stride), _p(sit)) == 0
        n = iters.value
        out = dict(X=X, J=J.value, iters=n, converged=bool(conv.value))
        if trace:
            out.update(Js=Jt[:n], alphas=at[:n].astype(np.int32))
        else:
            out["inner_iters"] = sit
        return out
