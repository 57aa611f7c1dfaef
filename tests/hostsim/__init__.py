"""ctypes front end of the TEST-ONLY host build of the device math (see hostsim.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_hostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "curiousrl_b200", "csrc", f)
                for f in ("ilqr_common.cuh", "ilqr_models.cuh", "ilqr_core.cuh")]

MODEL_ID = {"TwoLinkPlanarManipulator": 0, "ThreeLinkPlanarManipulator": 1, "RoboticArmTracking": 2}
DIMS = {0: (6, 2), 1: (8, 3), 2: (6, 2)}
U_BOX = {0: (-100.0, 100.0), 1: (-100.0, 100.0), 2: (-500.0, 500.0)}


def build(force=False):
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO]
        subprocess.run(cmd, check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class HostSim:
    """Single-trajectory calls into the host build; dtype 'f64' or 'f32'."""

    def __init__(self, model, dtype="f64"):
        self.mid = MODEL_ID[model] if isinstance(model, str) else int(model)
        self.n, self.m = DIMS[self.mid]
        self.d = self.n + self.m
        self.sfx = dtype
        self.np = np.float64 if dtype == "f64" else np.float32
        self.lo, self.hi = U_BOX[self.mid]

    def _fn(self, name):
        return getattr(lib(), "hs_%s_%s" % (name, self.sfx))

    def _params(self, target, T):
        p = np.ascontiguousarray(target, dtype=self.np)
        return (p, 0) if p.ndim == 1 else (p, p.shape[1])

    def rollout(self, x0, target, T, u0=None):
        x0 = np.ascontiguousarray(x0, dtype=self.np).reshape(-1)
        u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=self.np).reshape(T, self.m)
        p, st = self._params(target, T)
        X = np.zeros((T, self.d), dtype=self.np)
        J = ctypes.c_double()
        rc = self._fn("rollout")(self.mid, _p(x0), _p(u0), _p(p), st, ctypes.c_double(self.lo), ctypes.c_double(self.hi),
                                 T, _p(X), ctypes.byref(J))
        assert rc == 0
        return X, J.value

    def backward(self, X, target):
        X = np.ascontiguousarray(X, dtype=self.np)
        T = X.shape[0]
        p, st = self._params(target, T)
        K = np.zeros((T, self.m, self.n), dtype=self.np)
        k = np.zeros((T, self.m), dtype=self.np)
        assert self._fn("backward")(self.mid, _p(X), _p(p), st, T, _p(K), _p(k)) == 0
        return K, k

    def update(self, X, K, k, alpha, target):
        X = np.ascontiguousarray(X, dtype=self.np)
        K = np.ascontiguousarray(K, dtype=self.np)
        k = np.ascontiguousarray(k, dtype=self.np)
        T = X.shape[0]
        p, st = self._params(target, T)
        Xn = np.zeros_like(X)
        J = ctypes.c_double()
        rc = self._fn("update")(self.mid, _p(X), _p(K), _p(k), ctypes.c_double(alpha), _p(p), st,
                                ctypes.c_double(self.lo), ctypes.c_double(self.hi), T, _p(Xn), ctypes.byref(J))
        assert rc == 0
        return Xn, J.value

    def derivs(self, X, target):
        X = np.ascontiguousarray(X, dtype=self.np)
        T = X.shape[0]
        p, st = self._params(target, T)
        F = np.zeros((T, self.n, self.d), dtype=self.np)
        c = np.zeros((T, self.d), dtype=self.np)
        C = np.zeros((T, self.d, self.d), dtype=self.np)
        assert self._fn("derivs")(self.mid, _p(X), _p(p), st, T, _p(F), _p(c), _p(C)) == 0
        return F, c, C

    def backward_dense(self, F, c, C):
        F = np.ascontiguousarray(F, dtype=self.np)
        c = np.ascontiguousarray(c, dtype=self.np)
        C = np.ascontiguousarray(C, dtype=self.np)
        T = F.shape[0]
        n0, d0 = F.shape[1], F.shape[2]                   # any (n, m) the dense recursion is compiled for
        K = np.zeros((T, d0 - n0, n0), dtype=self.np)
        k = np.zeros((T, d0 - n0), dtype=self.np)
        n, m = C.shape[-1] - K.shape[1], K.shape[1]
        assert self._fn("backward_dense")(n, m, _p(F), _p(c), _p(C), T, _p(K), _p(k)) == 0
        return K, k

    def riccati_step(self, xu, p, V, v):
        """One structured Riccati step (the solve kernel's riccati_step) from the value function (V [n,n], v [n]) of the
        next time step; returns K [m,n], k [m] and the new (V, v)."""
        xu = np.ascontiguousarray(xu, dtype=self.np).reshape(-1)
        p = np.ascontiguousarray(p, dtype=self.np).reshape(-1)
        V = np.array(V, dtype=self.np).reshape(self.n, self.n).copy()
        v = np.array(v, dtype=self.np).reshape(self.n).copy()
        K = np.zeros((self.m, self.n), dtype=self.np)
        k = np.zeros(self.m, dtype=self.np)
        assert self._fn("riccati_step")(self.mid, _p(xu), _p(p), _p(V), _p(v), _p(K), _p(k)) == 0
        return K, k, V, v

    def solve(self, x0, target, T, max_iter=1000, is_check_stop=True, stopping_criterion=1e-6, max_line_search=10,
              gamma=0.5, line_search_method=0, stopping_method=0, J_init=np.inf, u0=None):
        x0 = np.ascontiguousarray(x0, dtype=self.np).reshape(-1)
        u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=self.np).reshape(T, self.m)
        p, st = self._params(target, T)
        X = np.zeros((T, self.d), dtype=self.np)
        J = ctypes.c_double()
        iters = ctypes.c_int()
        conv = ctypes.c_int()
        Jt = np.full(max_iter, np.nan)
        at = np.full(max_iter, -2, dtype=np.int32)
        rc = self._fn("solve")(self.mid, _p(x0), _p(u0), _p(p), st, ctypes.c_double(self.lo), ctypes.c_double(self.hi), T,
                               max_iter, int(is_check_stop), ctypes.c_double(stopping_criterion), max_line_search,
                               ctypes.c_double(gamma), line_search_method, stopping_method, ctypes.c_double(J_init),
                               _p(X), ctypes.byref(J), ctypes.byref(iters), ctypes.byref(conv), _p(Jt), _p(at))
        assert rc == 0
        n = iters.value
        return dict(X=X, J=J.value, iters=n, converged=bool(conv.value), Js=Jt[:n], alphas=at[:n])
