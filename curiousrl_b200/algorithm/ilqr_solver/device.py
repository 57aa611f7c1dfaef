"""Device session of one scenario: model id, dtype, horizon, bounds + typed calls into the C ABI.

Everything numeric on the iLQR path goes through `DeviceModel`, which only moves pointers of
torch CUDA tensors into libcuriousrl_b200.so.  PyTorch is used for device memory and streams,
nothing else.  There is no CPU path: constructing a `DeviceModel` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from ... import _native

_TORCH_DTYPE = {"float64": torch.float64, "float32": torch.float32}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@dataclass
class BatchResult:
    """Result of a batched solve (device tensors unless `to_host()` is called)."""
    trajectory: Optional[torch.Tensor]      # [B, T, n+m] final trajectories (None if not requested)
    obj_fun_value: torch.Tensor             # [B] float64 final cost
    iterations: torch.Tensor                # [B] int32 number of backward+forward pairs
    converged: torch.Tensor                 # [B] uint8: 1 = stopping criterion met, 0 = ran max_iter
    obj_fun_trace: Optional[torch.Tensor] = None   # [B, max_iter] float64 (NaN-padded) if trace=True
    alpha_trace: Optional[torch.Tensor] = None     # [B, max_iter] int8, accepted step index, -1 = rejected, -2 = unused

    def to_host(self):
        conv = lambda t: None if t is None else t.cpu().numpy()
        return BatchResult(*(conv(getattr(self, f)) for f in
                             ("trajectory", "obj_fun_value", "iterations", "converged", "obj_fun_trace", "alpha_trace")))


class DeviceModel:
    def __init__(self, model_name: str, T: int, u_lo: float, u_hi: float, dtype: str = "float64", device=None,
                 barrier: bool = False, bounds=None, scenario=None):
        """`scenario`: the scenario object - needed only for a scenario without a precompiled device model, which is
        then generated from its sympy expressions and compiled at run time (curiousrl_b200/jit.py)."""
        self.generic = model_name in _native.GENERIC_MODEL_IDS      # generated device model: FP64, bounds compiled in
        self.jit = model_name not in _native.MODEL_IDS and not self.generic
        if self.jit and scenario is None:
            raise Exception('Scenario "%s" has no precompiled device model in curiousrl_b200 (%s) and no scenario object was '
                            "given to compile one from; there is no CPU fallback." % (
                                model_name, ", ".join(list(_native.MODEL_IDS) + list(_native.GENERIC_MODEL_IDS))))
        self.generic = self.generic or self.jit
        if self.generic and dtype != "float64":
            raise ValueError('the generated device model of "%s" is compiled for float64 only' % model_name)
        if dtype not in _TORCH_DTYPE:
            raise ValueError("dtype must be 'float64' or 'float32'")
        if dtype == "float32" and model_name == "RoboticArmTracking" and int(T) > 100:
            # the FP64 reference does not reproduce ITSELF on this scenario beyond T = 100 (1e-15 noise moves its cost by
            # O(1)); north_star's FP32 bar (1e-4 on the final cost) cannot be stated there, so the mode is not offered
            raise ValueError('dtype="float32" is not offered for RoboticArmTracking with T > 100: the scenario is chaotic '
                             "for the float64 reference itself there; use float64")
        if self.jit:
            from ... import jit as _jit
            self.lib, self.lib_path = _jit.compile_scenario(scenario)
        else:
            self.lib = _native.load()
        if not torch.cuda.is_available():
            raise RuntimeError("curiousrl_b200 needs a CUDA device (B200); none is visible and there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if barrier and dtype != "float64":
            raise ValueError("the log-barrier cost is compiled for float64 only")
        self.model_name = model_name
        if self.jit:
            self.model_id = 101 if barrier else 100                    # CRL_GEN_USER_BARRIER / CRL_GEN_USER
        else:
            self.model_id = ((_native.GENERIC_BARRIER_MODEL_IDS if barrier else _native.GENERIC_MODEL_IDS) if self.generic else
                             _native.BARRIER_MODEL_IDS if barrier else _native.MODEL_IDS)[model_name]
        n_, m_, p_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _native.check(self.lib.crl_model_dims(self.model_id, n_, m_, p_), "crl_model_dims")
        self.n, self.m, self.p = n_.value, m_.value, p_.value
        self.d = self.n + self.m
        self.T = int(T)
        self.dtype_name, self.dtype, self.dtype_id = dtype, _TORCH_DTYPE[dtype], _native.DTYPE_IDS[dtype]
        self.u_lo, self.u_hi = float(u_lo), float(u_hi)
        # generated models: the box `constr` [d, 2] if it is not the scenario's default (host array, read by every call)
        self._bounds = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.float64).reshape(self.d, 2)
        if self._bounds is not None and not self.generic:
            raise ValueError("per-component bounds are for the generated models")
        self._workspace = None
        self._workspaces = {}      # stream handle -> workspace tensor
        self.launches = 0          # kernels enqueued through this object (bench accounting)

    # ------------------------------------------------------------------ tensor plumbing
    def real(self, x, shape=None) -> torch.Tensor:
        t = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
        t = t.to(device=self.device, dtype=self.dtype).contiguous()
        return t if shape is None else t.reshape(shape)

    def empty(self, *shape, dtype=None):
        return torch.empty(shape, device=self.device, dtype=self.dtype if dtype is None else dtype)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def problem(self, params: torch.Tensor, B: int) -> _native.Problem:
        """params: [p] shared, [B, p] per-trajectory constant, [T, p] shared time-varying (B must be given
        explicitly via shape [1, T, p]) or [B, T, p]."""
        if params.dim() == 1:
            sb, st = 0, 0
        elif params.dim() == 2:
            if params.shape[0] != B:
                raise ValueError("add_param of shape %s does not match batch %d" % (tuple(params.shape), B))
            sb, st = self.p, 0
        elif params.dim() == 3:
            if params.shape[1] != self.T or params.shape[0] not in (1, B):
                raise ValueError("add_param of shape %s does not match [B=%d, T=%d, p]" % (tuple(params.shape), B, self.T))
            sb, st = (0 if params.shape[0] == 1 and B != 1 else self.T * self.p), self.p
        else:
            raise ValueError("add_param must have 1, 2 or 3 dimensions")
        if params.shape[-1] != self.p:
            raise ValueError("add_param last dimension must be %d" % self.p)
        return _native.Problem(self.model_id, self.dtype_id, B, self.T, st, sb, params.data_ptr(), self.u_lo, self.u_hi,
                               self._bounds_ptr())

    def _bounds_ptr(self):
        return None if self._bounds is None else self._bounds.ctypes.data

    def _u0(self, u0, B):
        """None -> zeros; [T, m] shared; [B, T, m] per trajectory.  Returns (tensor-or-None, batch stride)."""
        if u0 is None:
            return None, 0
        u0 = self.real(u0)
        if u0.dim() >= 2 and u0.shape[-1] == 1 and u0.shape[-2] == self.m:   # reference layout [T, m, 1]
            u0 = u0.squeeze(-1)
        if u0.dim() == 2 and tuple(u0.shape) == (self.T, self.m):
            return u0.contiguous(), 0
        if u0.dim() == 3 and tuple(u0.shape) == (B, self.T, self.m):
            return u0.contiguous(), self.T * self.m
        raise ValueError("init_action must be [T, m], [T, m, 1] or [B, T, m]")

    # ------------------------------------------------------------------ stage calls (API layout)
    def rollout(self, x0, params, u0=None):
        x0 = self.real(x0).reshape(-1, self.n)
        B = x0.shape[0]
        params = self.real(params)
        u0t, u0s = self._u0(u0, B)
        X = self.empty(B, self.T, self.d)
        J = self.empty(B, dtype=torch.float64)
        prob = self.problem(params, B)
        _native.check(self.lib.crl_rollout(prob, _ptr(x0), _ptr(u0t), u0s, _ptr(X), _ptr(J), self._stream()), "crl_rollout")
        self.launches += 1
        return X, J

    def cost(self, X, params):
        X = self.real(X).reshape(-1, self.T, self.d)
        B = X.shape[0]
        params = self.real(params)
        J = self.empty(B, dtype=torch.float64)
        _native.check(self.lib.crl_cost(self.problem(params, B), _ptr(X), _ptr(J), self._stream()), "crl_cost")
        self.launches += 1
        return J

    def derivs(self, X, params, want=("F", "c", "C")):
        X = self.real(X).reshape(-1, self.T, self.d)
        B = X.shape[0]
        params = self.real(params)
        F = self.empty(B, self.T, self.n, self.d) if "F" in want else None
        c = self.empty(B, self.T, self.d) if "c" in want else None
        C = self.empty(B, self.T, self.d, self.d) if "C" in want else None
        _native.check(self.lib.crl_derivs(self.problem(params, B), _ptr(X), _ptr(F), _ptr(c), _ptr(C), self._stream()),
                      "crl_derivs")
        self.launches += 1
        return F, c, C

    def backward(self, X, params):
        X = self.real(X).reshape(-1, self.T, self.d)
        B = X.shape[0]
        params = self.real(params)
        K = self.empty(B, self.T, self.m, self.n)
        k = self.empty(B, self.T, self.m)
        _native.check(self.lib.crl_backward(self.problem(params, B), _ptr(X), _ptr(K), _ptr(k), self._stream()), "crl_backward")
        self.launches += 1
        return K, k

    def backward_dense(self, F, c, C):
        F = self.real(F).reshape(-1, self.T, self.n, self.d)
        B = F.shape[0]
        c = self.real(c).reshape(B, self.T, self.d)
        C = self.real(C).reshape(B, self.T, self.d, self.d)
        K = self.empty(B, self.T, self.m, self.n)
        k = self.empty(B, self.T, self.m)
        _native.check(self.lib.crl_backward_dense(self.n, self.m, self.dtype_id, B, self.T, _ptr(F), _ptr(c), _ptr(C),
                                                  _ptr(K), _ptr(k), self._stream()), "crl_backward_dense")
        self.launches += 1
        return K, k

    def update_traj(self, X, K, k, alpha, params):
        X = self.real(X).reshape(-1, self.T, self.d)
        B = X.shape[0]
        K = self.real(K).reshape(B, self.T, self.m, self.n)
        k = self.real(k).reshape(B, self.T, self.m)
        params = self.real(params)
        Xn = self.empty(B, self.T, self.d)
        Jn = self.empty(B, dtype=torch.float64)
        _native.check(self.lib.crl_update_traj(self.problem(params, B), _ptr(X), _ptr(K), _ptr(k), float(alpha), _ptr(Xn),
                                               _ptr(Jn), self._stream()), "crl_update_traj")
        self.launches += 1
        return Xn, Jn

    def forward_pass(self, opts: _native.SolverOpts, X, K, k, J_last, params):
        X = self.real(X).reshape(-1, self.T, self.d)
        B = X.shape[0]
        K = self.real(K).reshape(B, self.T, self.m, self.n)
        k = self.real(k).reshape(B, self.T, self.m)
        params = self.real(params)
        J_last = torch.as_tensor(J_last, dtype=torch.float64).to(self.device).reshape(B).contiguous()
        Xn = self.empty(B, self.T, self.d)
        Jn = self.empty(B, dtype=torch.float64)
        aidx = self.empty(B, dtype=torch.int32)
        stop = self.empty(B, dtype=torch.uint8)
        _native.check(self.lib.crl_forward_pass(self.problem(params, B), ctypes.byref(opts), _ptr(X), _ptr(K), _ptr(k),
                                                _ptr(J_last), _ptr(Xn), _ptr(Jn), _ptr(aidx), _ptr(stop), self._stream()),
                      "crl_forward_pass")
        self.launches += 1
        return Xn, Jn, aidx, stop

    # ------------------------------------------------------------------ the product path
    def workspace(self, prob, opts) -> torch.Tensor:
        need = self.lib.crl_solve_workspace_bytes(prob, ctypes.byref(opts))
        if need == 0:
            raise RuntimeError("curiousrl_b200: cannot size the solver workspace (no CUDA device?)")
        # one workspace per stream: the control block and the tiles belong to the solve that is in flight, and two
        # solves enqueued on different streams from one solver object would otherwise share them
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._workspaces.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)   # caching allocator: 512 B aligned
            self._workspaces[key] = ws
        self._workspace = ws                                                  # the most recent one (phase timers, tools)
        return ws

    def _check_out(self, out: BatchResult, B: int, max_iter: int):
        """A caller-supplied result object is written by raw pointer: refuse anything the kernel would overrun."""
        def need(name, t, shape, dtype, host_ok=False):
            if t is None:
                return
            where_ok = (t.device == self.device) or (host_ok and t.device.type == "cpu" and t.is_pinned())
            if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or not where_ok:
                raise ValueError("out.%s must be a contiguous %s tensor of shape %s on %s%s (got %s %s on %s)" % (
                    name, dtype, tuple(shape), self.device, " or in pinned host memory" if host_ok else "",
                    t.dtype, tuple(t.shape), t.device))
        need("trajectory", out.trajectory, (B, self.T, self.d), self.dtype, host_ok=True)
        need("obj_fun_value", out.obj_fun_value, (B,), torch.float64)
        need("iterations", out.iterations, (B,), torch.int32)
        need("converged", out.converged, (B,), torch.uint8)
        need("obj_fun_trace", out.obj_fun_trace, (B, max_iter), torch.float64)
        need("alpha_trace", out.alpha_trace, (B, max_iter), torch.int8)
        if out.obj_fun_value is None or out.iterations is None or out.converged is None:
            raise ValueError("out.obj_fun_value, out.iterations and out.converged are required")

    def solve(self, opts: _native.SolverOpts, x0, params, u0=None, J_init=None, return_trajectory=True,
              trace=False, out: Optional[BatchResult] = None) -> BatchResult:
        x0 = self.real(x0).reshape(-1, self.n)
        B = x0.shape[0]
        params = self.real(params)
        u0t, u0s = self._u0(u0, B)
        prob = self.problem(params, B)
        ws = self.workspace(prob, opts)
        if J_init is not None:
            J_init = torch.as_tensor(J_init, dtype=torch.float64).to(self.device).reshape(-1)
            if J_init.numel() == 1:
                J_init = J_init.expand(B)
            J_init = J_init.contiguous()
        if out is not None:
            self._check_out(out, B, opts.max_iter)
        if out is None:
            out = BatchResult(self.empty(B, self.T, self.d) if return_trajectory else None,
                              self.empty(B, dtype=torch.float64), self.empty(B, dtype=torch.int32),
                              self.empty(B, dtype=torch.uint8))
            if trace:
                out.obj_fun_trace = torch.full((B, opts.max_iter), float("nan"), dtype=torch.float64, device=self.device)
                out.alpha_trace = torch.full((B, opts.max_iter), -2, dtype=torch.int8, device=self.device)
        _native.check(self.lib.crl_solve(prob, ctypes.byref(opts), _ptr(x0), _ptr(u0t), u0s, _ptr(J_init),
                                         _ptr(out.trajectory), _ptr(out.obj_fun_value), _ptr(out.iterations),
                                         _ptr(out.converged), _ptr(out.obj_fun_trace), _ptr(out.alpha_trace),
                                         _ptr(ws), ws.numel(), self._stream()), "crl_solve")
        self.launches += 1       # one persistent kernel (+ a 256-byte memset of the work counter)
        # keep inputs alive until the stream has consumed them
        out._keepalive = (x0, params, u0t, J_init)
        return out

    def solve_stages(self, opts: _native.SolverOpts, x0, stage_params, u0=None, J_init=None, return_trajectory=True):
        """crl_solve_stages: every trajectory runs one inner loop per stage row inside ONE persistent kernel.
        stage_params [B, S, p] (or [1, S, p] shared).  Returns (BatchResult, stage_iterations [B, S] int32)."""
        x0 = self.real(x0).reshape(-1, self.n)
        B = x0.shape[0]
        stage_params = self.real(stage_params)
        if stage_params.dim() not in (3, 4) or stage_params.shape[-1] != self.p or stage_params.shape[0] not in (1, B):
            raise ValueError("stage_params must be [B, S, p] / [1, S, p], or [B, S, T, p] / [1, S, T, p] (generated models)")
        if stage_params.dim() == 4 and (not self.generic or stage_params.shape[2] != self.T):
            raise ValueError("time-varying stage rows [.., S, T, p] are for the generated models")
        S = int(stage_params.shape[1])
        per_b = S * self.p * (self.T if stage_params.dim() == 4 else 1)
        sb = 0 if (stage_params.shape[0] == 1 and B != 1) else per_b
        st = self.p if stage_params.dim() == 4 else 0
        prob = _native.Problem(self.model_id, self.dtype_id, B, self.T, st, sb, stage_params.data_ptr(), self.u_lo, self.u_hi,
                               self._bounds_ptr())
        u0t, u0s = self._u0(u0, B)
        ws = self.workspace(prob, opts)
        if J_init is not None:
            J_init = torch.as_tensor(J_init, dtype=torch.float64).to(self.device).reshape(-1)
            if J_init.numel() == 1:
                J_init = J_init.expand(B)
            J_init = J_init.contiguous()
        out = BatchResult(self.empty(B, self.T, self.d) if return_trajectory else None,
                          self.empty(B, dtype=torch.float64), self.empty(B, dtype=torch.int32),
                          self.empty(B, dtype=torch.uint8))
        stage_iters = torch.zeros((B, S), dtype=torch.int32, device=self.device)
        _native.check(self.lib.crl_solve_stages(prob, ctypes.byref(opts), S, _ptr(x0), _ptr(u0t), u0s, _ptr(J_init),
                                                _ptr(out.trajectory), _ptr(out.obj_fun_value), _ptr(out.iterations),
                                                _ptr(out.converged), _ptr(stage_iters), _ptr(ws), ws.numel(),
                                                self._stream()), "crl_solve_stages")
        self.launches += 1       # one persistent kernel for the whole barrier schedule
        out._keepalive = (x0, stage_params, u0t, J_init)
        return out, stage_iters
