"""LogBarrieriLQR with the reference's Python surface, executing on the B200.

Reference: /root/reference/CuriousRL/algorithm/ilqr_solver/log_barrier_ilqr.py:13-156 - the solver
Example/RoboticArmTrackingIteractive.py:11 uses.  Every finite bound of the scenario adds
`(-1/t) log(-(lo - x_u[i]))` / `(-1/t) log(-(x_u[i] - hi))` to the running cost (:91-95), the barrier
parameter t is appended to `add_param` as an extra column (:96-101), the line search is the
"feasibility" one (:21) and `solve()` runs one inner iLQR loop per t in `t = [0.5, ..., 100]`, resetting
the stored cost to +inf after each converged inner loop (:121-146).

Device side: the same kernels as BasiciLQR, instantiated for the barrier cost (`Barrier<Model>` in
csrc/ilqr_models.cuh, model ids CRL_*_BARRIER) - non-constant action entries of the cost gradient and
Hessian inside the fused Riccati pass, `log` terms in every rollout's cost.  One inner loop = one
`crl_solve` launch warm-started with the previous loop's actions.  FP64 only.

Quirks kept:
* the FIRST backward pass after t moves on still uses cost derivatives evaluated with the previous t
  (the reference computes C, c in the forward pass, before the outer loop updates add_param): the device
  parameter row carries that stale value as its last entry;
* the first inner loop never writes t (`if j != self._t[0]`, :122), so a second `solve()` on the same
  object starts with the t the previous solve ended on;
* `get_obj_add_param()` returns the [T, p + 1] array whose last column is t.
"""
from __future__ import annotations

import time as _time
from typing import List, Optional

import numpy as np
import torch

from ... import _native
from ...utils.Logger import logger
from .device import BatchResult, DeviceModel
from .solver import iLQRDynamicModel, iLQRObjectiveFunction, iLQRWrapper, _np, _publish_batch


class _RealCostOwner:
    """Parameter source of the barrier-free cost evaluator: the first p columns of the solver's add_param."""

    def __init__(self, algo):
        self._algo = algo

    def _params_for(self, B: int) -> torch.Tensor:
        dev = self._algo._real_dev
        if self._algo._real_add_param is None:            # scenario without add_param: one unused zero column
            return dev.real(np.zeros((dev.T, dev.p))).reshape(1, dev.T, dev.p)
        ap = np.asarray(self._algo._obj_fun._add_param, dtype=np.float64)[:, :dev.p]
        return dev.real(ap).reshape(1, dev.T, dev.p)


class LogBarrierBatchResult(BatchResult):
    """BatchResult + `inner_iterations` [B, len(t)] (iterations of every inner loop) and `real_obj_fun_value` [B]
    (the barrier-free cost of the final trajectories, what the reference reports, :138)."""
    inner_iterations: Optional[torch.Tensor] = None
    real_obj_fun_value: Optional[torch.Tensor] = None


class LogBarrieriLQR(iLQRWrapper):
    """Log-barrier iLQR (log_barrier_ilqr.py:13-156) on the GPU.  Constructor arguments as in the reference;
    `device`, `verify`, `grid_blocks`, `occupancy`, `coop_threshold` as for BasiciLQR."""

    def __init__(self, max_iter=1000, is_check_stop=True, stopping_criterion=1e-6, max_line_search=50, gamma=0.5,
                 t=[0.5, 1., 2., 5., 10., 20., 50., 100.], line_search_method="feasibility", stopping_method="relative",
                 device=None, verify=True, grid_blocks=0, occupancy=0, coop_threshold=0):
        super().__init__(stopping_criterion=stopping_criterion, max_line_search=max_line_search, gamma=gamma,
                         line_search_method=line_search_method, stopping_method=stopping_method,
                         max_iter=max_iter, is_check_stop=is_check_stop)
        self._t = list(t)
        self._max_iter = max_iter
        self._is_check_stop = is_check_stop
        self._device = device
        self._verify = verify
        self._grid_blocks = grid_blocks
        self._occupancy = occupancy          # scheduling knobs as for BasiciLQR; results do not depend on them
        self._coop_threshold = coop_threshold
        self._trajectory = None
        self._t_stale = None                 # the t the next first backward pass still sees
        self.last_result: Optional[BatchResult] = None
        self.inner_iterations: List[int] = []

    # ------------------------------------------------------------------ init
    def init(self, scenario) -> "LogBarrieriLQR":
        if not scenario.with_model() or scenario.is_action_discrete() or scenario.is_output_image():
            raise Exception('Scenario "' + scenario.name + '" cannot learn with LogBarrieriLQR')
        constr = np.asarray(scenario.constr, dtype=np.float64)
        n = int(scenario.init_state.shape[0])
        T = int(scenario.init_action.shape[0])
        self._generic = scenario.name in _native.GENERIC_MODEL_IDS
        jit = not self._generic and scenario.name not in _native.MODEL_IDS
        if jit:
            # any other scenario: plain and barrier models are generated from its sympy expressions and compiled now
            # (curiousrl_b200/jit.py); the barrier sits on every finite bound of the scenario's own box
            self._generic = True
            lo = hi = 0.0
        elif self._generic:
            # generated device model: the barrier sits on every finite bound of the compiled-in box, states included
            compiled = np.asarray(_native.model_bounds(_native.GENERIC_MODEL_IDS[scenario.name]), dtype=np.float64)
            if compiled.shape != constr.shape or not np.array_equal(compiled, constr):
                raise Exception('Scenario "%s": the barrier cost is generated for constr = %s' % (scenario.name, compiled.tolist()))
            lo = hi = 0.0
        else:
            if not (np.all(constr[:n, 0] == -np.inf) and np.all(constr[:n, 1] == np.inf)):
                raise Exception('Scenario "%s": finite state bounds are not supported by the device models' % scenario.name)
            if scenario.name not in _native.MODEL_IDS:
                raise Exception('Scenario "%s" has no device model in curiousrl_b200; there is no CPU fallback.' % scenario.name)
            lo, hi = _native.model_default_bounds(_native.MODEL_IDS[scenario.name])
            if not (np.all(constr[n:, 0] == lo) and np.all(constr[n:, 1] == hi)):
                raise Exception('Scenario "%s": the barrier cost is compiled for the action box [%g, %g]' % (scenario.name, lo, hi))
        self._dev = DeviceModel(scenario.name, T, lo, hi, device=self._device, barrier=True, scenario=scenario if jit else None)
        self._real_dev = DeviceModel(scenario.name, T, lo, hi, device=self._device, scenario=scenario if jit else None)
        self._dynamic_model = iLQRDynamicModel(self._dev, scenario.init_state, scenario.init_action, self)
        self._real_add_param = scenario.add_param          # None for a scenario without add_param
        if scenario.add_param is None:
            if not self._generic:
                raise Exception('Scenario "%s" has no add_param (target)' % scenario.name)
            add_param = self._t[0] * np.ones((T, 1), dtype=np.float64)                                            # :96-98
        else:
            add_param = np.hstack([np.asarray(scenario.add_param, dtype=np.float64), self._t[0] * np.ones((T, 1))])   # :100-101
        self._obj_fun = iLQRObjectiveFunction(self._dev, add_param, self)
        self._real_obj_fun = iLQRObjectiveFunction(self._real_dev, None, _RealCostOwner(self))
        self._t_stale = None
        if self._verify:
            # the device model is chosen by the scenario's NAME: check the scenario's own sympy definition against it, so
            # that a subclass / edited scenario of the same name cannot silently run the compiled default expressions
            if self._generic:
                self._verify_generic(scenario)
            self._verify_against(scenario)
        return self

    def _verify_generic(self, scenario):
        """Generated device model under the barrier cost: dynamics (one closed-loop step with zero gains) against the
        scenario's sympy f.  The cost + barrier terms are checked by _verify_against."""
        import sympy as sp
        dev, rng = self._real_dev, np.random.default_rng(2468)
        xu = scenario.x_u_var
        f_expr = scenario.dynamic_function
        f = sp.lambdify([xu], f_expr.tolist() if hasattr(f_expr, "tolist") else f_expr, "math")
        constr = np.asarray(scenario.constr, dtype=np.float64)
        lo, hi = np.maximum(constr[:, 0], -0.5), np.minimum(constr[:, 1], 0.5)
        prm = _RealCostOwner(self)._params_for(1)
        if dev.T > 1:
            for _ in range(3):
                X = rng.uniform(lo, hi, (1, dev.T, dev.d))
                Xn, _ = dev.update_traj(X, np.zeros((1, dev.T, dev.m, dev.n)), np.zeros((1, dev.T, dev.m)), 0.0, prm)
                x1 = np.clip(np.asarray(f(X[0, 0]), dtype=np.float64), constr[:dev.n, 0], constr[:dev.n, 1])
                if not np.allclose(Xn.cpu().numpy()[0, 1, :dev.n], x1, rtol=1e-9, atol=1e-9):
                    raise Exception('Scenario "%s": dynamic_function differs from the device model' % scenario.name)

    def _verify_against(self, scenario):
        """The barrier cost of the device model against the scenario's own symbolic cost + barrier terms."""
        import sympy as sp
        dev, rng = self._dev, np.random.default_rng(4321)
        xu, t_var = scenario.x_u_var, sp.symbols("t")
        expr = scenario.obj_fun
        for i, c in enumerate(np.asarray(scenario.constr, dtype=np.float64)):
            if not np.isinf(c[0]):
                expr = expr + (-1 / t_var) * sp.log(-(c[0] - xu[i]))
            if not np.isinf(c[1]):
                expr = expr + (-1 / t_var) * sp.log(-(xu[i] - c[1]))
        if self._generic:
            # the scenario's own (time-varying) parameter rows + t, states and actions inside the box (the barrier of a
            # generated model sits on the state bounds too)
            pv = tuple(scenario.add_param_var) if scenario.add_param_var is not None else ()
            l = sp.lambdify([xu, (*pv, t_var)], expr, "math")
            constr = np.asarray(scenario.constr, dtype=np.float64)
            lo, hi = np.maximum(constr[:, 0], -0.5), np.minimum(constr[:, 1], 0.5)
            P = self._params_for(1)
            Ph = P.cpu().numpy()[0]
            for _ in range(3):
                X = rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), (1, dev.T, dev.d))
                J = float(dev.cost(X, P).item())
                J_ref = sum(float(l(X[0, t], Ph[t, :len(pv) + 1])) for t in range(dev.T))
                if not np.isclose(J, J_ref, rtol=1e-9):
                    raise Exception('Scenario "%s": obj_fun + barrier differs from the device model' % scenario.name)
            return
        l = sp.lambdify([xu, (*scenario.add_param_var, t_var)], expr, "math")
        for _ in range(3):
            X = rng.uniform(-2.0, 2.0, (1, dev.T, dev.d))
            X[:, :, dev.n:] = rng.uniform(0.9 * dev.u_lo, 0.9 * dev.u_hi, (1, dev.T, dev.m))
            prm = np.concatenate([rng.uniform(-3.0, 3.0, dev.p - 2), [rng.uniform(0.5, 100.0)], [1.0]])
            J = float(dev.cost(X, dev.real(prm)).item())
            J_ref = sum(float(l(X[0, t], prm[:-1])) for t in range(dev.T))
            if not np.isclose(J, J_ref, rtol=1e-9):
                raise Exception('Scenario "%s": obj_fun + barrier differs from the device model' % scenario.name)

    # ------------------------------------------------------------------ parameters
    def _params_for(self, B: int) -> torch.Tensor:
        """Device parameter rows [1, T, p]: the reference's add_param (targets, t) + the stale t of the next first
        backward pass (= t unless an outer loop just moved it)."""
        ap = np.asarray(self._obj_fun._add_param, dtype=np.float64)
        stale = ap[:, -1:] if self._t_stale is None else np.full((ap.shape[0], 1), float(self._t_stale))
        return self._dev.real(np.hstack([ap, stale])).reshape(1, self._dev.T, self._dev.p)

    # ------------------------------------------------------------------ solve (B = 1, reference semantics)
    def solve(self):
        """One inner iLQR loop per barrier parameter (log_barrier_ilqr.py:104-149).  Returns None; results are left
        in `_trajectory`, `inner_iterations` and the logger's JSON channel."""
        dev = self._dev
        start = _time.time()
        x0 = np.asarray(self._dynamic_model._init_state, dtype=np.float64).reshape(1, dev.n)
        U = np.asarray(self._dynamic_model._init_action, dtype=np.float64).reshape(dev.T, dev.m)
        X0 = self._dynamic_model.eval_traj()
        logger.info("[+ +] Initial Obj.Val.: %.5e" % self._real_obj_fun.eval_obj_fun(X0))
        self.inner_iterations = []
        total = -1
        res = None
        for j in self._t:
            self._t_stale = None
            if j != self._t[0]:                                      # :122-125
                add_param = self.get_obj_add_param()
                self._t_stale = float(add_param[0, -1])              # what C, c were evaluated with
                add_param[:, -1] = j * np.ones((dev.T))
                self.set_obj_add_param(add_param)
            res = dev.solve(self._opts(self._max_iter, self._is_check_stop), x0, self._params_for(1), u0=U,
                            J_init=np.asarray([self._obj_fun_value_last], dtype=np.float64), trace=True)
            host = res.to_host()
            iters = int(host.iterations[0])
            for i in range(iters):
                total += 1
                logger.info("[+ +] Total Iter.No.%3d   Iter.No.%3d   Barrier Obj.Val.:%.5e" % (total, i, host.obj_fun_trace[0, i]))
            traj = host.trajectory[0].astype(np.float64)
            U = traj[:, dev.n:]
            self._trajectory = traj[:, :, None]
            self.inner_iterations.append(iters)
            self._obj_fun_value_last = float(host.obj_fun_value[0])
            logger.info("[+ +] Obj.Val.:%.5e" % self._real_obj_fun.eval_obj_fun(self._trajectory))
            logger.save_to_json(trajectory=self._trajectory.tolist())
            if bool(host.converged[0]) and self._is_check_stop:       # :140-146
                self.set_obj_fun_value(np.inf)
                logger.info("[+ +] Complete One Inner Loop! The log barrier parameter t is %.5f" % j + " in this iteration!")
        self._t_stale = None
        self.last_result = res
        logger.info("[+ +] Completed! All Time:%.5e" % (_time.time() - start))

    # ------------------------------------------------------------------ batched extension
    def solve_batch(self, init_state, add_param=None, init_action=None, return_trajectory=True,
                    fused=True) -> LogBarrierBatchResult:
        """B independent problems.  init_state [B, n]; add_param [B, p] constant targets (None = the scenario's);
        init_action None (the scenario's), [T, m] or [B, T, m].  Fresh problems: the stored cost starts at +inf.

        fused=True (default): the whole barrier schedule in ONE persistent kernel (`crl_solve_stages`) - every
        trajectory walks through the t's at its own pace, so the batch has one tail instead of one per t.
        fused=False: one `crl_solve` launch per t, every launch warm-started with the previous one's actions
        (the first implementation; same results bit for bit, kept for the equality test)."""
        dev, rdev = self._dev, self._real_dev
        x0 = dev.real(init_state).reshape(-1, dev.n)
        B = x0.shape[0]
        if self._generic:
            return self._solve_batch_generic(x0, add_param, init_action, return_trajectory, fused)
        if add_param is None:
            tgt = dev.real(np.asarray(self._obj_fun._add_param, dtype=np.float64)[0, :rdev.p]).reshape(1, rdev.p).expand(B, rdev.p)
        else:
            tgt = dev.real(add_param).reshape(B, rdev.p)
        u0 = self._dynamic_model._init_action if init_action is None else init_action
        if isinstance(u0, np.ndarray) and not u0.any():
            u0 = None
        opts = self._opts(self._max_iter, self._is_check_stop)
        S = len(self._t)
        if fused:
            # parameter rows [B, S, p]: (targets, t_s, t_{s-1}); the first loop's stale value is its own t
            ts = torch.tensor(self._t, dtype=dev.dtype, device=dev.device)
            prm = torch.empty((B, S, dev.p), device=dev.device, dtype=dev.dtype)
            prm[:, :, :rdev.p] = tgt.reshape(B, 1, rdev.p)
            prm[:, :, rdev.p] = ts
            prm[:, 0, rdev.p + 1] = ts[0]
            prm[:, 1:, rdev.p + 1] = ts[:-1]
            res, inner = dev.solve_stages(opts, x0, prm, u0=u0, J_init=None, return_trajectory=True)
        else:
            prm = torch.empty((B, dev.p), device=dev.device, dtype=dev.dtype)
            prm[:, :rdev.p] = tgt
            J_init = torch.full((B,), float("inf"), dtype=torch.float64, device=dev.device)
            inner = torch.zeros((B, S), dtype=torch.int32, device=dev.device)
            res = None
            t_prev = self._t[0]
            for s, t in enumerate(self._t):
                prm[:, rdev.p] = t
                prm[:, rdev.p + 1] = t_prev
                res = dev.solve(opts, x0, prm, u0=u0, J_init=J_init, return_trajectory=True)
                inner[:, s] = res.iterations
                u0 = res.trajectory[:, :, dev.n:].contiguous()
                if self._is_check_stop:
                    J_init = torch.where(res.converged.bool(), torch.full_like(res.obj_fun_value, float("inf")), res.obj_fun_value)
                else:
                    J_init = res.obj_fun_value.clone()
                t_prev = t
        out = LogBarrierBatchResult(res.trajectory if return_trajectory else None, res.obj_fun_value, inner.sum(dim=1).to(torch.int32),
                                    res.converged)
        out.inner_iterations = inner
        out.real_obj_fun_value = rdev.cost(res.trajectory, tgt.contiguous())
        self.last_result = out
        _publish_batch(self, out, inner_iterations=inner, real_obj_fun_value=out.real_obj_fun_value)
        return out

    def _solve_batch_generic(self, x0, add_param, init_action, return_trajectory, fused) -> LogBarrierBatchResult:
        """Generated device models: the scenario's own (time-varying) add_param is shared by the batch; stage rows
        [1, S, T, p] = (add_param columns, t_s, t_{s-1})."""
        if add_param is not None:
            raise ValueError("the generated-model scenarios take their cost parameters from the scenario (add_param=None)")
        dev, rdev = self._dev, self._real_dev
        B, S, T = x0.shape[0], len(self._t), dev.T
        base = np.asarray(self._obj_fun._add_param, dtype=np.float64)[:, :dev.p - 2]          # without the t column
        u0 = self._dynamic_model._init_action if init_action is None else init_action
        if isinstance(u0, np.ndarray) and not u0.any():
            u0 = None
        opts = self._opts(self._max_iter, self._is_check_stop)
        rows = np.empty((1, S, T, dev.p), dtype=np.float64)
        for s, t in enumerate(self._t):
            rows[0, s, :, :dev.p - 2] = base
            rows[0, s, :, dev.p - 2] = t
            rows[0, s, :, dev.p - 1] = self._t[max(s - 1, 0)]
        if fused:
            res, inner = dev.solve_stages(opts, x0, dev.real(rows), u0=u0, J_init=None, return_trajectory=True)
        else:
            J_init = torch.full((B,), float("inf"), dtype=torch.float64, device=dev.device)
            inner = torch.zeros((B, S), dtype=torch.int32, device=dev.device)
            res = None
            for s in range(S):
                res = dev.solve(opts, x0, dev.real(rows[:, s]), u0=u0, J_init=J_init, return_trajectory=True)
                inner[:, s] = res.iterations
                u0 = res.trajectory[:, :, dev.n:].contiguous()
                if self._is_check_stop:
                    J_init = torch.where(res.converged.bool(), torch.full_like(res.obj_fun_value, float("inf")), res.obj_fun_value)
                else:
                    J_init = res.obj_fun_value.clone()
        out = LogBarrierBatchResult(res.trajectory if return_trajectory else None, res.obj_fun_value, inner.sum(dim=1).to(torch.int32),
                                    res.converged)
        out.inner_iterations = inner
        out.real_obj_fun_value = rdev.cost(res.trajectory, _RealCostOwner(self)._params_for(B))
        self.last_result = out
        _publish_batch(self, out, inner_iterations=inner, real_obj_fun_value=out.real_obj_fun_value)
        return out
