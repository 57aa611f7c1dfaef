"""iLQR solver objects with the reference's Python surface, executing on the B200.

Same class names, constructor arguments and method signatures as the reference
(/root/reference/CuriousRL/algorithm/ilqr_solver/):

* `iLQRDynamicModel`      eval_traj / update_traj / eval_grad_dynamic_model, T/m/n   (ilqr_dynamic_model.py:25-94)
* `iLQRObjectiveFunction` eval_obj_fun / eval_grad_obj_fun / eval_hessian_obj_fun    (ilqr_obj_fun.py:43-84)
* `iLQRWrapper`           forward_pass / backward_pass / get_*/set_*                 (ilqr_wrapper.py:176-296)
* `BasiciLQR`             __init__ / init / solve                                    (basic_ilqr.py:13-104)

For one problem (B = 1) arrays go in and out as NumPy with the reference's shapes
([T, n+m, 1] trajectories, [T, m, n] gains ...), so the Example scripts and `scenario.play()`
work unchanged.  The batched extension (`solve_batch`, and a leading batch axis accepted by
every evaluator method) takes/returns torch CUDA tensors without host copies.

Nothing here computes on the CPU: each method is one call into the C ABI.
"""
from __future__ import annotations

import time as _time
from typing import Optional

import numpy as np
import torch

from ... import _native
from ...utils.Logger import logger
from ..algo_wrapper import AlgoWrapper
from .device import BatchResult, DeviceModel


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.float64).numpy()


class iLQRDynamicModel:
    """Dynamics evaluator bound to a device model (reference: ilqr_dynamic_model.py:4-94)."""

    def __init__(self, dev: DeviceModel, init_state, init_action, owner):
        self._dev = dev
        self._init_state = init_state          # array [n, 1], replaced by set_init_state()
        self._init_action = init_action        # array [T, m, 1]
        self._owner = owner                    # the solver (source of the current add_param)
        self._n, self._m, self._T = dev.n, dev.m, dev.T

    T = property(lambda self: self._T)
    m = property(lambda self: self._m)
    n = property(lambda self: self._n)

    def eval_traj(self, init_state=None, action_traj=None):
        """Open-loop rollout.  NumPy in -> [T, n+m, 1] NumPy out; torch [B, n] in -> [B, T, n+m] torch out."""
        x0 = self._init_state if init_state is None else init_state
        u0 = self._init_action if action_traj is None else action_traj
        if isinstance(x0, torch.Tensor):
            X, _ = self._dev.rollout(x0, self._owner._params_for(x0.reshape(-1, self._n).shape[0]), u0)
            return X
        X, _ = self._dev.rollout(np.asarray(x0, dtype=np.float64).reshape(1, self._n), self._owner._params_for(1),
                                 np.asarray(u0, dtype=np.float64).reshape(self._T, self._m))
        return _np(X)[0][:, :, None]

    def update_traj(self, old_traj, K_matrix_all, k_vector_all, alpha):
        """Closed-loop rollout for one step size (ilqr_dynamic_model.py:56-71)."""
        if isinstance(old_traj, torch.Tensor):
            B = old_traj.reshape(-1, self._T, self._dev.d).shape[0]
            return self._dev.update_traj(old_traj, K_matrix_all, k_vector_all, alpha, self._owner._params_for(B))[0]
        Xn, _ = self._dev.update_traj(np.asarray(old_traj, dtype=np.float64).reshape(1, self._T, -1), np.asarray(K_matrix_all),
                                      np.asarray(k_vector_all).reshape(1, self._T, self._m), alpha, self._owner._params_for(1))
        return _np(Xn)[0][:, :, None]

    def eval_grad_dynamic_model(self, trajectory):
        if isinstance(trajectory, torch.Tensor):
            B = trajectory.reshape(-1, self._T, self._dev.d).shape[0]
            return self._dev.derivs(trajectory, self._owner._params_for(B), want=("F",))[0]
        F = self._dev.derivs(np.asarray(trajectory, dtype=np.float64).reshape(1, self._T, -1), self._owner._params_for(1),
                             want=("F",))[0]
        return _np(F)[0]


class iLQRObjectiveFunction:
    """Cost evaluator bound to a device model (reference: ilqr_obj_fun.py:5-84)."""

    def __init__(self, dev: DeviceModel, add_param, owner):
        self._dev = dev
        self._add_param = add_param            # array [T, p]; aliased with the scenario's (reference quirk Q8)
        self._owner = owner

    def _eval(self, trajectory, what):
        T, d = self._dev.T, self._dev.d
        if isinstance(trajectory, torch.Tensor):
            X = trajectory.reshape(-1, T, d)
            prm = self._owner._params_for(X.shape[0])
            return self._dev.cost(X, prm) if what == "J" else self._dev.derivs(X, prm, want=(what,))[("F", "c", "C").index(what)]
        X = np.asarray(trajectory, dtype=np.float64).reshape(1, T, d)
        prm = self._owner._params_for(1)
        if what == "J":
            return float(self._dev.cost(X, prm).item())
        out = _np(self._dev.derivs(X, prm, want=(what,))[("F", "c", "C").index(what)])[0]
        return out[:, :, None] if what == "c" else out

    def eval_obj_fun(self, trajectory):
        return self._eval(trajectory, "J")

    def eval_grad_obj_fun(self, trajectory):
        return self._eval(trajectory, "c")

    def eval_hessian_obj_fun(self, trajectory):
        return self._eval(trajectory, "C")


class iLQRWrapper(AlgoWrapper):
    """Line search, stopping tests, forward/backward passes and state accessors (ilqr_wrapper.py:12-296)."""

    def __init__(self, stopping_criterion, max_line_search, gamma, line_search_method, stopping_method, **kwargs):
        if line_search_method not in _native.LINE_SEARCH_IDS:
            raise ValueError('line_search_method must be "vanilla", "feasibility" or "none"')
        if stopping_method not in _native.STOPPING_IDS:
            raise ValueError('stopping_method must be "vanilla" or "relative"')
        super().__init__(stopping_criterion=stopping_criterion, max_line_search=max_line_search, gamma=gamma,
                         line_search_method=line_search_method, stopping_method=stopping_method, **kwargs)
        self._stopping_criterion = stopping_criterion
        self._max_line_search = max_line_search
        self._gamma = gamma
        self._line_search_method = line_search_method
        self._stopping_method = stopping_method
        self._obj_fun_value_last = np.inf
        self._dev: Optional[DeviceModel] = None
        self._dynamic_model = None
        self._obj_fun = None

    obj_fun = property(lambda self: self._obj_fun)
    dynamic_model = property(lambda self: self._dynamic_model)

    # -- helpers ---------------------------------------------------------------------------
    def _opts(self, max_iter=1, is_check_stop=True) -> _native.SolverOpts:
        return _native.SolverOpts(int(max_iter), int(bool(is_check_stop)), float(self._stopping_criterion),
                                  int(self._max_line_search), float(self._gamma),
                                  _native.LINE_SEARCH_IDS[self._line_search_method],
                                  _native.STOPPING_IDS[self._stopping_method], int(getattr(self, "_grid_blocks", 0)),
                                  int(getattr(self, "_occupancy", 0)), int(getattr(self, "_coop_threshold", 0)), 0)

    def _params_for(self, B: int) -> torch.Tensor:
        """The solver's current add_param ([T, p], reference semantics) as a device tensor shared by B problems."""
        ap = self._obj_fun._add_param
        if ap is None:                                   # scenario without add_param: zeros [T, 1] (ilqr_obj_fun.py:91-92)
            ap = np.zeros((self._dev.T, self._dev.p))
        return self._dev.real(np.asarray(ap, dtype=np.float64)).reshape(1, self._dev.T, self._dev.p)

    # -- the two passes (ilqr_wrapper.py:176-230) --------------------------------------------
    def forward_pass(self, trajectory, K_matrix, k_vector):
        """Line search + stopping test + derivatives at the new trajectory; updates the stored cost.

        NumPy: ([T,d,1], [T,m,n], [T,m,1]) -> (traj [T,d,1], C [T,d,d], c [T,d,1], F [T,n,d], J, is_stop).
        """
        dev, T = self._dev, self._dev.T
        X = np.asarray(trajectory, dtype=np.float64).reshape(1, T, dev.d)
        prm = self._params_for(1)
        Xn, Jn, _aidx, stop = dev.forward_pass(self._opts(), X, np.asarray(K_matrix), np.asarray(k_vector).reshape(1, T, dev.m),
                                               np.asarray([self._obj_fun_value_last], dtype=np.float64), prm)
        F, c, C = dev.derivs(Xn, prm)
        J = float(Jn.item())
        self._obj_fun_value_last = J
        return (_np(Xn)[0][:, :, None], _np(C)[0], _np(c)[0][:, :, None], _np(F)[0], J, bool(stop.item()))

    def backward_pass(self, C_matrix, c_vector, F_matrix):
        """Riccati recursion on caller-supplied matrices -> (K [T,m,n], k [T,m,1])."""
        dev = self._dev
        K, k = dev.backward_dense(np.asarray(F_matrix), np.asarray(c_vector).reshape(1, dev.T, dev.d), np.asarray(C_matrix))
        return _np(K)[0], _np(k)[0][:, :, None]

    # -- accessors (ilqr_wrapper.py:258-296) -------------------------------------------------
    def get_obj_fun_value(self):
        return self._obj_fun_value_last

    def set_obj_fun_value(self, obj_fun_value):
        self._obj_fun_value_last = obj_fun_value

    def set_init_state(self, new_state):
        self._dynamic_model._init_state = new_state

    def set_obj_add_param(self, new_add_param):
        self._obj_fun._add_param = new_add_param

    def get_obj_add_param(self):
        return self._obj_fun._add_param

    def _host_trajectory(self, B):
        """Pinned host buffer [B, T, n+m] for `solve_batch(trajectory_to_host=True)` (cached: pinning 0.6 GB takes longer
        than the solve)."""
        dev = self._dev
        buf = getattr(self, "_host_traj_buf", None)
        if buf is None or tuple(buf.shape) != (B, dev.T, dev.d) or buf.dtype != dev.dtype:
            buf = torch.empty((B, dev.T, dev.d), dtype=dev.dtype, pin_memory=True)
            self._host_traj_buf = buf
        return buf

    # -- batched re-solve (receding horizon) ---------------------------------------------------
    def resolve_batch(self, previous, at_step, add_param=None, warm_start=False, **solve_kwargs):
        """Batched form of the Examples' re-solve (Example/ThreeLinkPlanarManipulatorInteractive.py:18-31: new target,
        `set_obj_fun_value(inf)`, `set_init_state(<state the arm is in>)`, `solve()`): every problem restarts from the
        state its `previous` solution (a BatchResult with trajectories) reaches at `at_step` (an int, or one int per
        problem), with `add_param` as its new target (None = as `solve_batch`) and the stored cost at +inf.

        warm_start=False restarts from the scenario's `init_action` like the reference; warm_start=True starts from
        the remaining actions of the previous solution shifted by `at_step` and padded with zeros (receding
        horizon).  Everything stays on the device."""
        dev = self._dev
        X = previous.trajectory
        if X is None:
            raise ValueError("resolve_batch needs the trajectories of the previous solve (return_trajectory=True)")
        B = X.shape[0]
        steps = torch.as_tensor(at_step, device=X.device, dtype=torch.long).reshape(-1)
        if steps.numel() == 1:
            steps = steps.expand(B)
        if steps.numel() != B or int(steps.min()) < 0 or int(steps.max()) >= dev.T:
            raise ValueError("at_step must be an int or [B] ints in [0, T)")
        x0 = X[torch.arange(B, device=X.device), steps, :dev.n].contiguous()
        u0 = None
        if warm_start:
            t_idx = torch.arange(dev.T, device=X.device).reshape(1, dev.T) + steps.reshape(B, 1)        # [B, T]
            valid = (t_idx < dev.T).unsqueeze(-1)
            U = X[:, :, dev.n:]
            u0 = torch.where(valid, torch.gather(U, 1, t_idx.clamp(max=dev.T - 1).unsqueeze(-1).expand(B, dev.T, dev.m)),
                             torch.zeros((), dtype=U.dtype, device=U.device)).contiguous()
        return self.solve_batch(x0, add_param, init_action=u0, **solve_kwargs)


def _publish_batch(algo, res, **extra):
    """Push a batched result to the logger's batched channel (the counterpart of `logger.save_to_json(trajectory=...)`
    at the end of `solve()`); by reference unless `logger.set_is_save_json(True)`."""
    fields = dict(trajectory=res.trajectory, obj_fun_value=res.obj_fun_value, iterations=res.iterations,
                  converged=res.converged)
    fields.update(extra)
    logger.save_batch(meta={"solver": type(algo).__name__, "model": algo._dev.model_name, "T": algo._dev.T}, **fields)


class BasiciLQR(iLQRWrapper):
    """Classical iLQR (basic_ilqr.py:12-104) on the GPU.

    Extra, defaulted constructor arguments: `dtype` ("float64" | "float32": arithmetic of the
    rollouts and the Riccati pass; costs are always accumulated in float64), `device` (torch
    device, default: current CUDA device), `verify` (check the scenario's symbolic definition
    against the device model at init), `grid_blocks` (persistent-kernel grid, 0 = auto),
    `coop_threshold` (scheduling of the tail of a batched solve, see include/curiousrl_b200.h;
    results do not depend on it).
    """

    def __init__(self, max_iter=1000, is_check_stop=True, stopping_criterion=1e-6, max_line_search=50, gamma=0.5,
                 line_search_method="vanilla", stopping_method="relative", dtype="float64", device=None, verify=True,
                 grid_blocks=0, occupancy=0, coop_threshold=0):
        super().__init__(stopping_criterion=stopping_criterion, max_line_search=max_line_search, gamma=gamma,
                         line_search_method=line_search_method, stopping_method=stopping_method,
                         max_iter=max_iter, is_check_stop=is_check_stop)
        self._max_iter = max_iter
        self._is_check_stop = is_check_stop
        self._dtype = dtype
        self._device = device
        self._verify = verify
        self._grid_blocks = grid_blocks
        self._occupancy = occupancy      # resident CTAs/SM variant of the solve kernel (0 = default)
        self._coop_threshold = coop_threshold
        self._trajectory = None
        self.last_result: Optional[BatchResult] = None

    def init(self, scenario) -> "BasiciLQR":
        if not scenario.with_model() or scenario.is_action_discrete() or scenario.is_output_image():
            raise Exception('Scenario "' + scenario.name + '" cannot learn with BasiciLQR')
        constr = np.asarray(scenario.constr, dtype=np.float64)
        n = int(scenario.init_state.shape[0])
        T = int(scenario.init_action.shape[0])
        if scenario.name in _native.GENERIC_MODEL_IDS:
            # generated device model: the whole box - state bounds included - is compiled in
            # the scenario's default box is compiled in; any other one (is_with_constraints=False) is passed per call
            compiled = np.asarray(_native.model_bounds(_native.GENERIC_MODEL_IDS[scenario.name]), dtype=np.float64)
            if compiled.shape != constr.shape:
                raise Exception('Scenario "%s" does not have the shape of the device model' % scenario.name)
            self._dev = DeviceModel(scenario.name, T, 0.0, 0.0, dtype=self._dtype, device=self._device,
                                    bounds=None if np.array_equal(compiled, constr) else constr)
        elif scenario.name not in _native.MODEL_IDS:
            # any other DynamicModelWrapper (basic_ilqr.py:50-66 takes them all): its device model is generated from the
            # scenario's sympy expressions and compiled now (curiousrl_b200/jit.py; cached in-tree); box compiled in
            self._dev = DeviceModel(scenario.name, T, 0.0, 0.0, dtype=self._dtype, device=self._device, scenario=scenario)
        else:
            if not (np.all(constr[:n, 0] == -np.inf) and np.all(constr[:n, 1] == np.inf)):
                raise Exception('Scenario "%s": finite state bounds are not supported by the device models' % scenario.name)
            if not (np.all(constr[n:, 0] == constr[n, 0]) and np.all(constr[n:, 1] == constr[n, 1])):
                raise Exception('Scenario "%s": per-action bounds must be identical' % scenario.name)
            self._dev = DeviceModel(scenario.name, T, constr[n, 0], constr[n, 1], dtype=self._dtype, device=self._device)
        if (self._dev.n, self._dev.m) != (n, int(len(scenario.x_u_var)) - n):
            raise Exception('Scenario "%s" does not have the shape of the device model' % scenario.name)
        self._dynamic_model = iLQRDynamicModel(self._dev, scenario.init_state, scenario.init_action, self)
        self._obj_fun = iLQRObjectiveFunction(self._dev, scenario.add_param, self)
        if self._verify:
            if self._dev.generic:
                self._verify_generic(scenario)
            else:
                self._verify_against(scenario)
        return self

    def _verify_generic(self, scenario):
        """Generated device model against the scenario's own sympy f and cost: random trajectories inside the box,
        cost with the scenario's (time-varying) add_param, one closed-loop step with zero gains for the dynamics."""
        import sympy as sp
        dev, rng = self._dev, np.random.default_rng(12345)
        xu = scenario.x_u_var
        pv = scenario.add_param_var if scenario.add_param_var is not None else sp.symbols("no_use")
        f_expr = scenario.dynamic_function
        f = sp.lambdify([xu], f_expr.tolist() if hasattr(f_expr, "tolist") else f_expr, "math")
        l = sp.lambdify([xu, pv], scenario.obj_fun, "math")
        constr = np.asarray(scenario.constr, dtype=np.float64)
        lo, hi = np.maximum(constr[:, 0], -0.5), np.minimum(constr[:, 1], 0.5)
        P = _np(self._params_for(1))[0]
        for _ in range(3):
            X = rng.uniform(lo, hi, (1, dev.T, dev.d))
            J = float(dev.cost(X, self._params_for(1)).item())
            J_ref = sum(float(l(X[0, t], P[t] if scenario.add_param_var is not None else 0.0)) for t in range(dev.T))
            if not np.isclose(J, J_ref, rtol=1e-9):
                raise Exception('Scenario "%s": obj_fun differs from the device model' % scenario.name)
            if dev.T > 1:
                Xn, _ = dev.update_traj(X, np.zeros((1, dev.T, dev.m, dev.n)), np.zeros((1, dev.T, dev.m)), 0.0,
                                        self._params_for(1))
                x1 = np.clip(np.asarray(f(X[0, 0]), dtype=np.float64), constr[:dev.n, 0], constr[:dev.n, 1])
                if not np.allclose(_np(Xn)[0, 1, :dev.n], x1, rtol=1e-9, atol=1e-9):
                    raise Exception('Scenario "%s": dynamic_function differs from the device model' % scenario.name)

    def _verify_against(self, scenario):
        """Evaluate the scenario's own sympy f and cost at random points and compare with the device model."""
        import sympy as sp
        dev, rng = self._dev, np.random.default_rng(12345)
        xu = scenario.x_u_var
        f_expr = scenario.dynamic_function
        f_expr = f_expr.tolist() if hasattr(f_expr, "tolist") else f_expr
        f = sp.lambdify([xu], f_expr, "math")
        l = sp.lambdify([xu, scenario.add_param_var], scenario.obj_fun, "math")
        tol = 1e-9 if self._dtype == "float64" else 1e-3
        hi = min(50.0, dev.u_hi)
        for _ in range(3):
            x0 = rng.uniform(-2.0, 2.0, dev.n)
            u0 = rng.uniform(-hi, hi, dev.m)
            tgt = rng.uniform(-3.0, 3.0, dev.p)
            U = np.zeros((dev.T, dev.m))
            U[0] = u0
            X, J = dev.rollout(x0[None], dev.real(tgt), U)
            Xh = _np(X)[0]
            x1 = np.asarray(f(np.concatenate([x0, u0])), dtype=np.float64)
            if dev.T > 1 and not np.allclose(Xh[1, :dev.n], x1, rtol=tol, atol=tol):
                raise Exception('Scenario "%s": dynamic_function differs from the device model' % scenario.name)
            J_ref = sum(float(l(Xh[t], tgt)) for t in range(dev.T))
            if not np.isclose(float(J.item()), J_ref, rtol=max(tol, 1e-6) if self._dtype == "float32" else 1e-9):
                raise Exception('Scenario "%s": obj_fun differs from the device model' % scenario.name)

    # ------------------------------------------------------------------ solve (B = 1, reference semantics)
    def _solve_recording(self):
        """`solve()` with one record per iteration in the logger's JSON channel, as the reference writes them
        (basic_ilqr.py:81-95: `logger.save_to_json(trajectory=...)` inside the loop, replayed by
        `read_from_json(folder, no_iter)` / `scenario.play(folder, no_iter)`): the loop runs on the host with one
        backward and one forward-pass call per iteration.  Same device functions as the fused kernel, hence the same
        numbers (tests/test_gpu_solve_parity.py::test_recorded_solve_equals_the_fused_one)."""
        dev = self._dev
        start = _time.time()
        x0 = np.asarray(self._dynamic_model._init_state, dtype=np.float64).reshape(1, dev.n)
        u0 = np.asarray(self._dynamic_model._init_action, dtype=np.float64)
        prm = self._params_for(1)
        X, J0 = dev.rollout(x0, prm, u0 if u0.any() else None)
        logger.info("[+ +] Initial Obj.Val.: %.5e" % float(J0.item()))
        J_last = torch.tensor([self._obj_fun_value_last], dtype=torch.float64, device=dev.device)
        opts = self._opts()
        for i in range(self._max_iter):
            K, k = dev.backward(X, prm)
            X, J_last, _aidx, stop = dev.forward_pass(opts, X, K, k, J_last, prm)
            self._trajectory = _np(X)[0][:, :, None]
            logger.info("[+ +] Iter.No.%3d   Obj.Val.:%.5e" % (i, float(J_last.item())))
            logger.save_to_json(trajectory=self._trajectory.tolist())
            if bool(stop.item()) and self._is_check_stop:
                break
        self._obj_fun_value_last = float(J_last.item())
        self.last_result = None
        logger.info("[+ +] Completed! All Time:%.5e" % (_time.time() - start))

    def solve(self, record_iterations=None):
        """Solve the scenario's problem.  Returns None; results are left in `_trajectory`,
        `get_obj_fun_value()` and the logger's JSON channel, as in basic_ilqr.py:68-97.

        record_iterations: True = one JSON record per iteration like the reference (host loop, two kernels per
        iteration); False = the whole solve in one kernel and one record for the final trajectory.  Default: True when the
        logger saves its records to disk (`logger.set_is_save_json(True)` - the per-iteration replay is what those
        files are for), False otherwise (in memory only the last record can be read back anyway)."""
        if record_iterations is None:
            record_iterations = bool(logger.IS_SAVE_JSON)
        if record_iterations:
            if self._dev.dtype_name != "float64":
                raise ValueError("record_iterations needs dtype='float64'")
            return self._solve_recording()
        dev = self._dev
        start = _time.time()
        x0 = np.asarray(self._dynamic_model._init_state, dtype=np.float64).reshape(1, dev.n)
        u0 = np.asarray(self._dynamic_model._init_action, dtype=np.float64)
        res = dev.solve(self._opts(self._max_iter, self._is_check_stop), x0, self._params_for(1),
                        u0=u0 if u0.any() else None,
                        J_init=np.asarray([self._obj_fun_value_last], dtype=np.float64), trace=True)
        host = res.to_host()
        iters = int(host.iterations[0])
        for i in range(iters):
            logger.info("[+ +] Iter.No.%3d   Obj.Val.:%.5e" % (i, host.obj_fun_trace[0, i]))
        self._trajectory = host.trajectory[0].astype(np.float64)[:, :, None]
        self._obj_fun_value_last = float(host.obj_fun_value[0])
        self.last_result = res
        logger.save_to_json(trajectory=self._trajectory.tolist())     # what scenario.play() reads
        logger.info("[+ +] Completed! All Time:%.5e" % (_time.time() - start))

    # ------------------------------------------------------------------ batched extension
    def solve_batch(self, init_state, add_param=None, init_action=None, obj_fun_value=None, return_trajectory=True,
                    trace=False, out: Optional[BatchResult] = None, trajectory_to_host=False) -> BatchResult:
        """Solve B independent problems in one persistent kernel.

        trajectory_to_host=True: the kernel writes every finished trajectory straight into PINNED HOST memory (mapped
        into the device's address space; 256-byte coalesced stores over PCIe) while the other trajectories are still
        being iterated - the device-to-host copy of the [B, T, n+m] result overlaps the solve instead of following it.
        `result.trajectory` is then a pinned CPU tensor, complete once the stream has been synchronised; the buffer
        is owned by the solver and reused by the next call with the same batch size.

        init_state [B, n]; add_param [B, p] or [B, 1, p] (constant target per problem), [B, T, p], [T, p] or
        [1, T, p] (one time-varying table shared by the batch; [T, p] is refused when B == T because it cannot be
        told from [B, p]) or None (the scenario's); init_action None (the scenario's), [T, m] or [B, T, m]; obj_fun_value: initial
        "last cost" (scalar or [B]; default +inf, i.e. `set_obj_fun_value(np.inf)` for every problem).
        NumPy or torch inputs; outputs are torch tensors on the solver's device.
        """
        dev = self._dev
        x0 = dev.real(init_state).reshape(-1, dev.n)
        B = x0.shape[0]
        if add_param is None:
            prm = self._params_for(B)
        else:
            prm = dev.real(add_param)
            if prm.dim() == 2 and tuple(prm.shape) == (dev.T, dev.p):
                if B == dev.T:
                    raise ValueError("add_param of shape [%d, %d] is ambiguous when the batch size equals the horizon: pass "
                                     "[1, T, p] for one time-varying table shared by the batch or [B, 1, p] / [B, T, p] for "
                                     "per-problem parameters" % (dev.T, dev.p))
                prm = prm.reshape(1, dev.T, dev.p)
            elif prm.dim() == 3 and tuple(prm.shape) == (B, 1, dev.p):
                prm = prm.reshape(B, dev.p)                       # per-problem constants, spelled unambiguously
        u0 = self._dynamic_model._init_action if init_action is None else init_action
        if isinstance(u0, np.ndarray) and not u0.any():
            u0 = None                      # zeros: let the kernel skip the loads
        if trajectory_to_host and return_trajectory:
            if out is None:
                out = BatchResult(None, dev.empty(B, dtype=torch.float64), dev.empty(B, dtype=torch.int32),
                                  dev.empty(B, dtype=torch.uint8))
                if trace:
                    out.obj_fun_trace = torch.full((B, self._max_iter), float("nan"), dtype=torch.float64, device=dev.device)
                    out.alpha_trace = torch.full((B, self._max_iter), -2, dtype=torch.int8, device=dev.device)
            if out.trajectory is None or out.trajectory.device.type != "cpu":
                out.trajectory = self._host_trajectory(B)
        res = dev.solve(self._opts(self._max_iter, self._is_check_stop), x0, prm, u0=u0,
                        J_init=obj_fun_value,
                        return_trajectory=return_trajectory, trace=trace, out=out)
        self.last_result = res
        _publish_batch(self, res)
        return res
