"""CPU oracle of the batched-iLQR hot path.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; nothing under
`curiousrl_b200/` does.

It restates, function by function, what the reference computes on the path
(all citations relative to /root/reference/CuriousRL/algorithm/ilqr_solver/):

* `Problem.__init__`      symbolic differentiation + lambdify of f, df/d(x,u), l, grad l, hess l
                          (ilqr_dynamic_model.py:25-37, ilqr_obj_fun.py:43-52)
* `rollout`               open-loop rollout with the clamp-after-use order     (ilqr_dynamic_model.py:96-109)
* `update_traj`           closed-loop rollout for one step size alpha           (ilqr_dynamic_model.py:111-134)
* `cost`                  J = sum_t l(X[t], p_t), sequential float64 sum       (ilqr_obj_fun.py:86-95)
* `derivs`                F[t], c[t], C[t] (+1e-100 I)                         (ilqr_dynamic_model.py:136-147, ilqr_obj_fun.py:97-119)
* `backward`              Riccati recursion, general solve, no regularisation  (ilqr_wrapper.py:232-256)
* `line_search`           first improving alpha = 1, g, g^2, ...               (ilqr_wrapper.py:74-146)
* `is_stop`               relative / vanilla stopping tests                    (ilqr_wrapper.py:148-174)
* `solve`                 outer loop of BasiciLQR.solve                        (basic_ilqr.py:68-97)
* `BarrierScenario`       LogBarrieriLQR's cost: obj_fun + (-1/t) log terms, t appended to add_param
                          (log_barrier_ilqr.py:84-101)
* `solve_log_barrier`     LogBarrieriLQR.solve: one inner loop per barrier parameter  (log_barrier_ilqr.py:104-149)

The arithmetic substrate is the reference's own: the same sympy expressions go
through `sympy.lambdify` (printers "math" for the dynamics, "numpy" for the
cost, as in the reference), matrix products are NumPy `@` (BLAS dgemm) and the
gain solve is `numpy.linalg.solve` (LAPACK dgesv).  With `use_numba=True` the
loops and lambdas are wrapped in `numba.njit` exactly as the reference does,
which makes the oracle as fast as the reference (used for the CPU baseline);
without it everything is plain Python/NumPy (used by the unit tests, no JIT
start-up cost).

Parity pin: `tests/test_oracle_golden.py` checks this module against vectors
produced in the build container by the unmodified reference
(`oracle/make_golden.py`, fixtures in `tests/golden/`).

Trajectories here are [T, d] (d = n + m); the reference carries a trailing
singleton axis ([T, d, 1]).
"""
from __future__ import annotations

import numpy as np
import sympy as sp


def _lambdify(args, expr, printer):
    # sympy >= 1.13 refuses N-dim arrays in the strict printers; nested lists are
    # what older versions emitted and what the reference's np.asarray(...) expects.
    if isinstance(expr, (sp.NDimArray, np.ndarray)):
        expr = expr.tolist()
    return sp.lambdify(args, expr, printer)


class Problem:
    """Callable f, F, l, grad l, hess l of one scenario + its sizes and box constraints.

    `scenario` is any object with the reference's DynamicModelWrapper properties
    (dynamic_model.py:55-101) - the reference's own classes or curiousrl_b200's.
    """

    def __init__(self, scenario, use_numba: bool = False):
        xu = scenario.x_u_var
        self.name = scenario.name
        self.n = int(scenario.init_state.shape[0])
        self.m = int(len(xu)) - self.n
        self.d = self.n + self.m
        self.T = int(scenario.init_action.shape[0])
        self.constr = np.asarray(scenario.constr, dtype=np.float64)
        self.init_state = np.asarray(scenario.init_state, dtype=np.float64).reshape(-1)
        self.init_action = np.asarray(scenario.init_action, dtype=np.float64).reshape(self.T, self.m)
        unused = sp.symbols("no_use")
        pv = scenario.add_param_var
        if pv is None:                                   # ilqr_obj_fun.py:44-45
            pv = unused
        if scenario.add_param is None:                   # ilqr_obj_fun.py:91-92 (and :103-104, :115-116)
            self.add_param = np.zeros((self.T, 1))
        else:
            self.add_param = np.asarray(scenario.add_param, dtype=np.float64)
        f_sym = scenario.dynamic_function
        # ilqr_dynamic_model.py:33-35 - dynamics never see add_param
        f = _lambdify([xu, unused], f_sym, "math")
        F = _lambdify([xu, unused], sp.transpose(sp.derive_by_array(f_sym, xu)), "math")
        # ilqr_obj_fun.py:46-51
        l_sym = scenario.obj_fun
        g_sym = sp.derive_by_array(l_sym, xu)
        h_sym = sp.derive_by_array(g_sym, xu)
        l = _lambdify([xu, pv], l_sym, "numpy")
        g = _lambdify([xu, pv], g_sym, "numpy")
        H = _lambdify([xu, pv], np.asarray(h_sym) + 1e-100 * np.eye(h_sym.shape[0]), "numpy")
        self.use_numba = use_numba
        if use_numba:
            from numba import njit
            f, F, l, g, H = (njit(z) for z in (f, F, l, g, H))
            self._k = _numba_kernels()
        else:
            self._k = _PY_KERNELS
        self.f, self.F, self.l, self.g, self.H = f, F, l, g, H

    def params(self, target=None):
        """[T, p] cost parameters: the scenario's own, or a constant target broadcast over time."""
        if target is None:
            return self.add_param
        target = np.asarray(target, dtype=np.float64)
        if target.ndim == 1:
            return np.ascontiguousarray(np.broadcast_to(target, (self.T, target.shape[0])))
        return target


# ----------------------------------------------------------------------------- kernels
# Written once in a numba-compatible subset of Python; used as-is (pure Python) or jitted.

def _rollout(f, x0, U, n, m, constr):
    T = U.shape[0]
    X = np.zeros((T, n + m))
    dummy = np.zeros(1)
    X[0, :n] = x0
    X[0, n:] = U[0]
    for t in range(T - 1):
        X[t + 1, :n] = np.asarray(f(X[t], dummy), dtype=np.float64)   # uses the UNclamped row t
        X[t + 1, n:] = U[t + 1]
        for i in range(n + m):                                         # row t clamped afterwards
            X[t, i] = min(max(constr[i, 0], X[t, i]), constr[i, 1])
    return X


def _update_traj(f, X, K, k, alpha, n, m, constr):
    T = K.shape[0]
    Xn = np.zeros((T, n + m))
    dummy = np.zeros(1)
    Xn[0] = X[0]
    for t in range(T - 1):
        dx = Xn[t, :n] - X[t, :n]
        du = (K[t] @ dx.reshape(n, 1)).reshape(m) + alpha * k[t]     # 2-D @ 2-D like the reference (:123): BLAS dgemm, not dgemv
        Xn[t, n:] = X[t, n:] + du
        for i in range(m):                                              # action clamp before dynamics
            Xn[t, n + i] = min(max(constr[n + i, 0], Xn[t, n + i]), constr[n + i, 1])
        Xn[t + 1, :n] = np.asarray(f(Xn[t], dummy), dtype=np.float64)
        for i in range(n):                                              # state clamp after dynamics
            Xn[t + 1, i] = min(max(constr[i, 0], Xn[t + 1, i]), constr[i, 1])
    return Xn                                                           # Xn[T-1, n:] stays 0


def _cost(l, X, P):
    J = 0.0
    for t in range(X.shape[0]):
        J = J + l(X[t], P[t])
    return J


def _derivs(F, g, H, X, P, n):
    T, d = X.shape
    Fm = np.zeros((T, n, d))
    c = np.zeros((T, d))
    C = np.zeros((T, d, d))
    dummy = np.zeros(1)
    for t in range(T):
        Fm[t] = np.asarray(F(X[t], dummy), dtype=np.float64)
        c[t] = np.asarray(g(X[t], P[t]), dtype=np.float64)
        C[t] = np.asarray(H(X[t], P[t]), dtype=np.float64)
    return Fm, c, C


def _backward(Fm, c, C, n, m):
    T = Fm.shape[0]
    V = np.zeros((n, n))
    v = np.zeros((n, 1))
    K = np.zeros((T, m, n))
    k = np.zeros((T, m, 1))
    for t in range(T - 1, -1, -1):
        # transposes stay *views* (as in the reference): BLAS then runs its 'T' variants,
        # whose summation order differs from a contiguous copy at the last bit
        Q = C[t] + Fm[t].T @ V @ Fm[t]                 # (F^T V) F, left to right
        q = c[t] + Fm[t].T @ v
        Quu = Q[n:, n:].copy()
        Qux = Q[n:, :n].copy()
        qu = q[n:].copy()
        K[t] = -np.linalg.solve(Quu, Qux)              # general LU (dgesv); no regularisation
        k[t] = -np.linalg.solve(Quu, qu)
        V = Q[:n, :n] + Qux.T @ K[t] + K[t].T @ Qux + K[t].T @ Quu @ K[t]   # V is not re-symmetrised
        v = q[:n] + Qux.T @ k[t] + K[t].T @ qu + K[t].T @ Quu @ k[t]
    return K, k


_PY_KERNELS = dict(rollout=_rollout, update_traj=_update_traj, cost=_cost, derivs=_derivs, backward=_backward)
_NUMBA_KERNELS = None


def _numba_kernels():
    global _NUMBA_KERNELS
    if _NUMBA_KERNELS is None:
        from numba import njit
        _NUMBA_KERNELS = {name: njit(fn) for name, fn in _PY_KERNELS.items()}
    return _NUMBA_KERNELS


# ----------------------------------------------------------------------------- stage API

def rollout(prob: Problem, x0=None, U=None):
    x0 = prob.init_state if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1)
    U = prob.init_action if U is None else np.asarray(U, dtype=np.float64).reshape(prob.T, prob.m)
    return prob._k["rollout"](prob.f, x0, U, prob.n, prob.m, prob.constr)


def update_traj(prob: Problem, X, K, k, alpha):
    return prob._k["update_traj"](prob.f, X, K, k.reshape(prob.T, prob.m), float(alpha), prob.n, prob.m, prob.constr)


def cost(prob: Problem, X, P=None):
    return float(prob._k["cost"](prob.l, X, prob.params(P)))


def derivs(prob: Problem, X, P=None):
    return prob._k["derivs"](prob.F, prob.g, prob.H, X, prob.params(P), prob.n)


def backward(prob: Problem, Fm, c, C):
    c3 = np.ascontiguousarray(c).reshape(prob.T, prob.d, 1)
    K, k = prob._k["backward"](Fm, c3, C, prob.n, prob.m)
    return K, k.reshape(prob.T, prob.m)


def line_search(prob: Problem, X, K, k, J_last, P, max_line_search=10, gamma=0.5, method="vanilla"):
    """Return (X_new, J_new, accepted_index or -1, rollouts_done, J_tried[list])."""
    tried = []
    if method == "none":
        Xn = update_traj(prob, X, K, k, 1.0)
        Jn = cost(prob, Xn, P)
        return Xn, Jn, 0, 1, [Jn]
    alpha = 1.0
    for a in range(max_line_search):
        Xn = update_traj(prob, X, K, k, alpha)
        Jn = cost(prob, Xn, P)
        tried.append(Jn)
        alpha = alpha * gamma
        delta = Jn - J_last
        ok = delta < 0
        if method == "feasibility":
            ok = ok and not np.isnan(delta)
        if ok:
            return Xn, Jn, a, a + 1, tried
    return X, J_last, -1, max_line_search, tried


def is_stop(J, J_last, criterion=1e-6, method="relative"):
    with np.errstate(invalid="ignore", divide="ignore"):
        delta = np.float64(J) - np.float64(J_last)
        if method == "vanilla":
            return bool(abs(delta) < criterion)
        return bool(abs(delta / np.float64(J_last)) < criterion)


def solve(prob: Problem, x0=None, target=None, U0=None, max_iter=1000, is_check_stop=True,
          stopping_criterion=1e-6, max_line_search=10, gamma=0.5, line_search_method="vanilla",
          stopping_method="relative", J_last=np.inf):
    """One solve; returns a dict with X, J, iters and the per-iteration trace."""
    P = prob.params(target)
    X = rollout(prob, x0, U0)
    J0 = cost(prob, X, P)
    Fm, c, C = derivs(prob, X, P)
    Js, alphas, rollouts = [], [], 0
    iters, stopped = 0, False
    for _ in range(max_iter):
        K, k = backward(prob, Fm, c, C)
        X, J, a_idx, n_roll, _tried = line_search(prob, X, K, k, J_last, P, max_line_search, gamma,
                                                  line_search_method)
        stop = is_stop(J, J_last, stopping_criterion, stopping_method)
        Fm, c, C = derivs(prob, X, P)
        J_last = J
        Js.append(J)
        alphas.append(a_idx)
        rollouts += n_roll
        iters += 1
        if stop and is_check_stop:
            stopped = True
            break
    return dict(X=X, J=J_last, J0=J0, iters=iters, Js=np.asarray(Js), alphas=np.asarray(alphas, dtype=np.int32),
                rollouts=rollouts, stopped=stopped)


# ----------------------------------------------------------------------------- LogBarrieriLQR

class BarrierScenario:
    """The scenario LogBarrieriLQR.init builds its barrier objective from (log_barrier_ilqr.py:84-101):
    every finite bound adds (-1/t) log(-(lo - x_u[i])) / (-1/t) log(-(x_u[i] - hi)); the barrier parameter
    t becomes the last add_param column, initialised with t0."""

    def __init__(self, scenario, t0):
        xu = scenario.x_u_var
        t_var = sp.symbols("t")
        constr = np.asarray(scenario.constr, dtype=np.float64)
        obj = scenario.obj_fun
        for i, c in enumerate(constr):
            if not np.isinf(c[0]):
                obj = obj + (-1 / t_var) * sp.log(-(c[0] - xu[i]))
            if not np.isinf(c[1]):
                obj = obj + (-1 / t_var) * sp.log(-(xu[i] - c[1]))
        self.name = scenario.name
        self.x_u_var = xu
        self.init_state = scenario.init_state
        self.init_action = scenario.init_action
        self.constr = scenario.constr
        self.dynamic_function = scenario.dynamic_function
        self.obj_fun = obj
        T = int(scenario.init_action.shape[0])
        if scenario.add_param_var is None:                       # log_barrier_ilqr.py:86-88
            self.add_param_var = (t_var,)
        else:
            self.add_param_var = (*scenario.add_param_var, t_var)
        if scenario.add_param is None:                           # :96-98
            self.add_param = t0 * np.ones((T, 1))
        else:
            self.add_param = np.hstack([np.asarray(scenario.add_param, dtype=np.float64), t0 * np.ones((T, 1))])
        self.has_params = scenario.add_param is not None


def solve_log_barrier(prob_barrier: Problem, prob_real: Problem, x0=None, target=None, t_list=(0.5, 1., 2., 5., 10., 20., 50., 100.),
                      max_iter=1000, is_check_stop=True, stopping_criterion=1e-6, max_line_search=10, gamma=0.5,
                      line_search_method="feasibility", stopping_method="relative"):
    """LogBarrieriLQR.solve (log_barrier_ilqr.py:104-149) for a fresh solver object.  `prob_barrier` is the
    Problem of a BarrierScenario, `prob_real` the plain one (reported cost).  Note the order of operations the
    reference has: C, c, F are evaluated inside the forward pass, so the first backward pass after t moves on
    still sees derivatives taken with the previous t."""
    T = prob_barrier.T
    P_real = prob_real.params(target)
    if prob_barrier.add_param.shape[1] == 1:                     # scenario without add_param: the row is (t) alone
        P = t_list[0] * np.ones((T, 1))
    else:
        P = np.hstack([P_real, t_list[0] * np.ones((T, 1))])
    X = rollout(prob_barrier, x0, None)
    Fm, c, C = derivs(prob_barrier, X, P)
    J_last = np.inf
    inner, Js, Jreal, alphas = [], [], [], []
    for j in t_list:
        if j != t_list[0]:
            P = P.copy()
            P[:, -1] = j
        n_it = 0
        for _ in range(max_iter):
            K, k = backward(prob_barrier, Fm, c, C)
            X, J, a_idx, _n, _tried = line_search(prob_barrier, X, K, k, J_last, P, max_line_search, gamma, line_search_method)
            stop = is_stop(J, J_last, stopping_criterion, stopping_method)
            Fm, c, C = derivs(prob_barrier, X, P)
            J_last = J
            Js.append(J)
            Jreal.append(cost(prob_real, X, P_real))
            alphas.append(a_idx)
            n_it += 1
            if stop and is_check_stop:
                J_last = np.inf
                break
        inner.append(n_it)
    return dict(X=X, Jreal=Jreal[-1], inner_iters=np.asarray(inner), Js=np.asarray(Js), Jreals=np.asarray(Jreal),
                alphas=np.asarray(alphas, dtype=np.int32))
