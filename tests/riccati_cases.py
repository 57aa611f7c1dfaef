"""Inputs shared by the CPU (host build) and GPU (C ABI) tests of the pivoted LU and of the teacher-forced Riccati
step (reference: CuriousRL/algorithm/ilqr_solver/ilqr_wrapper.py:240-255)."""
import numpy as np

EPS = np.finfo(np.float64).eps
DENSE_SIZES = [(6, 2), (8, 3), (4, 1), (5, 1), (4, 2)]          # (n, m) the dense recursion is compiled for


def lu_blocks(m, rng):
    """Quu blocks (m x m) that exercise the pivoting: name -> matrix."""
    out = {}
    out["random_nonsymmetric"] = rng.standard_normal((m, m))
    s = rng.standard_normal((m, m))
    s = s + s.T
    out["symmetric_indefinite"] = s - 0.5 * np.eye(m) * np.abs(np.linalg.eigvalsh(s)).max() * 0.1
    out["negative_definite"] = -(s @ s.T + np.eye(m))
    if m >= 2:
        a = rng.standard_normal((m, m))
        a[0, 0] = 1e-3 * a[1, 0]                                  # |a10| >> |a00|: rows 0 and 1 are exchanged
        out["row_exchange"] = a
        a = rng.standard_normal((m, m))
        a[0, 0] = 0.0                                             # exact zero pivot: must exchange
        out["zero_leading_pivot"] = a
        a = rng.standard_normal((m, m))
        a[1, 0] = -a[0, 0]                                        # tied |pivots|: LAPACK (idamax) keeps the first
        out["tied_pivots"] = a
        a = rng.standard_normal((m, m))
        a[1] = a[0] * (1 + 1e-9) + 1e-9 * rng.standard_normal(m)  # nearly dependent rows: tiny second pivot
        out["near_singular"] = a
        u, _, vt = np.linalg.svd(rng.standard_normal((m, m)))
        out["cond_1e10"] = (u * np.logspace(0, -10, m)) @ vt
        out["scaled_1e15"] = rng.standard_normal((m, m)) * np.logspace(15, -3, m)[:, None]   # badly scaled rows
    if m >= 3:
        a = rng.standard_normal((m, m))
        a[:, 0] = [1e-6, 0.5, -2.0]                               # the LAST row holds the first pivot
        a[0, 1] = 5.0                                             # ... and the second elimination step exchanges again
        out["two_exchanges"] = a
    return out


def lu_problem(n, m, A, rng):
    """T = 1 inputs of backward_pass(C, c, F) whose gains are K = -solve(A, B), k = -solve(A, b)."""
    d = n + m
    B = rng.standard_normal((m, n)) * np.logspace(-2, 2, n)[None, :]
    b = rng.standard_normal(m)
    C = rng.standard_normal((d, d))
    C[n:, n:] = A
    C[n:, :n] = B
    c = rng.standard_normal(d)
    c[n:] = b
    F = np.zeros((1, n, d))                                       # V = 0 at the last step: Q = C whatever F is
    return F, c[None], C[None], -np.linalg.solve(A, B), -np.linalg.solve(A, b)


def reference_value_functions(F, c, C, n, m):
    """The reference's recursion (ilqr_wrapper.py:240-255, the oracle's restatement of it) with the value function of
    every step kept: V[t], v[t] are what step t - 1 consumes (V[T] = 0)."""
    T = F.shape[0]
    V = np.zeros((T + 1, n, n))
    v = np.zeros((T + 1, n, 1))
    K = np.zeros((T, m, n))
    k = np.zeros((T, m))
    cond = np.zeros(T)
    K_scale = np.zeros(T)
    k_scale = np.zeros(T)
    lam_min = np.zeros(T)
    swap = np.zeros(T, dtype=bool)
    for t in range(T - 1, -1, -1):
        Q = C[t] + F[t].T @ V[t + 1] @ F[t]
        q = c[t].reshape(-1, 1) + F[t].T @ v[t + 1]
        Quu, Qux, qu = Q[n:, n:].copy(), Q[n:, :n].copy(), q[n:].copy()
        cond[t] = np.linalg.cond(Quu)
        lam_min[t] = np.linalg.eigvalsh(0.5 * (Quu + Quu.T)).min()
        swap[t] = m > 1 and np.abs(Quu[1:, 0]).max() > abs(Quu[0, 0])
        Kt = -np.linalg.solve(Quu, Qux)
        kt = -np.linalg.solve(Quu, qu)
        K[t], k[t] = Kt, kt[:, 0]
        # magnitudes the rounding errors of a step are relative to: the gains themselves, and |Quu^-1| times the sum of
        # the absolute terms that make up Qux / qu (they can cancel: near convergence qu -> 0 while its terms do not)
        inv = np.abs(np.linalg.inv(Quu)).sum(axis=1).max()
        Fa = np.abs(F[t])
        Qux_terms = np.abs(C[t][n:, :n]) + Fa[:, n:].T @ np.abs(V[t + 1]) @ Fa[:, :n]
        qu_terms = np.abs(c[t][n:]).reshape(-1, 1) + Fa[:, n:].T @ np.abs(v[t + 1])
        K_scale[t] = max(np.abs(Kt).max(), inv * Qux_terms.max())
        k_scale[t] = max(np.abs(kt).max(), inv * qu_terms.max())
        V[t] = Q[:n, :n] + Qux.T @ Kt + Kt.T @ Qux + Kt.T @ Quu @ Kt
        v[t] = q[:n] + Qux.T @ kt + Kt.T @ qu + Kt.T @ Quu @ kt
    return dict(V=V, v=v[:, :, 0], K=K, k=k, cond=cond, K_scale=K_scale, k_scale=k_scale, lam_min=lam_min, swap=swap)


def teacher_forced_windows(F, c, C, V, v, L):
    """Dense inputs [B, L + 1, ...] for backward_pass: window b = steps t0 .. t0 + L - 1 of the recursion, followed by
    one synthetic step that hands the reference's value function of step t0 + L to the recursion EXACTLY:
    F = [I 0], C = blockdiag(V, I), c = (v, 0) give Q = C, K = 0, k = 0 and V' = V, v' = v with no rounding."""
    T, n, d = F.shape
    m = d - n
    starts = list(range(0, T - L + 1))
    B = len(starts)
    Fw = np.zeros((B, L + 1, n, d))
    cw = np.zeros((B, L + 1, d))
    Cw = np.zeros((B, L + 1, d, d))
    for b, t0 in enumerate(starts):
        Fw[b, :L], cw[b, :L], Cw[b, :L] = F[t0:t0 + L], c[t0:t0 + L], C[t0:t0 + L]
        Fw[b, L, :, :n] = np.eye(n)
        Cw[b, L, :n, :n] = V[t0 + L]
        Cw[b, L, n:, n:] = np.eye(m)
        cw[b, L, :n] = v[t0 + L]
    return np.asarray(starts), Fw, cw, Cw


def rel_rows(a, b):
    """max-norm relative error per leading index."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    num = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
    den = np.maximum(np.abs(b).reshape(b.shape[0], -1).max(axis=1), 1e-300)
    return num / den
