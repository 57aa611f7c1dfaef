// Per-trajectory numerical core of the batched iLQR path: one lane = one trajectory.
//
// Each function restates one stage of the reference (paths relative to
// /root/reference/CuriousRL/algorithm/ilqr_solver/):
//   rollout_open    ilqr_dynamic_model.py:96-109   (+ cost, ilqr_obj_fun.py:86-95)
//   rollout_closed  ilqr_dynamic_model.py:111-134  (+ cost)
//   backward_fused  ilqr_wrapper.py:232-256 with F, c, C (ilqr_dynamic_model.py:136-147,
//                   ilqr_obj_fun.py:97-119) recomputed from X on the fly, never materialised
//   derivs_dense    the materialised F, c, C of the reference (API / stage-parity only)
//   backward_dense  ilqr_wrapper.py:232-256 on caller-supplied dense F, c, C (API only)
//
// Everything is a __host__ __device__ template over <Model, real, view strides>; T is a
// run-time argument, all n/m-sized loops are unrolled at compile time so that matrices live
// in registers and structural zeros of the Jacobian are pruned by constant folding.
#pragma once

#include "ilqr_models.cuh"

namespace crl {

// Cost parameters (the per-time-step target) of one trajectory: p(t, i) = ptr[t*stride_t + i].
// stride_t = 0 broadcasts one target over the horizon.
template <class R>
struct ParamRef {
    const R* ptr;
    int stride_t;
    CRL_HD R get(int t, int i) const { return ptr[(size_t)t * stride_t + i]; }
};

template <class Mdl, class R>
CRL_HD double step_cost(const R* xu, const ParamRef<R>& prm, int t) {
    double xd[Mdl::D], pd[Mdl::P];
    CRL_UNROLL for (int i = 0; i < Mdl::D; ++i) xd[i] = (double)xu[i];
    CRL_UNROLL for (int i = 0; i < Mdl::P; ++i) pd[i] = (double)prm.get(t, i);
    return Mdl::cost(xd, pd);
}

// Run-time action bounds (lets RoboticArm(is_with_constraints=False) share the device model).
// The clamp is applied unconditionally, also against +-inf, exactly as the reference does
// (min(max(-inf, u), inf)); keeping it branch-free also keeps a time step one basic block.
template <class R>
struct Bounds {
    R lo, hi;          // same box for every action, as in all three scenarios
};

// -----------------------------------------------------------------------------------------
// Open-loop rollout from x0 under the action sequence U (nullptr = zeros).
// Quirk kept: row t is clamped only AFTER it produced row t+1, and row T-1 is never clamped
// (ilqr_dynamic_model.py:104-108).  State bounds are +-inf in all three scenarios.
// Returns J = sum_t l(X[t]) of the stored trajectory.
template <class Mdl, class R, class XV>
CRL_HD double rollout_open_v(const R* x0, int sx0, const R* U, int su_t, const ParamRef<R>& prm,
                             const Bounds<R>& ub, int T, XV X) {
    constexpr int N = Mdl::N, M = Mdl::M, D = Mdl::D;
    R xu[D], xn[N];
    CRL_UNROLL for (int i = 0; i < N; ++i) xu[i] = x0[i * sx0];
    CRL_UNROLL for (int a = 0; a < M; ++a) xu[N + a] = U ? U[a] : R(0);
    double J = 0.0;
    for (int t = 0; t < T; ++t) {
        const bool last = (t == T - 1);
        if (!last) model_step<Mdl, R>(xu, xn);            // from the unclamped row
        if (!last) {
            CRL_UNROLL for (int a = 0; a < M; ++a) xu[N + a] = clamp_(xu[N + a], ub.lo, ub.hi);
        }
        R* row = X.row(t);
        CRL_UNROLL for (int i = 0; i < D; ++i) X.st(row, i, xu[i]);
        J = J + step_cost<Mdl, R>(xu, prm, t);
        if (!last) {
            CRL_UNROLL for (int i = 0; i < N; ++i) xu[i] = xn[i];
            CRL_UNROLL for (int a = 0; a < M; ++a) xu[N + a] = U ? U[(size_t)(t + 1) * su_t + a] : R(0);
        }
    }
    return J;
}

template <class Mdl, class R, int SC>
CRL_HD double rollout_open(const R* x0, int sx0, const R* U, int su_t, const ParamRef<R>& prm,
                           const Bounds<R>& ub, int T, View<R, Mdl::D, SC> X) {
    return rollout_open_v<Mdl, R>(x0, sx0, U, su_t, prm, ub, T, X);
}

// -----------------------------------------------------------------------------------------
// Where a closed-loop rollout gets its per-step inputs (old states/actions, K, k) from.
// DirectFeed loads them from global memory through strided views (API layout, host build,
// cooperative straggler mode).  The solve kernel's TileFeed (ilqr_kernels.cu) streams whole
// time-step blocks of the warp tile into shared memory one step ahead with bulk async copies.
// KN = number of state columns stored per row of K (N for the API layout; NS in the solver's
// workspace, where the output columns are structurally zero and not stored).
template <class Mdl, class R, int KN, int SCX, int SCK>
struct DirectFeed {
    static constexpr int kKN = KN;
    View<R, Mdl::D, SCX> Xold;
    View<R, Mdl::M * KN, SCK> Kv;
    View<R, Mdl::M, SCK> kv;
    CRL_HD void begin(int) {}
    CRL_HD void end() {}
    CRL_HD void row0(R* xu) const {
        const R* row = Xold.row(0);
        CRL_UNROLL for (int i = 0; i < Mdl::D; ++i) xu[i] = Xold.ld(row, i);
    }
    CRL_HD void get(int t, R* xo, R* uo, R* Kc, R* kc) const {
        const R* xr = Xold.row(t);
        const R* Kr = Kv.row(t);
        const R* kr = kv.row(t);
        CRL_UNROLL for (int j = 0; j < KN; ++j) xo[j] = Xold.ld(xr, j);
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) uo[a] = Xold.ld(xr, Mdl::N + a);
        CRL_UNROLL for (int j = 0; j < Mdl::M * KN; ++j) Kc[j] = Kv.ld(Kr, j);
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) kc[a] = kv.ld(kr, a);
    }
};

// Closed-loop rollout  u = u_old + K (x_new - x_old) + alpha k  for one step size
// (ilqr_dynamic_model.py:111-134) with the cost of the new trajectory (ilqr_obj_fun.py:86-95).
template <class Mdl, class R, class Feed, class XV>
CRL_HD double rollout_closed_f(Feed& feed, R alpha, const ParamRef<R>& prm, const Bounds<R>& ub, int T,
                               XV Xnew) {
    constexpr int N = Mdl::N, M = Mdl::M, D = Mdl::D, KN = Feed::kKN;
    R xu[D], xn[N];
    feed.begin(T - 1);
    feed.row0(xu);                                                        // row 0 copied whole (:121)
    double J = 0.0;
    for (int t = 0; t < T - 1; ++t) {
        R xo_c[KN], uo_c[M], K_c[M * KN], k_c[M];
        feed.get(t, xo_c, uo_c, K_c, k_c);
        R dx[KN];
        CRL_UNROLL for (int j = 0; j < KN; ++j) dx[j] = xu[j] - xo_c[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            R acc = K_c[a * KN] * dx[0];
            CRL_UNROLL for (int j = 1; j < KN; ++j) acc = fmadd(K_c[a * KN + j], dx[j], acc);
            R du = acc + alpha * k_c[a];
            R u = uo_c[a] + du;
            u = clamp_(u, ub.lo, ub.hi);
            xu[N + a] = u;
        }
        R* row = Xnew.row(t);
        CRL_UNROLL for (int i = 0; i < D; ++i) Xnew.st(row, i, xu[i]);
        J = J + step_cost<Mdl, R>(xu, prm, t);
        model_step<Mdl, R>(xu, xn);
        CRL_UNROLL for (int i = 0; i < N; ++i) xu[i] = xn[i];
    }
    feed.end();
    if (T > 1) {
        CRL_UNROLL for (int a = 0; a < M; ++a) xu[N + a] = R(0);         // last action stays 0 (:133)
    }
    R* row = Xnew.row(T - 1);
    CRL_UNROLL for (int i = 0; i < D; ++i) Xnew.st(row, i, xu[i]);
    J = J + step_cost<Mdl, R>(xu, prm, T - 1);
    return J;
}

template <class Mdl, class R, int KN, int SCX, int SCK, int SCXN>
CRL_HD double rollout_closed(View<R, Mdl::D, SCX> Xold, View<R, Mdl::M * KN, SCK> Kv, View<R, Mdl::M, SCK> kv,
                             R alpha, const ParamRef<R>& prm, const Bounds<R>& ub, int T,
                             View<R, Mdl::D, SCXN> Xnew) {
    DirectFeed<Mdl, R, KN, SCX, SCK> feed{Xold, Kv, kv};
    return rollout_closed_f<Mdl, R>(feed, alpha, prm, ub, T, Xnew);
}

// -----------------------------------------------------------------------------------------
// NA step sizes in ONE pass over the stored data: the old trajectory, K and k are read once
// and shared by all candidates (this is what makes the line search move its algorithmic bytes
// only), and the NA independent rollouts give each lane NA-way instruction-level parallelism.
// Per candidate the arithmetic is exactly rollout_closed's, so every J[a] is bit-identical to
// a separate single-alpha call.  Only the candidate with index `keep` (per lane) is written
// to Xnew (STORE = false: none); a different winner is re-materialised by one rollout_closed call.
template <class Mdl, class R, int NA, class Feed, int SCXN, bool STORE = true>
CRL_HD void rollout_multi_f(Feed& feed, const R (&alpha)[NA], const ParamRef<R>& prm, const Bounds<R>& ub, int T,
                            int keep, View<R, Mdl::D, SCXN> Xnew, double (&J)[NA]) {
    constexpr int N = Mdl::N, M = Mdl::M, D = Mdl::D, KN = Feed::kKN;
    R xu[NA][D];
    feed.begin(T - 1);
    {
        R x0[D];
        feed.row0(x0);
        CRL_UNROLL for (int i = 0; i < D; ++i) {
            CRL_UNROLL for (int a = 0; a < NA; ++a) xu[a][i] = x0[i];
        }
    }
    CRL_UNROLL for (int a = 0; a < NA; ++a) J[a] = 0.0;
    for (int t = 0; t < T - 1; ++t) {
        R xo_c[KN], uo_c[M], K_c[M * KN], k_c[M];
        feed.get(t, xo_c, uo_c, K_c, k_c);
        CRL_UNROLL for (int c = 0; c < NA; ++c) {
            R dx[KN];
            CRL_UNROLL for (int j = 0; j < KN; ++j) dx[j] = xu[c][j] - xo_c[j];
            CRL_UNROLL for (int a = 0; a < M; ++a) {
                R acc = K_c[a * KN] * dx[0];
                CRL_UNROLL for (int j = 1; j < KN; ++j) acc = fmadd(K_c[a * KN + j], dx[j], acc);
                R du = acc + alpha[c] * k_c[a];
                R u = uo_c[a] + du;
                u = clamp_(u, ub.lo, ub.hi);
                xu[c][N + a] = u;
            }
        }
        if (STORE) {
            R* row = Xnew.row(t);
            CRL_UNROLL for (int i = 0; i < D; ++i) {
                R v = xu[0][i];
                CRL_UNROLL for (int c = 1; c < NA; ++c) v = (keep == c) ? xu[c][i] : v;
                Xnew.st(row, i, v);
            }
        }
        {
            constexpr int NG = Mdl::NANG;
            R ang[NA * NG], sn[NA * NG], cs[NA * NG];
            CRL_UNROLL for (int c = 0; c < NA; ++c) Mdl::template angles<R>(xu[c], ang + c * NG);
            sincos_batch<NA * NG, R>(ang, sn, cs);       // all candidates' trigonometry as one straight-line block
            CRL_UNROLL for (int c = 0; c < NA; ++c) {
                R xn[N];
                J[c] = J[c] + step_cost<Mdl, R>(xu[c], prm, t);
                Mdl::template step_trig<R>(xu[c], sn + c * NG, cs + c * NG, xn);
                CRL_UNROLL for (int i = 0; i < N; ++i) xu[c][i] = xn[i];
            }
        }
    }
    feed.end();
    CRL_UNROLL for (int c = 0; c < NA; ++c) {
        if (T > 1) {
            CRL_UNROLL for (int a = 0; a < M; ++a) xu[c][N + a] = R(0);
        }
        J[c] = J[c] + step_cost<Mdl, R>(xu[c], prm, T - 1);
    }
    if (STORE) {
        R* row = Xnew.row(T - 1);
        CRL_UNROLL for (int i = 0; i < D; ++i) {
            R v = xu[0][i];
            CRL_UNROLL for (int c = 1; c < NA; ++c) v = (keep == c) ? xu[c][i] : v;
            Xnew.st(row, i, v);
        }
    }
}

template <class Mdl, class R, int NA, int KN, int SCX, int SCK, int SCXN, bool STORE = true>
CRL_HD void rollout_multi(View<R, Mdl::D, SCX> Xold, View<R, Mdl::M * KN, SCK> Kv, View<R, Mdl::M, SCK> kv,
                          const R (&alpha)[NA], const ParamRef<R>& prm, const Bounds<R>& ub, int T, int keep,
                          View<R, Mdl::D, SCXN> Xnew, double (&J)[NA]) {
    DirectFeed<Mdl, R, KN, SCX, SCK> feed{Xold, Kv, kv};
    rollout_multi_f<Mdl, R, NA, DirectFeed<Mdl, R, KN, SCX, SCK>, SCXN, STORE>(feed, alpha, prm, ub, T, keep, Xnew, J);
}

// -----------------------------------------------------------------------------------------
// acc (+)= val * f  where f is a Jacobian entry of compile-time kind.
template <class R>
CRL_HD void acc_kind(R& acc, bool& have, R val, int kind, double cst, R var) {
    if (kind == KZ) return;
    if (kind == K1) {
        acc = have ? acc + val : val;
    } else {
        const R f = (kind == KC) ? R(cst) : var;
        acc = have ? fmadd(val, f, acc) : val * f;
    }
    have = true;
}

// Solve  a x = b  (M x M, NR right-hand sides) by LU with partial pivoting, the algorithm
// behind np.linalg.solve / dgesv (ilqr_wrapper.py:246-247): first largest |pivot| per column,
// reciprocal-scaled multipliers, no early-out on tiny or negative pivots.
// Split into the factorisation (depends on `a` only) and its application to right-hand sides, so
// that the warp-cooperative path (ilqr_coop.cuh) can factor once per lane and apply the result to
// one column per lane: per element the arithmetic is the same sequence in both forms.
template <int M, class R>
struct LuFactors {
    R u[M][M];          // rows after elimination (upper triangle is U; the rest is not used)
    R l[M][M];          // multipliers l[i][k], i > k
    R inv[M];           // reciprocal pivots
    bool sw[M][M];      // sw[k][i]: rows k and i were exchanged at step k (i > k)
};

template <int M, class R>
CRL_HD void lu_factor(const R (&a)[M][M], LuFactors<M, R>& f) {
    CRL_UNROLL for (int i = 0; i < M; ++i)
        CRL_UNROLL for (int j = 0; j < M; ++j) f.u[i][j] = a[i][j];
    CRL_UNROLL for (int k = 0; k < M; ++k) {
        int p = k;
        R best = fabs_(f.u[k][k]);
        CRL_UNROLL for (int i = k + 1; i < M; ++i) {
            const R v = fabs_(f.u[i][k]);
            if (v > best) { best = v; p = i; }
        }
        CRL_UNROLL for (int i = k + 1; i < M; ++i) {
            const bool sw = (p == i);
            f.sw[k][i] = sw;
            CRL_UNROLL for (int j = 0; j < M; ++j) {
                const R x = f.u[k][j], y = f.u[i][j];
                f.u[k][j] = sw ? y : x;
                f.u[i][j] = sw ? x : y;
            }
        }
        f.inv[k] = R(1) / f.u[k][k];
        CRL_UNROLL for (int i = k + 1; i < M; ++i) {
            const R l = f.u[i][k] * f.inv[k];
            f.l[i][k] = l;
            CRL_UNROLL for (int j = k + 1; j < M; ++j) f.u[i][j] = fmadd(-l, f.u[k][j], f.u[i][j]);
        }
    }
}

template <int M, int NR, class R>
CRL_HD void lu_apply(const LuFactors<M, R>& f, R (&b)[M][NR]) {
    CRL_UNROLL for (int k = 0; k < M; ++k) {
        CRL_UNROLL for (int i = k + 1; i < M; ++i) {
            const bool sw = f.sw[k][i];
            CRL_UNROLL for (int r = 0; r < NR; ++r) {
                const R x = b[k][r], y = b[i][r];
                b[k][r] = sw ? y : x;
                b[i][r] = sw ? x : y;
            }
        }
        CRL_UNROLL for (int i = k + 1; i < M; ++i) {
            CRL_UNROLL for (int r = 0; r < NR; ++r) b[i][r] = fmadd(-f.l[i][k], b[k][r], b[i][r]);
        }
    }
    CRL_UNROLL for (int j = M - 1; j >= 0; --j) {
        CRL_UNROLL for (int r = 0; r < NR; ++r) b[j][r] = b[j][r] * f.inv[j];
        CRL_UNROLL for (int i = 0; i < j; ++i) {
            CRL_UNROLL for (int r = 0; r < NR; ++r) b[i][r] = fmadd(-f.u[i][j], b[j][r], b[i][r]);
        }
    }
}

template <int M, int NR, class R>
CRL_HD void lu_solve(const R (&a)[M][M], R (&b)[M][NR]) {
    LuFactors<M, R> f;
    lu_factor<M, R>(a, f);
    lu_apply<M, NR, R>(f, b);
}

// Value-function state carried down the horizon.  Because the output entries y never feed
// back (F[:, y] = 0) the exact recursion keeps V block diagonal: V = diag(Vss, diag(vyy)),
// with vyy = C_yy after the first step, and K[:, y] = 0.
// Vss is kept SYMMETRIC (upper triangle, mirrored on read).  The reference never symmetrises
// V, but its asymmetry is pure rounding noise (~1e-16 relative) - well inside the 1e-15
// perturbation under which its own results are reproducible - and
// symmetric storage halves the live registers of the pass.
template <class Mdl, class R>
struct ValueFn {
    R Vss[Mdl::NS][Mdl::NS];      // only [i][j] with i <= j is meaningful
    R vs[Mdl::NS];
    R vyy[Mdl::NO];
    R vy[Mdl::NO];
    CRL_HD R V(int i, int j) const { return i <= j ? Vss[i][j] : Vss[j][i]; }
    CRL_HD void clear() {
        CRL_UNROLL for (int i = 0; i < Mdl::NS; ++i) {
            vs[i] = R(0);
            CRL_UNROLL for (int j = 0; j < Mdl::NS; ++j) Vss[i][j] = R(0);
        }
        CRL_UNROLL for (int y = 0; y < Mdl::NO; ++y) { vyy[y] = R(0); vy[y] = R(0); }
    }
};

// One Riccati step at row `xu` (time t): consumes/produces `vf`, emits K [M][NS] and k [M].
//   Q = C + (F^T V) F,  q = c + F^T v                       (ilqr_wrapper.py:240-241)
//   K = -Quu^-1 Qux,  k = -Quu^-1 qu                         (:246-247)
//   V = Qxx + Qux^T K + K^T Qux + (K^T Quu) K                (:248-251)
//   v = qx  + Qux^T k + K^T qu  + (K^T Quu) k                (:252-255)
// with the products restricted to the structurally non-zero entries (see ilqr_models.cuh)
// and the symmetric blocks (Qxx, V) formed in their upper triangle only.
template <class Mdl, class R>
CRL_HD void riccati_step(const R* xu, const R* p, ValueFn<Mdl, R>& vf, R (&Kout)[Mdl::M][Mdl::NS], R (&kout)[Mdl::M]) {
    constexpr int NS = Mdl::NS, NO = Mdl::NO, N = Mdl::N, M = Mdl::M;
    R A[NS][NS], B[NS][M], G[NO][NS];
    model_jac<Mdl, R>(xu, A, B, G);

    // MA = A^T Vss (NS x NS), MB = B^T Vss (M x NS): sums over l ascending
    R MA[NS][NS], MB[M][NS];
    CRL_UNROLL for (int i = 0; i < NS; ++i)
        CRL_UNROLL for (int k = 0; k < NS; ++k) {
            R acc = R(0); bool have = false;
            CRL_UNROLL for (int l = 0; l < NS; ++l) acc_kind(acc, have, vf.V(l, k), Mdl::akind(l, i), Mdl::aconst(l, i), A[l][i]);
            MA[i][k] = acc;
        }
    CRL_UNROLL for (int a = 0; a < M; ++a)
        CRL_UNROLL for (int k = 0; k < NS; ++k) {
            R acc = R(0); bool have = false;
            CRL_UNROLL for (int l = 0; l < NS; ++l) acc_kind(acc, have, vf.V(l, k), Mdl::bkind(l, a), Mdl::bconst(l, a), B[l][a]);
            MB[a][k] = acc;
        }

    // Q blocks (Qss: upper triangle)
    R Qss[NS][NS], Qus[M][NS + 1], Quu[M][M], qs[NS];     // Qus carries qu as extra column NS
    CRL_UNROLL for (int i = 0; i < NS; ++i)
        CRL_UNROLL for (int j = i; j < NS; ++j) {
            R acc = R(0); bool have = false;
            CRL_UNROLL for (int k = 0; k < NS; ++k) acc_kind(acc, have, MA[i][k], Mdl::akind(k, j), Mdl::aconst(k, j), A[k][j]);
            CRL_UNROLL for (int y = 0; y < NO; ++y)
                if (Mdl::gkind(y, i) != KZ && Mdl::gkind(y, j) != KZ) {
                    const R mg = G[y][i] * vf.vyy[y];
                    acc = have ? fmadd(mg, G[y][j], acc) : mg * G[y][j];
                    have = true;
                }
            Qss[i][j] = (i == j) ? R(cost_hess<Mdl>(i)) + acc : acc;
        }
    CRL_UNROLL for (int a = 0; a < M; ++a) {
        CRL_UNROLL for (int j = 0; j < NS; ++j) {
            R acc = R(0); bool have = false;
            CRL_UNROLL for (int k = 0; k < NS; ++k) acc_kind(acc, have, MB[a][k], Mdl::akind(k, j), Mdl::aconst(k, j), A[k][j]);
            Qus[a][j] = acc;
        }
        CRL_UNROLL for (int b = 0; b < M; ++b) {
            R acc = R(0); bool have = false;
            CRL_UNROLL for (int k = 0; k < NS; ++k) acc_kind(acc, have, MB[a][k], Mdl::bkind(k, b), Mdl::bconst(k, b), B[k][b]);
            Quu[a][b] = (a == b) ? cost_hess_u<Mdl, R>(a, xu[N + a], p) + acc : acc;
        }
    }
    // q = c + F^T v
    CRL_UNROLL for (int i = 0; i < NS; ++i) {
        R acc = R(0); bool have = false;
        CRL_UNROLL for (int l = 0; l < NS; ++l) acc_kind(acc, have, vf.vs[l], Mdl::akind(l, i), Mdl::aconst(l, i), A[l][i]);
        CRL_UNROLL for (int y = 0; y < NO; ++y)
            if (Mdl::gkind(y, i) != KZ) { acc = have ? fmadd(vf.vy[y], G[y][i], acc) : vf.vy[y] * G[y][i]; have = true; }
        qs[i] = cost_grad<Mdl, R>(i, xu[i], p) + acc;
    }
    R qu[M];
    CRL_UNROLL for (int a = 0; a < M; ++a) {
        R acc = R(0); bool have = false;
        CRL_UNROLL for (int l = 0; l < NS; ++l) acc_kind(acc, have, vf.vs[l], Mdl::bkind(l, a), Mdl::bconst(l, a), B[l][a]);
        qu[a] = cost_grad<Mdl, R>(N + a, xu[N + a], p) + acc;
        Qus[a][NS] = qu[a];
    }

    // gains: solve Quu [K | k] = [Qus | qu], then negate
    R LU[M][M], X[M][NS + 1];
    CRL_UNROLL for (int a = 0; a < M; ++a) {
        CRL_UNROLL for (int b = 0; b < M; ++b) LU[a][b] = Quu[a][b];
        CRL_UNROLL for (int j = 0; j <= NS; ++j) X[a][j] = Qus[a][j];
    }
    lu_solve<M, NS + 1, R>(LU, X);
    CRL_UNROLL for (int a = 0; a < M; ++a) {
        CRL_UNROLL for (int j = 0; j < NS; ++j) Kout[a][j] = -X[a][j];
        kout[a] = -X[a][NS];
    }

    // W = K^T Quu (NS x M)
    R W[NS][M];
    CRL_UNROLL for (int i = 0; i < NS; ++i)
        CRL_UNROLL for (int b = 0; b < M; ++b) {
            R acc = Kout[0][i] * Quu[0][b];
            CRL_UNROLL for (int a = 1; a < M; ++a) acc = fmadd(Kout[a][i], Quu[a][b], acc);
            W[i][b] = acc;
        }
    // v = ((qs + Qus^T k) + K^T qu) + W k
    CRL_UNROLL for (int i = 0; i < NS; ++i) {
        R t1 = Qus[0][i] * kout[0], t2 = Kout[0][i] * qu[0], t3 = W[i][0] * kout[0];
        CRL_UNROLL for (int a = 1; a < M; ++a) {
            t1 = fmadd(Qus[a][i], kout[a], t1);
            t2 = fmadd(Kout[a][i], qu[a], t2);
            t3 = fmadd(W[i][a], kout[a], t3);
        }
        vf.vs[i] = ((qs[i] + t1) + t2) + t3;
    }
    // V = ((Qss + S) + S^T) + W K with S = Qus^T K, upper triangle only
    CRL_UNROLL for (int i = 0; i < NS; ++i)
        CRL_UNROLL for (int j = i; j < NS; ++j) {
            R sij = Qus[0][i] * Kout[0][j], sji = Qus[0][j] * Kout[0][i], r = W[i][0] * Kout[0][j];
            CRL_UNROLL for (int a = 1; a < M; ++a) {
                sij = fmadd(Qus[a][i], Kout[a][j], sij);
                sji = fmadd(Qus[a][j], Kout[a][i], sji);
                r = fmadd(W[i][a], Kout[a][j], r);
            }
            vf.Vss[i][j] = ((Qss[i][j] + sij) + sji) + r;
        }
    // output block: Vyy <- Cyy, vy <- c_y(x_t)
    CRL_UNROLL for (int y = 0; y < NO; ++y) {
        vf.vyy[y] = R(cost_hess<Mdl>(NS + y));
        vf.vy[y] = cost_grad<Mdl, R>(NS + y, xu[NS + y], p);
    }
}

// Backward pass over a stored trajectory; F, c, C are recomputed per step.
// K is written with KN state columns per row (KN = N: output columns get -0.0, as the
// reference's -solve(Quu, 0) does; KN = NS: compact workspace form).
template <class Mdl, class R, int SCX>
struct DirectRowFeed {                       // rows of a stored trajectory, last to first
    View<R, Mdl::D, SCX> X;
    int T;
    CRL_HD void begin(int) {}
    CRL_HD void end() {}
    CRL_HD void get(int i, R* xu) const {    // i-th row handed out = time step T-1-i
        const R* row = X.row(T - 1 - i);
        CRL_UNROLL for (int c = 0; c < Mdl::D; ++c) xu[c] = X.ld(row, c);
    }
};

// `first` (barrier models only): this is the first backward pass of a solve - its cost derivatives use
// the stale barrier parameter p[kT + 1] (see Barrier in ilqr_models.cuh).
template <class Mdl, class R, int KN, class RowFeed, int SCK>
CRL_HD void backward_fused_f(RowFeed& rows, const ParamRef<R>& prm, int T, View<R, Mdl::M * KN, SCK> Kv,
                             View<R, Mdl::M, SCK> kv, bool first = false) {
    constexpr int NS = Mdl::NS, M = Mdl::M, D = Mdl::D, P = Mdl::P;
    ValueFn<Mdl, R> vf;
    vf.clear();
    rows.begin(T);
    for (int t = T - 1; t >= 0; --t) {
        R xu[D], p[P], Kt[M][NS], kt[M];
        rows.get(T - 1 - t, xu);
        CRL_UNROLL for (int i = 0; i < P; ++i) p[i] = prm.get(t, i);
        if constexpr (Mdl::kBarrier) {
            if (first) p[Mdl::kT] = p[Mdl::kT + 1];
        }
        riccati_step<Mdl, R>(xu, p, vf, Kt, kt);
        R* Kr = Kv.row(t);
        R* kr = kv.row(t);
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            CRL_UNROLL for (int j = 0; j < KN; ++j) Kv.st(Kr, a * KN + j, j < NS ? Kt[a][j] : R(-0.0));
            kv.st(kr, a, kt[a]);
        }
    }
    rows.end();
}

template <class Mdl, class R, int KN, int SCX, int SCK>
CRL_HD void backward_fused(View<R, Mdl::D, SCX> X, const ParamRef<R>& prm, int T,
                           View<R, Mdl::M * KN, SCK> Kv, View<R, Mdl::M, SCK> kv) {
    DirectRowFeed<Mdl, R, SCX> rows{X, T};
    backward_fused_f<Mdl, R, KN>(rows, prm, T, Kv, kv);
}

// Materialised derivatives of one row (API / stage parity): F [N][D], c [D], C [D][D] row-major.
template <class Mdl, class R>
CRL_HD void derivs_dense(const R* xu, const R* p, R* F, R* c, R* C) {
    constexpr int NS = Mdl::NS, NO = Mdl::NO, N = Mdl::N, M = Mdl::M, D = Mdl::D;
    R A[NS][NS], B[NS][M], G[NO][NS];
    model_jac<Mdl, R>(xu, A, B, G);
    for (int i = 0; i < N * D; ++i) F[i] = R(0);
    CRL_UNROLL for (int i = 0; i < NS; ++i) {
        CRL_UNROLL for (int j = 0; j < NS; ++j) {
            const int kd = Mdl::akind(i, j);
            F[i * D + j] = kd == KZ ? R(0) : (kd == K1 ? R(1) : (kd == KC ? R(Mdl::aconst(i, j)) : A[i][j]));
        }
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            const int kd = Mdl::bkind(i, a);
            F[i * D + N + a] = kd == KZ ? R(0) : (kd == K1 ? R(1) : (kd == KC ? R(Mdl::bconst(i, a)) : B[i][a]));
        }
    }
    CRL_UNROLL for (int y = 0; y < NO; ++y)
        CRL_UNROLL for (int j = 0; j < NS; ++j)
            F[(NS + y) * D + j] = Mdl::gkind(y, j) == KZ ? R(0) : G[y][j];
    for (int i = 0; i < D * D; ++i) C[i] = R(0);
    CRL_UNROLL for (int i = 0; i < D; ++i) {
        c[i] = cost_grad<Mdl, R>(i, xu[i], p);
        C[i * D + i] = i >= N ? cost_hess_u<Mdl, R>(i - N, xu[i], p) : R(cost_hess<Mdl>(i));
    }
}

// One step of the dense Riccati recursion (ilqr_wrapper.py:240-255) on row-major Ft [N][D], ct [D], Ct [D][D]:
// updates V, v in place and writes the gains Kt [M][N], kt [M].  Products accumulate in the order of the
// reference's BLAS calls (first term a product, the rest FMAs).
template <int N, int M, class R>
CRL_HD void dense_riccati_step(const R* Ft, const R* ct, const R* Ct, R (&V)[N][N], R (&v)[N], R* Kt, R* kt) {
    constexpr int D = N + M;
    R Mx[D][N], Q[D][D], q[D];
    for (int i = 0; i < D; ++i)
        for (int kk = 0; kk < N; ++kk) {
            R acc = Ft[0 * D + i] * V[0][kk];
            for (int l = 1; l < N; ++l) acc = fmadd(Ft[l * D + i], V[l][kk], acc);
            Mx[i][kk] = acc;
        }
    for (int i = 0; i < D; ++i) {
        for (int j = 0; j < D; ++j) {
            R acc = Mx[i][0] * Ft[0 * D + j];
            for (int kk = 1; kk < N; ++kk) acc = fmadd(Mx[i][kk], Ft[kk * D + j], acc);
            Q[i][j] = Ct[i * D + j] + acc;
        }
        R acc = Ft[0 * D + i] * v[0];
        for (int l = 1; l < N; ++l) acc = fmadd(Ft[l * D + i], v[l], acc);
        q[i] = ct[i] + acc;
    }
    R LU[M][M], X[M][N + 1];
    for (int a = 0; a < M; ++a) {
        for (int b = 0; b < M; ++b) LU[a][b] = Q[N + a][N + b];
        for (int j = 0; j < N; ++j) X[a][j] = Q[N + a][j];
        X[a][N] = q[N + a];
    }
    lu_solve<M, N + 1, R>(LU, X);
    for (int a = 0; a < M; ++a) {
        for (int j = 0; j < N; ++j) Kt[a * N + j] = -X[a][j];
        kt[a] = -X[a][N];
    }
    R W[N][M];
    for (int i = 0; i < N; ++i)
        for (int b = 0; b < M; ++b) {
            R acc = Kt[0 * N + i] * Q[N + 0][N + b];
            for (int a = 1; a < M; ++a) acc = fmadd(Kt[a * N + i], Q[N + a][N + b], acc);
            W[i][b] = acc;
        }
    R Vn[N][N], vn[N];
    for (int i = 0; i < N; ++i) {
        R t1 = Q[N][i] * kt[0], t2 = Kt[i] * q[N], t3 = W[i][0] * kt[0];
        for (int a = 1; a < M; ++a) {
            t1 = fmadd(Q[N + a][i], kt[a], t1);
            t2 = fmadd(Kt[a * N + i], q[N + a], t2);
            t3 = fmadd(W[i][a], kt[a], t3);
        }
        vn[i] = ((q[i] + t1) + t2) + t3;
        for (int j = 0; j < N; ++j) {
            R s1 = Q[N][i] * Kt[j], s2 = Kt[i] * Q[N][j], r = W[i][0] * Kt[j];
            for (int a = 1; a < M; ++a) {
                s1 = fmadd(Q[N + a][i], Kt[a * N + j], s1);
                s2 = fmadd(Kt[a * N + i], Q[N + a][j], s2);
                r = fmadd(W[i][a], Kt[a * N + j], r);
            }
            Vn[i][j] = ((Q[i][j] + s1) + s2) + r;
        }
    }
    for (int i = 0; i < N; ++i) { v[i] = vn[i]; for (int j = 0; j < N; ++j) V[i][j] = Vn[i][j]; }
}

// Dense Riccati recursion on caller-supplied F [T][N][D], c [T][D], C [T][D][D] (row-major,
// one trajectory).  The reference's backward_pass(C, c, F) signature; not on the hot path.
template <int N, int M, class R>
CRL_HD void backward_dense(const R* F, const R* c, const R* C, int T, R* K, R* k) {
    constexpr int D = N + M;
    R V[N][N], v[N];
    for (int i = 0; i < N; ++i) { v[i] = R(0); for (int j = 0; j < N; ++j) V[i][j] = R(0); }
    for (int t = T - 1; t >= 0; --t)
        dense_riccati_step<N, M, R>(F + (size_t)t * N * D, c + (size_t)t * D, C + (size_t)t * D * D, V, v,
                                    K + (size_t)t * M * N, k + (size_t)t * M);
}

// Accept / stop logic of one iteration (ilqr_wrapper.py:98-100, 148-174, 204-213).
// `J_try` is the cost of the candidate for the current alpha; returns true if it is taken.
CRL_HD bool ls_accept(double J_try, double J_last, bool feasibility) {
    const double delta = J_try - J_last;
    bool ok = delta < 0.0;
    if (feasibility) ok = ok && !(delta != delta);
    return ok;
}
CRL_HD bool stop_test(double J_new, double J_last, double criterion, bool relative) {
    const double delta = J_new - J_last;
    const double m = relative ? ::fabs(delta / J_last) : ::fabs(delta);
    return m < criterion;
}

}  // namespace crl
