// Batched iLQR for GENERATED device models (cart-pole, vehicle, car-parking scenarios).
//
// The hand-written arm models (ilqr_core.cuh) exploit their structure: constant diagonal cost Hessian, output
// states, one action box, state bounds at +-inf.  The generated models (ilqr_generated.cuh, written by
// curiousrl_b200/codegen.py) have none of it, so this is the reference's iteration in its general form
// (citations relative to /root/reference/CuriousRL/algorithm/ilqr_solver/):
//
//   gen_rollout_open    ilqr_dynamic_model.py:96-109   every component of row t clamped AFTER it produced row t+1
//   gen_backward        ilqr_wrapper.py:232-256 with F, c, C (ilqr_dynamic_model.py:136-147, ilqr_obj_fun.py:97-119)
//                       recomputed per time step from X and the time-varying parameter row, dense Riccati step
//   gen_rollout_closed  ilqr_dynamic_model.py:111-134  action clamp BEFORE the dynamics, state clamp AFTER them
//   gen_solve_one       basic_ilqr.py:68-97 + ilqr_wrapper.py:74-214 for one trajectory
//
// One thread = one trajectory; FP64 only; all views are strided (SC = 32: [t][component][lane] warp tiles in the
// solver workspace, SC = 1: dense rows in the API layout and in the test-only host build).
#pragma once

#include "ilqr_core.cuh"
#ifdef CRL_USER_MODEL
// run-time compiled model (curiousrl_b200/jit.py): the generated header of ONE scenario, structs GenUser / GenUserBarrier
#include CRL_USER_HEADER
#else
#include "ilqr_generated.cuh"
#endif

#ifndef CRL_GEN_MAX_D
#define CRL_GEN_MAX_D 8          // components (n + m) a GenBox holds; the run-time compiled unit raises it to 16
#endif

namespace crl {

// The box `constr` of a problem, component by component (states first).  The generated models carry the scenario's
// default box (G::lo / G::hi); a caller may pass another one (e.g. is_with_constraints=False: all +-inf).
struct GenBox {
    double lo[CRL_GEN_MAX_D], hi[CRL_GEN_MAX_D];
};
template <class G>
CRL_HD GenBox gen_default_box() {
    static_assert(G::D <= CRL_GEN_MAX_D, "GenBox holds up to CRL_GEN_MAX_D components");
    GenBox b;
    for (int i = 0; i < CRL_GEN_MAX_D; ++i) { b.lo[i] = i < G::D ? G::lo(i) : 0.0; b.hi[i] = i < G::D ? G::hi(i) : 0.0; }
    return b;
}

// min(max(lo, v), hi) with Python's semantics (ilqr_dynamic_model.py:107-108, :127-128, :130-131): max(a, b) is
// "b if b > a else a", min(a, b) is "b if b < a else a" - so a NaN is replaced by the lower bound, as in the reference.
CRL_HD double gen_clamp(double v, int i, const GenBox& box) {
    const double lo = box.lo[i], hi = box.hi[i];
    const double w = v > lo ? v : lo;
    return hi < w ? hi : w;
}

// Parameter row of time step t: p(t, i) = ptr[t * stride_t + i]; the problem's box rides along.
struct GenParams {
    const double* ptr;
    int stride_t;
    GenBox box;
    template <int P>
    CRL_HD void load(int t, double (&p)[P]) const {
        for (int i = 0; i < P; ++i) p[i] = ptr[(size_t)t * stride_t + i];
    }
};

// Open-loop rollout from x0 under U (nullptr = zeros; element (t, a) at U[t * su_t + a]); returns J.
template <class G, class XV>
CRL_HD double gen_rollout_open(const double* x0, const double* U, int su_t, GenParams prm, int T, XV X) {
    constexpr int N = G::N, M = G::M, D = G::D;
    double xu[D], xn[N], p[G::P];
    for (int i = 0; i < N; ++i) xu[i] = x0[i];
    for (int a = 0; a < M; ++a) xu[N + a] = U ? U[a] : 0.0;
    double J = 0.0;
    for (int t = 0; t < T; ++t) {
        const bool last = (t == T - 1);
        if (!last) {
            G::step(xu, xn);                                                   // from the UNclamped row
            for (int i = 0; i < D; ++i) xu[i] = gen_clamp(xu[i], i, prm.box);   // row t clamped afterwards
        }
        double* row = X.row(t);
        for (int i = 0; i < D; ++i) X.st(row, i, xu[i]);
        prm.load(t, p);
        J = J + G::cost(xu, p);
        if (!last) {
            for (int i = 0; i < N; ++i) xu[i] = xn[i];
            for (int a = 0; a < M; ++a) xu[N + a] = U ? U[(size_t)(t + 1) * su_t + a] : 0.0;
        }
    }
    return J;
}

// J of a stored trajectory (ilqr_obj_fun.py:86-95)
template <class G, class XV>
CRL_HD double gen_cost(XV X, GenParams prm, int T) {
    double xu[G::D], p[G::P], J = 0.0;
    for (int t = 0; t < T; ++t) {
        const double* row = X.row(t);
        for (int i = 0; i < G::D; ++i) xu[i] = X.ld(row, i);
        prm.load(t, p);
        J = J + G::cost(xu, p);
    }
    return J;
}

// F, c, C of one row (the reference's materialised derivatives)
template <class G>
CRL_HD void gen_derivs(const double* xu, const double* p, double* F, double* c, double* C) {
    G::jac(xu, F);
    G::grad(xu, p, c);
    G::hess(xu, p, C);
}

// Backward pass: K [T][M*N], k [T][M] from X.
// `first`: the first backward pass of an inner loop of LogBarrieriLQR still sees cost derivatives taken with the
// previous barrier parameter (the reference evaluates C, c in the forward pass, before its outer loop moves t on,
// log_barrier_ilqr.py:121-133): barrier models take t from the stale entry p[kT + 1] then.
template <class G, class XV, class KV, class SV>
CRL_HD void gen_backward(XV X, GenParams prm, int T, KV K, SV k, bool first = false) {
    constexpr int N = G::N, M = G::M, D = G::D;
    double V[N][N], v[N];
    for (int i = 0; i < N; ++i) { v[i] = 0.0; for (int j = 0; j < N; ++j) V[i][j] = 0.0; }
    double xnext[D];
    {
        const double* row = X.row(T - 1);
        for (int i = 0; i < D; ++i) xnext[i] = X.ld(row, i);
    }
    for (int t = T - 1; t >= 0; --t) {
        double xu[D], p[G::P], F[N * D], c[D], C[D * D], Kt[M * N], kt[M];
        for (int i = 0; i < D; ++i) xu[i] = xnext[i];
        if (t > 0) {                                       // the row of step t-1 is loaded while step t computes
            const double* row = X.row(t - 1);
            for (int i = 0; i < D; ++i) xnext[i] = X.ld(row, i);
        }
        prm.load(t, p);
        if constexpr (G::kBarrier) {
            if (first) p[G::kT] = p[G::kT + 1];
        }
        gen_derivs<G>(xu, p, F, c, C);
        dense_riccati_step<N, M, double>(F, c, C, V, v, Kt, kt);
        double* Kr = K.row(t);
        double* kr = k.row(t);
        for (int i = 0; i < M * N; ++i) K.st(Kr, i, Kt[i]);
        for (int a = 0; a < M; ++a) k.st(kr, a, kt[a]);
    }
}

// Closed-loop rollout for one step size; writes Xn, returns its cost.
// `store = false`: cost only (the line-search helpers of gen_solve_warp); same arithmetic, nothing written.
template <class G, class XV, class KV, class SV, class XW>
CRL_HD double gen_rollout_closed(XV X, KV K, SV k, double alpha, GenParams prm, int T, XW Xn, bool store = true) {
    constexpr int N = G::N, M = G::M, D = G::D;
    double xu[D], xo[D], xn[N], p[G::P];
    {
        const double* r0 = X.row(0);
        for (int i = 0; i < D; ++i) xu[i] = X.ld(r0, i);          // row 0 copied, its action overwritten below
    }
    double J = 0.0;
    // old row, K and k of the next step are loaded while the current step computes (they do not depend on it)
    double xo_n[D], K_n[M * N], k_n[M], Kc[M * N], kc[M];
    auto fetch = [&](int t) {
        const double* ro = X.row(t);
        for (int i = 0; i < D; ++i) xo_n[i] = X.ld(ro, i);
        const double* Kr = K.row(t);
        for (int i = 0; i < M * N; ++i) K_n[i] = K.ld(Kr, i);
        const double* kr = k.row(t);
        for (int a = 0; a < M; ++a) k_n[a] = k.ld(kr, a);
    };
    if (T > 1) fetch(0);
    for (int t = 0; t < T; ++t) {
        const bool last = (t == T - 1);
        if (!last) {
            for (int i = 0; i < D; ++i) xo[i] = xo_n[i];
            for (int i = 0; i < M * N; ++i) Kc[i] = K_n[i];
            for (int a = 0; a < M; ++a) kc[a] = k_n[a];
            if (t + 1 < T - 1) fetch(t + 1);
            double dx[N];
            for (int j = 0; j < N; ++j) dx[j] = xu[j] - xo[j];
            for (int a = 0; a < M; ++a) {
                // delta_u = K dx + alpha k;  u = u_old + delta_u  (ilqr_dynamic_model.py:123-126)
                double acc = Kc[a * N + 0] * dx[0];
                for (int j = 1; j < N; ++j) acc = fmadd(Kc[a * N + j], dx[j], acc);
                const double du = acc + alpha * kc[a];
                xu[N + a] = gen_clamp(xo[N + a] + du, N + a, prm.box);
            }
            G::step(xu, xn);
        } else {
            if (T > 1) { for (int a = 0; a < M; ++a) xu[N + a] = 0.0; }        // last action stays 0 (:133)
        }
        if (store) {
            double* rw = Xn.row(t);
            for (int i = 0; i < D; ++i) Xn.st(rw, i, xu[i]);
        }
        prm.load(t, p);
        J = J + G::cost(xu, p);
        if (!last) {
            for (int i = 0; i < N; ++i) xu[i] = gen_clamp(xn[i], i, prm.box);   // state clamp after the dynamics
        }
    }
    return J;
}

struct GenOpts {
    int max_iter, check_stop, max_ls, ls_method, stop_method;
    double criterion, gamma;
};

// The whole solve of one trajectory.  Xa / Xb: two trajectory buffers (the result ends in *final_in_a ? Xa : Xb).
// n_stage > 1 (barrier models): LogBarrieriLQR's outer loop (log_barrier_ilqr.py:121-146) - one inner loop per
// parameter set `prm.ptr + s * stage_stride`, the trajectory is kept, the iteration counter restarts and the stored
// cost is reset to +inf after a loop that ended by the stopping criterion; stage_iters[s] = iterations of loop s.
template <class G, class XV, class KV, class SV>
CRL_HD void gen_solve_one(const double* x0, const double* U, int su_t, GenParams prm, int T, const GenOpts& op,
                          double J_init, XV Xa, XV Xb, KV K, SV k, double* J_out, int* iters_out, int* conv_out,
                          bool* final_in_a, double* J_trace, signed char* alpha_trace, int n_stage = 1,
                          long long stage_stride = 0, int* stage_iters = nullptr) {
    gen_rollout_open<G>(x0, U, su_t, prm, T, Xa);
    bool in_a = true;
    double J_last = J_init;
    const int tries = op.ls_method == 2 /* CRL_LS_NONE */ ? 1 : op.max_ls;
    int total = 0;
    bool stopped = false;
    for (int s = 0; s < n_stage; ++s) {
        GenParams ps = prm;
        ps.ptr = prm.ptr + (size_t)s * stage_stride;
        if (s > 0 && stopped) J_last = (double)INFINITY;                       // log_barrier_ilqr.py:141
        int it = 0;
        stopped = false;
        while (it < op.max_iter) {
            XV Xc = in_a ? Xa : Xb, Xn = in_a ? Xb : Xa;
            gen_backward<G>(Xc, ps, T, K, k, it == 0);
            bool accepted = false;
            double J_new = J_last, alpha = 1.0;
            int taken = -1;
            for (int a = 0; a < tries && !accepted; ++a) {
                const double Jt = gen_rollout_closed<G>(Xc, K, k, alpha, ps, T, Xn);
                if (op.ls_method == 2 || ls_accept(Jt, J_last, op.ls_method == 1)) {
                    accepted = true;
                    J_new = Jt;
                    taken = a;
                }
                alpha = alpha * op.gamma;
            }
            const bool stop = stop_test(J_new, J_last, op.criterion, op.stop_method == 0);
            J_last = J_new;
            if (accepted) in_a = !in_a;
            if (J_trace) J_trace[it] = J_new;
            if (alpha_trace) alpha_trace[it] = (signed char)taken;
            ++it;
            if (stop && op.check_stop) { stopped = true; break; }
        }
        if (stage_iters) stage_iters[s] = it;
        total += it;
    }
    *J_out = J_last;
    *iters_out = total;
    *conv_out = stopped ? 1 : 0;
    *final_in_a = in_a;
}

}  // namespace crl
