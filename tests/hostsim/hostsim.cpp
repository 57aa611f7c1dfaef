// TEST-ONLY host build of the kernels' numerical core (curiousrl_b200/csrc/ilqr_core.cuh).
//
// The product ships no CPU path: libcuriousrl_b200.so contains only CUDA kernels.  This file
// compiles the very same __host__ __device__ templates with g++ so that the `-m "not gpu"`
// tests can check the device math (structured Riccati step, analytic Jacobians, rollouts,
// accept/stop logic) against the oracle in a container without a GPU.  Built on demand by
// tests/hostsim/__init__.py into tests/hostsim/_hostsim.so.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../curiousrl_b200/csrc/ilqr_core.cuh"

using namespace crl;

namespace {

template <class Mdl, class R>
struct Sim {
    static Bounds<R> bounds(double lo, double hi) {
        Bounds<R> b;
        b.lo = (R)lo; b.hi = (R)hi;
        return b;
    }
    static double rollout(const R* x0, const R* u0, const R* params, int stride_t, double lo, double hi, int T, R* X) {
        View<R, Mdl::D, 1> Xv{X};
        return rollout_open<Mdl, R, 1>(x0, 1, u0, Mdl::M, ParamRef<R>{params, stride_t}, bounds(lo, hi), T, Xv);
    }
    static void backward(const R* X, const R* params, int stride_t, int T, R* K, R* k) {
        View<R, Mdl::D, 1> Xv{const_cast<R*>(X)};
        View<R, Mdl::M * Mdl::N, 1> Kv{K};
        View<R, Mdl::M, 1> kv{k};
        backward_fused<Mdl, R, Mdl::N, 1, 1>(Xv, ParamRef<R>{params, stride_t}, T, Kv, kv);
    }
    static double update(const R* X, const R* K, const R* k, double alpha, const R* params, int stride_t, double lo,
                         double hi, int T, R* Xn) {
        View<R, Mdl::D, 1> Xo{const_cast<R*>(X)}, Xw{Xn};
        View<R, Mdl::M * Mdl::N, 1> Kv{const_cast<R*>(K)};
        View<R, Mdl::M, 1> kv{const_cast<R*>(k)};
        return rollout_closed<Mdl, R, Mdl::N, 1, 1, 1>(Xo, Kv, kv, (R)alpha, ParamRef<R>{params, stride_t},
                                                       bounds(lo, hi), T, Xw);
    }
    static void derivs(const R* X, const R* params, int stride_t, int T, R* F, R* c, R* C) {
        constexpr int N = Mdl::N, D = Mdl::D;
        for (int t = 0; t < T; ++t) {
            R p[Mdl::P];
            for (int i = 0; i < Mdl::P; ++i) p[i] = params[(size_t)t * stride_t + i];
            derivs_dense<Mdl, R>(X + (size_t)t * D, p, F + (size_t)t * N * D, c + (size_t)t * D, C + (size_t)t * D * D);
        }
    }
    static void riccati(const R* xu, const R* p, R* V, R* v, R* K, R* k) {
        constexpr int N = Mdl::N, NS = Mdl::NS, NO = Mdl::NO, M = Mdl::M;
        ValueFn<Mdl, R> vf;
        vf.clear();
        for (int i = 0; i < NS; ++i) {
            vf.vs[i] = v[i];
            for (int j = i; j < NS; ++j) vf.Vss[i][j] = V[i * N + j];
        }
        for (int y = 0; y < NO; ++y) { vf.vyy[y] = V[(NS + y) * N + NS + y]; vf.vy[y] = v[NS + y]; }
        R Kt[M][NS], kt[M];
        riccati_step<Mdl, R>(xu, p, vf, Kt, kt);
        for (int a = 0; a < M; ++a) {
            for (int j = 0; j < N; ++j) K[a * N + j] = j < NS ? Kt[a][j] : R(-0.0);
            k[a] = kt[a];
        }
        for (int i = 0; i < N * N; ++i) V[i] = R(0);
        for (int i = 0; i < NS; ++i) {
            v[i] = vf.vs[i];
            for (int j = 0; j < NS; ++j) V[i * N + j] = vf.V(i, j);
        }
        for (int y = 0; y < NO; ++y) { V[(NS + y) * N + NS + y] = vf.vyy[y]; v[NS + y] = vf.vy[y]; }
    }
    // One lane of solve_kernel, run sequentially: same stage functions, same compact K layout,
    // same accept / stop / carry logic.
    static void solve(const R* x0, const R* u0, const R* params, int stride_t, double lo, double hi, int T,
                      int max_iter, int check_stop, double criterion, int max_ls, double gamma, int ls_method,
                      int stop_method, double J_init, R* X_out, double* J_out, int* iters_out, int* conv_out,
                      double* J_trace, int* alpha_trace) {
        constexpr int D = Mdl::D, M = Mdl::M, NS = Mdl::NS;
        std::vector<R> xa((size_t)T * D), xb((size_t)T * D), Kb((size_t)T * M * NS), kb((size_t)T * M);
        ParamRef<R> prm{params, stride_t};
        Bounds<R> ub = bounds(lo, hi);
        int cur = 0;
        View<R, D, 1> X0{xa.data()}, X1{xb.data()};
        View<R, M * NS, 1> Kv{Kb.data()};
        View<R, M, 1> kv{kb.data()};
        rollout_open<Mdl, R, 1>(x0, 1, u0, M, prm, ub, T, X0);
        double J_last = J_init;
        int it = 0, pred = 0;
        while (true) {
            View<R, D, 1> Xc = cur ? X1 : X0, Xn = cur ? X0 : X1;
            backward_fused<Mdl, R, NS, 1, 1>(Xc, prm, T, Kv, kv);
            bool accepted = false, replay = false;
            double J_new = J_last, alpha = 1.0, alpha_taken = 1.0;
            int taken = -1;
            const int tries = (ls_method == 2) ? 1 : max_ls;
            constexpr int CH = 4;                       // = kLsChunk of ilqr_kernels.cu
            int a0 = 0;
            while (a0 < tries && !accepted) {
                const bool wide = (tries - a0 >= 2) && (a0 > 0 || pred > 0);
                if (wide) {
                    R al[CH];
                    double ald[CH], Jc[CH];
                    for (int c = 0; c < CH; ++c) { ald[c] = alpha; al[c] = (R)alpha; alpha = alpha * gamma; }
                    int keep = pred - a0;
                    keep = keep < 0 ? 0 : (keep > CH - 1 ? CH - 1 : keep);
                    rollout_multi<Mdl, R, CH, NS, 1, 1, 1>(Xc, Kv, kv, al, prm, ub, T, keep, Xn, Jc);
                    for (int c = 0; c < CH; ++c)
                        if (!accepted && a0 + c < tries && ls_accept(Jc[c], J_last, ls_method == 1)) {
                            accepted = true; J_new = Jc[c]; taken = a0 + c; alpha_taken = ald[c]; replay = (c != keep);
                        }
                    a0 += CH;
                } else {
                    const double Jt = rollout_closed<Mdl, R, NS, 1, 1, 1>(Xc, Kv, kv, (R)alpha, prm, ub, T, Xn);
                    if (ls_method == 2 || ls_accept(Jt, J_last, ls_method == 1)) { accepted = true; J_new = Jt; taken = a0; }
                    alpha = alpha * gamma;
                    a0 += 1;
                }
            }
            if (replay) rollout_closed<Mdl, R, NS, 1, 1, 1>(Xc, Kv, kv, (R)alpha_taken, prm, ub, T, Xn);
            if (accepted) pred = taken;
            const bool stop = stop_test(J_new, J_last, criterion, stop_method == 0);
            J_last = J_new;
            if (J_trace) J_trace[it] = J_new;
            if (alpha_trace) alpha_trace[it] = taken;
            ++it;
            const bool finished = (stop && check_stop) || it >= max_iter;
            if (finished) {
                const R* src = accepted ? Xn.base : Xc.base;
                for (int i = 0; i < T * D; ++i) X_out[i] = src[i];
                *J_out = J_last; *iters_out = it; *conv_out = (stop && check_stop) ? 1 : 0;
                return;
            }
            if (!accepted) for (int i = 0; i < T * D; ++i) Xn.base[i] = Xc.base[i];
            cur ^= 1;
        }
    }
};

}  // namespace

#define HS_DISPATCH(model, ...)                                       \
    switch (model) {                                                  \
        case 0: { using Mdl = TwoLink; __VA_ARGS__; } break;          \
        case 1: { using Mdl = ThreeLink; __VA_ARGS__; } break;        \
        case 2: { using Mdl = RoboticArm; __VA_ARGS__; } break;       \
        default: return -2;                                           \
    }

#define HS_API(SUFFIX, R)                                                                                              \
    extern "C" int hs_rollout_##SUFFIX(int model, const R* x0, const R* u0, const R* params, int stride_t, double lo,  \
                                       double hi, int T, R* X, double* J) {                                            \
        HS_DISPATCH(model, *J = (Sim<Mdl, R>::rollout(x0, u0, params, stride_t, lo, hi, T, X)));                       \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int hs_backward_##SUFFIX(int model, const R* X, const R* params, int stride_t, int T, R* K, R* k) {     \
        HS_DISPATCH(model, (Sim<Mdl, R>::backward(X, params, stride_t, T, K, k)));                                     \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int hs_update_##SUFFIX(int model, const R* X, const R* K, const R* k, double alpha, const R* params,    \
                                      int stride_t, double lo, double hi, int T, R* Xn, double* J) {                   \
        HS_DISPATCH(model, *J = (Sim<Mdl, R>::update(X, K, k, alpha, params, stride_t, lo, hi, T, Xn)));               \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int hs_derivs_##SUFFIX(int model, const R* X, const R* params, int stride_t, int T, R* F, R* c, R* C) { \
        HS_DISPATCH(model, (Sim<Mdl, R>::derivs(X, params, stride_t, T, F, c, C)));                                    \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int hs_backward_dense_##SUFFIX(int n, int m, const R* F, const R* c, const R* C, int T, R* K, R* k) {   \
        if (n == 6 && m == 2) backward_dense<6, 2, R>(F, c, C, T, K, k);                                               \
        else if (n == 8 && m == 3) backward_dense<8, 3, R>(F, c, C, T, K, k);                                          \
        else if (n == 4 && m == 1) backward_dense<4, 1, R>(F, c, C, T, K, k);                                          \
        else if (n == 5 && m == 1) backward_dense<5, 1, R>(F, c, C, T, K, k);                                          \
        else if (n == 4 && m == 2) backward_dense<4, 2, R>(F, c, C, T, K, k);                                          \
        else return -2;                                                                                                \
        return 0;                                                                                                      \
    }                                                                                                                  \
    /* one STRUCTURED Riccati step (riccati_step, what the solve kernel runs) from a caller-supplied value function: \
       V [N][N] dense (block diagonal: state block + diagonal output block), v [N]; both are updated in place */     \
    extern "C" int hs_riccati_step_##SUFFIX(int model, const R* xu, const R* p, R* V, R* v, R* K, R* k) {             \
        HS_DISPATCH(model, (Sim<Mdl, R>::riccati(xu, p, V, v, K, k)));                                                 \
        return 0;                                                                                                      \
    }                                                                                                                  \
    extern "C" int hs_solve_##SUFFIX(int model, const R* x0, const R* u0, const R* params, int stride_t, double lo,    \
                                     double hi, int T, int max_iter, int check_stop, double criterion, int max_ls,     \
                                     double gamma, int ls_method, int stop_method, double J_init, R* X_out,            \
                                     double* J_out, int* iters_out, int* conv_out, double* J_trace, int* alpha_trace) { \
        HS_DISPATCH(model, (Sim<Mdl, R>::solve(x0, u0, params, stride_t, lo, hi, T, max_iter, check_stop, criterion,   \
                                               max_ls, gamma, ls_method, stop_method, J_init, X_out, J_out, iters_out, \
                                               conv_out, J_trace, alpha_trace)));                                      \
        return 0;                                                                                                      \
    }

extern "C" void hs_sincos_f64(const double* x, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) sincos_batch<1, double>(x + i, s + i, c + i);
}
extern "C" void hs_sincos_f32(const float* x, int n, float* s, float* c) {
    for (int i = 0; i < n; ++i) sincos_batch<1, float>(x + i, s + i, c + i);
}

HS_API(f64, double)
HS_API(f32, float)
