// Kernels for the GENERATED device models: CartPoleSwingUp1/2, VehicleTracking, CarParking.
//
// Device math: ilqr_generic.cuh (the reference's iteration in its general form) over ilqr_generated.cuh (written by
// curiousrl_b200/codegen.py from the scenarios' sympy expressions).  One thread = one trajectory, FP64.
//
//   stage kernels   API layout (dense rows), the reference's per-stage operator surface
//   gen_solve_kernel  the whole solve of a trajectory in one thread (basic_ilqr.py:68-97); the two trajectory
//                   buffers and the gains live in the workspace as warp tiles [t][component][32 lanes], so every
//                   access of a warp is one coalesced 256-byte row.  No tail engine: these scenarios are the
//                   widening row, correctness first; divergence between the lanes of a warp costs what the longest
//                   lane takes.
#include "ilqr_generic_host.h"
#include "ilqr_generic.cuh"

namespace crl {

namespace {

constexpr int kGenBlock = 128;

struct GenProblemDev {
    const double* params;
    long long stride_b;
    int stride_t;
    long long B;
    int T;
    GenBox box;
    __device__ GenParams prm(long long b) const { return GenParams{params + b * stride_b, stride_t, box}; }
};

// the scenario's default box, or the caller's (crl_problem::bounds_host, [d][2] rows of (lo, hi) in host memory)
GenProblemDev to_gen(const crl_problem_t* p) {
    GenProblemDev d{(const double*)p->params, p->params_stride_b, p->params_stride_t, p->batch, p->horizon, GenBox{}};
    int n = 0, m = 0;
    gen_dims(p->model, &n, &m, nullptr);
    gen_bounds(p->model, d.box.lo, d.box.hi);
    if (p->bounds_host) {
        for (int i = 0; i < n + m; ++i) { d.box.lo[i] = p->bounds_host[2 * i]; d.box.hi[i] = p->bounds_host[2 * i + 1]; }
    }
    return d;
}
GenOpts to_gen(const crl_solver_opts_t* o) {
    return GenOpts{o->max_iter, o->is_check_stop, o->max_line_search, o->line_search_method, o->stopping_method,
                   o->stopping_criterion, o->gamma};
}
unsigned blocks(long long work) { return (unsigned)((work + kGenBlock - 1) / kGenBlock); }

template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_rollout_kernel(GenProblemDev pb, const double* x0, const double* u0,
                                                                long long u0_sb, double* X, double* J) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    const double Jb = gen_rollout_open<G>(x0 + b * G::N, u0 ? u0 + b * u0_sb : nullptr, G::M, pb.prm(b), pb.T,
                                          View<double, G::D, 1>{X + b * pb.T * G::D});
    if (J) J[b] = Jb;
}

template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_cost_kernel(GenProblemDev pb, const double* X, double* J) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    J[b] = gen_cost<G>(View<double, G::D, 1>{const_cast<double*>(X) + b * pb.T * G::D}, pb.prm(b), pb.T);
}

template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_derivs_kernel(GenProblemDev pb, const double* X, double* F, double* c,
                                                               double* C) {
    constexpr int N = G::N, D = G::D;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (b, t)
    if (idx >= pb.B * pb.T) return;
    const long long b = idx / pb.T;
    const int t = (int)(idx % pb.T);
    double xu[D], p[G::P], Fl[N * D], cl[D], Cl[D * D];
    for (int i = 0; i < D; ++i) xu[i] = X[idx * D + i];
    pb.prm(b).load(t, p);
    gen_derivs<G>(xu, p, Fl, cl, Cl);
    if (F) for (int i = 0; i < N * D; ++i) F[idx * N * D + i] = Fl[i];
    if (c) for (int i = 0; i < D; ++i) c[idx * D + i] = cl[i];
    if (C) for (int i = 0; i < D * D; ++i) C[idx * D * D + i] = Cl[i];
}

template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_backward_kernel(GenProblemDev pb, const double* X, double* K, double* k) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    gen_backward<G>(View<double, G::D, 1>{const_cast<double*>(X) + b * pb.T * G::D}, pb.prm(b), pb.T,
                    View<double, G::M * G::N, 1>{K + b * pb.T * G::M * G::N}, View<double, G::M, 1>{k + b * pb.T * G::M});
}

template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_update_kernel(GenProblemDev pb, const double* X, const double* K,
                                                               const double* k, double alpha, double* Xn, double* Jn) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    const double J = gen_rollout_closed<G>(View<double, G::D, 1>{const_cast<double*>(X) + b * pb.T * G::D},
                                           View<double, G::M * G::N, 1>{const_cast<double*>(K) + b * pb.T * G::M * G::N},
                                           View<double, G::M, 1>{const_cast<double*>(k) + b * pb.T * G::M}, alpha,
                                           pb.prm(b), pb.T, View<double, G::D, 1>{Xn + b * pb.T * G::D});
    if (Jn) Jn[b] = J;
}

// line search + stopping test of one iteration (ilqr_wrapper.py:74-214 without the derivative re-evaluation)
template <class G>
__global__ void __launch_bounds__(kGenBlock) gen_forward_kernel(GenProblemDev pb, GenOpts op, const double* X,
                                                                const double* K, const double* k, const double* J_last,
                                                                double* Xn, double* J_new, int* alpha_index,
                                                                uint8_t* stop) {
    constexpr int D = G::D;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    const View<double, D, 1> Xo{const_cast<double*>(X) + b * pb.T * D}, Xw{Xn + b * pb.T * D};
    const View<double, G::M * G::N, 1> Kv{const_cast<double*>(K) + b * pb.T * G::M * G::N};
    const View<double, G::M, 1> kv{const_cast<double*>(k) + b * pb.T * G::M};
    const double Jl = J_last[b];
    double Jn = Jl, alpha = 1.0;
    int taken = -1;
    const int tries = (op.ls_method == CRL_LS_NONE) ? 1 : op.max_ls;
    for (int a = 0; a < tries && taken < 0; ++a) {
        const double Jt = gen_rollout_closed<G>(Xo, Kv, kv, alpha, pb.prm(b), pb.T, Xw);
        alpha = alpha * op.gamma;
        if (op.ls_method == CRL_LS_NONE || ls_accept(Jt, Jl, op.ls_method == CRL_LS_FEASIBILITY)) {
            taken = a;
            Jn = Jt;
        }
    }
    if (taken < 0) {   // no step accepted: the trajectory is returned unchanged (ilqr_wrapper.py:100)
        for (int i = 0; i < pb.T * D; ++i) Xw.base[i] = Xo.base[i];
    }
    J_new[b] = Jn;
    alpha_index[b] = taken;
    stop[b] = stop_test(Jn, Jl, op.criterion, op.stop_method == CRL_STOP_RELATIVE) ? 1 : 0;
}

template <class G>
constexpr long long gen_tile_reals(int T) {
    return (long long)T * 32 * (2 * G::D + G::M * G::N + G::M);
}

// The solves of the 32 trajectories of a warp, iteration by iteration in lock step (same results as gen_solve_one per
// trajectory, bit for bit: every candidate of the line search is gen_rollout_closed's arithmetic).  What it adds is the
// LINE-SEARCH HELPERS: a lane whose trajectory has finished, or has its step for this iteration, would idle while the
// warp runs one pass per step size for its slowest lane - and every trajectory's LAST iteration rejects all
// max_line_search step sizes (ncu of the per-thread kernel on VehicleTracking: 14.3 of 32 lanes active in the rollouts).
// Here the idle lanes evaluate the NEXT step sizes of the trajectories that are still searching, cost only, reading the
// owner's columns of X, K, k (same sectors: no extra traffic): with n_pend searching lanes and n_help idle ones a pass
// covers 1 + n_help / n_pend step sizes per trajectory.  If a helper's step size is the first improving one the owner
// repeats that rollout with stores (one more pass for the warp, instead of one per skipped step size).
template <class G>
__device__ void gen_solve_warp(bool valid, const double* x0, const double* U, int su_t, GenParams prm, int T, const GenOpts& op,
                               double J_init, View<double, G::D, 32> Xa, View<double, G::D, 32> Xb,
                               View<double, G::M * G::N, 32> K, View<double, G::M, 32> k, double* J_out, int* iters_out,
                               uint8_t* conv_out, bool* final_in_a, double* J_trace, signed char* alpha_trace, int n_stage,
                               long long stage_stride, int* stage_iters) {
    constexpr unsigned kAll = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (valid) gen_rollout_open<G>(x0, U, su_t, prm, T, Xa);
    bool active = valid, in_a = true, stopped = false;
    double J_last = J_init;
    const int tries = op.ls_method == 2 /* CRL_LS_NONE */ ? 1 : op.max_ls;
    int s = 0, it = 0, total = 0;
    GenParams ps = prm;
    while (__any_sync(kAll, active)) {
        const View<double, G::D, 32> Xc = in_a ? Xa : Xb, Xn = in_a ? Xb : Xa;
        if (active) gen_backward<G>(Xc, ps, T, K, k, it == 0);
        __syncwarp();                                           // the helpers read the owner's K, k (and its X of the last pass)
        bool accepted = false, replay = false;
        double J_new = J_last, alpha = 1.0, alpha_taken = 1.0;
        int taken = -1, a0 = 0;
        while (a0 < tries) {
            const bool pending = active && !accepted;
            const unsigned pm = __ballot_sync(kAll, pending);
            if (pm == 0u) break;
            const int n_pend = __popc(pm);
            int per = (32 - n_pend) / n_pend;                   // helpers per searching trajectory
            if (per > tries - a0 - 1) per = tries - a0 - 1;
            const unsigned hm = ~pm;
            const int rank_p = __popc(pm & lt), rank_h = __popc(hm & lt);
            const bool helper = !pending && rank_h < per * n_pend;
            const int owner = helper ? (int)__fns(pm, 0, rank_h % n_pend + 1) : lane;
            const int j = helper ? 1 + rank_h / n_pend : 0;     // this lane's step size: index a0 + j
            // the owner's problem
            const double Jl_o = __shfl_sync(kAll, J_last, owner);
            const bool in_a_o = __shfl_sync(kAll, (int)in_a, owner) != 0;
            GenParams ps_o = ps;
            ps_o.ptr = (const double*)(uintptr_t)__shfl_sync(kAll, (unsigned long long)(uintptr_t)ps.ptr, owner);
            double Jt = 0.0;
            bool ok = false;
            if (pending || helper) {
                double al = alpha;
                for (int i = 0; i < j; ++i) al = al * op.gamma;
                const int shift = owner - lane;
                const View<double, G::D, 32> Xc_o{(in_a_o ? Xa.base : Xb.base) + shift};
                const View<double, G::M * G::N, 32> K_o{K.base + shift};
                const View<double, G::M, 32> k_o{k.base + shift};
                Jt = gen_rollout_closed<G>(Xc_o, K_o, k_o, al, ps_o, T, Xn, pending);
                ok = op.ls_method == 2 || ls_accept(Jt, Jl_o, op.ls_method == 1);
            }
            const unsigned okm = __ballot_sync(kAll, ok);
            int w = -1, src = lane;
            if (pending) {
                if ((okm >> lane) & 1u) {
                    w = 0;
                } else {
                    for (int jj = 0; jj < per && w < 0; ++jj) {
                        const unsigned hl = __fns(hm, 0, rank_p + jj * n_pend + 1);
                        if (hl < 32u && ((okm >> hl) & 1u)) { w = jj + 1; src = (int)hl; }
                    }
                }
            }
            const double Jw = __shfl_sync(kAll, Jt, src);
            if (pending && w >= 0) {
                accepted = true;
                taken = a0 + w;
                J_new = Jw;
                if (w > 0) {
                    replay = true;
                    double al = alpha;
                    for (int i = 0; i < w; ++i) al = al * op.gamma;
                    alpha_taken = al;
                }
            }
            for (int i = 0; i < 1 + per; ++i) alpha = alpha * op.gamma;
            a0 += 1 + per;
        }
        if (__any_sync(kAll, replay)) {
            if (replay) gen_rollout_closed<G>(Xc, K, k, alpha_taken, ps, T, Xn, true);
        }
        if (active) {
            const bool stop = stop_test(J_new, J_last, op.criterion, op.stop_method == 0);
            J_last = J_new;
            if (accepted) in_a = !in_a;
            if (J_trace) J_trace[it] = J_new;
            if (alpha_trace) alpha_trace[it] = (signed char)taken;
            ++it;
            if (stop && op.check_stop) stopped = true;
            if (stopped || it >= op.max_iter) {               // this inner loop is over (log_barrier_ilqr.py:121-146)
                if (stage_iters) stage_iters[s] = it;
                total += it;
                ++s;
                if (s < n_stage) {
                    ps.ptr = prm.ptr + (size_t)s * stage_stride;
                    if (stopped) J_last = (double)INFINITY;
                    it = 0;
                    stopped = false;
                } else {
                    active = false;
                }
            }
        }
        __syncwarp();
    }
    if (valid) {
        *J_out = J_last;
        *iters_out = total;
        *conv_out = stopped ? (uint8_t)1 : (uint8_t)0;
    }
    *final_in_a = in_a;
}

#ifndef CRL_GEN_MINB
#define CRL_GEN_MINB 4
#endif
// 4 resident CTAs per SM (128 registers per thread): 75,776 resident threads, so the benchmark batch of 65,536 is one wave
// WARPSYNC = true: gen_solve_warp (lock step + line-search helpers, the default); false: gen_solve_one per thread
// (opts->reserved == 2, "warps free-running": the checker of the other one)
template <class G, bool WARPSYNC>
__global__ void __launch_bounds__(kGenBlock, CRL_GEN_MINB) gen_solve_kernel(GenProblemDev pb, GenOpts op, const double* x0,
                                                              const double* u0, long long u0_sb, const double* J_init,
                                                              double* X_out, double* J_out, int* iters_out,
                                                              uint8_t* conv_out, double* J_trace, int8_t* alpha_trace,
                                                              double* ws, int n_stage, long long stage_stride,
                                                              int* stage_iters) {
    constexpr int D = G::D, M = G::M, N = G::N;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int T = pb.T;
    double* const tile = ws + (b >> 5) * gen_tile_reals<G>(T) + lane;
    const View<double, D, 32> Xa{tile}, Xb{tile + (size_t)T * D * 32};
    const View<double, M * N, 32> Kv{tile + (size_t)2 * T * D * 32};
    const View<double, M, 32> kv{tile + (size_t)2 * T * D * 32 + (size_t)T * M * N * 32};
    bool in_a = true;
    if constexpr (WARPSYNC) {
        const bool valid = b < pb.B;
        const long long bb = valid ? b : 0;
        gen_solve_warp<G>(valid, x0 + bb * N, u0 ? u0 + bb * u0_sb : nullptr, M, pb.prm(bb), T, op,
                          J_init ? J_init[bb] : (double)INFINITY, Xa, Xb, Kv, kv, J_out + bb, iters_out + bb,
                          conv_out + bb, &in_a, J_trace ? J_trace + bb * op.max_iter : nullptr,
                          alpha_trace ? (signed char*)alpha_trace + bb * op.max_iter : nullptr, n_stage, stage_stride,
                          stage_iters ? stage_iters + bb * n_stage : nullptr);
    } else if (b < pb.B) {
        double J = 0.0;
        int iters = 0, conv = 0;
        gen_solve_one<G>(x0 + b * N, u0 ? u0 + b * u0_sb : nullptr, M, pb.prm(b), T, op,
                         J_init ? J_init[b] : (double)INFINITY, Xa, Xb, Kv, kv, &J, &iters, &conv, &in_a,
                         J_trace ? J_trace + b * op.max_iter : nullptr,
                         alpha_trace ? (signed char*)alpha_trace + b * op.max_iter : nullptr, n_stage, stage_stride,
                         stage_iters ? stage_iters + b * n_stage : nullptr);
        J_out[b] = J;
        iters_out[b] = iters;
        conv_out[b] = (uint8_t)conv;
    }
    if (X_out) {
        // final trajectories leave the tile through the whole warp: 32 consecutive elements per store
        __syncwarp();
        const long long b0 = b - lane;
        for (int owner = 0; owner < 32; ++owner) {
            const long long bo = b0 + owner;
            if (bo >= pb.B) break;
            const bool a_o = __shfl_sync(0xffffffffu, (int)in_a, owner) != 0;
            const double* col = (a_o ? Xa.base : Xb.base) - lane + owner;
            double* dst = X_out + bo * T * D;
            // 8 loads in flight before the stores (one load per store is one DRAM round trip per element: the compiler
            // must assume that the tile and X_out alias)
            for (int e0 = lane; e0 < T * D; e0 += 32 * 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (e0 + 32 * j < T * D) v[j] = col[(size_t)(e0 + 32 * j) * 32];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (e0 + 32 * j < T * D) dst[e0 + 32 * j] = v[j];
            }
        }
    }
}

}  // namespace

#ifdef CRL_USER_MODEL
// the run-time compiled unit holds exactly one scenario: ids CRL_GEN_USER / CRL_GEN_USER_BARRIER
#define CRL_GEN_DISPATCH(model, ...)                                                           \
    switch (model) {                                                                           \
        case CRL_GEN_USER: { using G = GenUser; __VA_ARGS__; } break;                          \
        case CRL_GEN_USER_BARRIER: { using G = GenUserBarrier; __VA_ARGS__; } break;           \
        default: break;                                                                        \
    }

bool gen_is_model(int model) { return model == CRL_GEN_USER || model == CRL_GEN_USER_BARRIER; }
bool gen_is_barrier(int model) { return model == CRL_GEN_USER_BARRIER; }
#else
#define CRL_GEN_DISPATCH(model, ...)                                                           \
    switch (model) {                                                                           \
        case CRL_GEN_CART_POLE_1: { using G = GenCartPoleSwingUp1; __VA_ARGS__; } break;       \
        case CRL_GEN_CART_POLE_2: { using G = GenCartPoleSwingUp2; __VA_ARGS__; } break;       \
        case CRL_GEN_VEHICLE_TRACKING: { using G = GenVehicleTracking; __VA_ARGS__; } break;   \
        case CRL_GEN_CAR_PARKING: { using G = GenCarParking; __VA_ARGS__; } break;             \
        case CRL_GEN_CART_POLE_1_BARRIER: { using G = GenCartPoleSwingUp1Barrier; __VA_ARGS__; } break;      \
        case CRL_GEN_CART_POLE_2_BARRIER: { using G = GenCartPoleSwingUp2Barrier; __VA_ARGS__; } break;      \
        case CRL_GEN_VEHICLE_TRACKING_BARRIER: { using G = GenVehicleTrackingBarrier; __VA_ARGS__; } break;  \
        case CRL_GEN_CAR_PARKING_BARRIER: { using G = GenCarParkingBarrier; __VA_ARGS__; } break;            \
        default: break;                                                                        \
    }

bool gen_is_model(int model) { return model >= CRL_GEN_CART_POLE_1 && model <= CRL_GEN_CAR_PARKING_BARRIER; }
bool gen_is_barrier(int model) { return model >= CRL_GEN_CART_POLE_1_BARRIER && model <= CRL_GEN_CAR_PARKING_BARRIER; }
#endif

int gen_dims(int model, int* n, int* m, int* p) {
    if (!gen_is_model(model)) return -1;
    CRL_GEN_DISPATCH(model, { if (n) *n = G::N; if (m) *m = G::M; if (p) *p = G::P; });
    return 0;
}

int gen_bounds(int model, double* lo, double* hi) {
    if (!gen_is_model(model)) return -1;
    CRL_GEN_DISPATCH(model, { for (int i = 0; i < G::D; ++i) { if (lo) lo[i] = G::lo(i); if (hi) hi[i] = G::hi(i); } });
    return 0;
}

int gen_launch_rollout(const crl_problem_t* prob, const double* x0, const double* u0, long long u0_sb, double* X,
                       double* J, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_rollout_kernel<G><<<blocks(prob->batch), kGenBlock, 0, s>>>(to_gen(prob), x0, u0, u0_sb, X, J)));
    return (int)cudaGetLastError();
}
int gen_launch_cost(const crl_problem_t* prob, const double* X, double* J, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_cost_kernel<G><<<blocks(prob->batch), kGenBlock, 0, s>>>(to_gen(prob), X, J)));
    return (int)cudaGetLastError();
}
int gen_launch_derivs(const crl_problem_t* prob, const double* X, double* F, double* c, double* C, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_derivs_kernel<G><<<blocks(prob->batch * prob->horizon), kGenBlock, 0, s>>>(to_gen(prob), X, F, c, C)));
    return (int)cudaGetLastError();
}
int gen_launch_backward(const crl_problem_t* prob, const double* X, double* K, double* k, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_backward_kernel<G><<<blocks(prob->batch), kGenBlock, 0, s>>>(to_gen(prob), X, K, k)));
    return (int)cudaGetLastError();
}
int gen_launch_update(const crl_problem_t* prob, const double* X, const double* K, const double* k, double alpha,
                      double* Xn, double* Jn, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_update_kernel<G><<<blocks(prob->batch), kGenBlock, 0, s>>>(to_gen(prob), X, K, k, alpha, Xn, Jn)));
    return (int)cudaGetLastError();
}
int gen_launch_forward(const crl_problem_t* prob, const crl_solver_opts_t* opts, const double* X, const double* K,
                       const double* k, const double* J_last, double* Xn, double* J_new, int32_t* alpha_index,
                       uint8_t* stop, cudaStream_t s) {
    CRL_GEN_DISPATCH(prob->model, (gen_forward_kernel<G><<<blocks(prob->batch), kGenBlock, 0, s>>>(
                                      to_gen(prob), to_gen(opts), X, K, k, J_last, Xn, J_new, alpha_index, stop)));
    return (int)cudaGetLastError();
}

size_t gen_solve_workspace_bytes(const crl_problem_t* prob) {
    long long reals = 0;
    CRL_GEN_DISPATCH(prob->model, (reals = gen_tile_reals<G>(prob->horizon)));
    const size_t warps = (size_t)((prob->batch + 31) / 32);
    return (warps > 0 ? warps : 1) * (size_t)reals * sizeof(double);
}

int gen_launch_solve(const crl_problem_t* prob, const crl_solver_opts_t* opts, const double* x0, const double* u0,
                     long long u0_sb, const double* J_init, double* X, double* J, int32_t* iters, uint8_t* converged,
                     double* J_trace, int8_t* alpha_trace, void* workspace, cudaStream_t s, int n_stage,
                     int32_t* stage_iters) {
    // whole warps: the result write-out is a warp-cooperative loop
    const unsigned grid = blocks(((prob->batch + 31) / 32) * 32);
    if (opts->reserved == 2) {
        CRL_GEN_DISPATCH(prob->model, (gen_solve_kernel<G, false><<<grid, kGenBlock, 0, s>>>(
                                          to_gen(prob), to_gen(opts), x0, u0, u0_sb, J_init, X, J, iters, converged, J_trace,
                                          alpha_trace, (double*)workspace, n_stage,
                                          prob->params_stride_t ? (long long)prob->horizon * G::P : (long long)G::P,
                                          stage_iters)));
    } else {
        CRL_GEN_DISPATCH(prob->model, (gen_solve_kernel<G, true><<<grid, kGenBlock, 0, s>>>(
                                          to_gen(prob), to_gen(opts), x0, u0, u0_sb, J_init, X, J, iters, converged, J_trace,
                                          alpha_trace, (double*)workspace, n_stage,
                                          prob->params_stride_t ? (long long)prob->horizon * G::P : (long long)G::P,
                                          stage_iters)));
    }
    return (int)cudaGetLastError();
}

}  // namespace crl
