// CUDA kernels + C ABI of the batched-iLQR path (sm_100a).
//
// Mapping: one lane = one trajectory.  The matrices of the Riccati recursion (n <= 8,
// m <= 3) live in that lane's registers; a warp therefore advances 32 trajectories in
// lock step and every global access is a 32-lane row of one component (SoA tiles), i.e.
// one or two full 128-byte lines.  Tensor cores are not used: the contractions are 2x2 ...
// 8x8 with structural zeros, far below one MMA tile.
//
// Kernels
//   stage kernels (API layout [B][T][..], one thread per trajectory): rollout, cost, derivs,
//     backward, backward_dense, update_traj, forward_pass - the reference's operator surface,
//     used for B = 1 compatibility calls and stage-parity tests;
//   solve_kernel: the product path.  Persistent grid; each lane pulls trajectories from a
//     global counter, keeps its working set (X_cur, X_new, K, k) in a warp-tile workspace slot
//     and iterates {backward, multi-alpha line search, accept/stop} until its trajectory
//     stops, then writes the result and pulls the next one.  No host round trip, no
//     compaction pass, memory footprint independent of B.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/curiousrl_b200.h"
#include "ilqr_core.cuh"
#include "ilqr_coop.cuh"
#include "ilqr_generic_host.h"

namespace crl {

constexpr int kBlock = 128;       // threads per CTA (4 warps)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCoopMax = 6;        // a drained warp with <= this many live trajectories works on them cooperatively
#ifndef CRL_LS_CHUNK
#define CRL_LS_CHUNK 4
#endif
constexpr int kLsChunk = CRL_LS_CHUNK;   // step sizes evaluated per pass of the line search
#ifndef CRL_HELP_MAX
#define CRL_HELP_MAX 16
#endif
#ifndef CRL_SPLIT_ROLLOUT
#define CRL_SPLIT_ROLLOUT 1
#endif
#ifndef CRL_SPLIT_UNROLL
#define CRL_SPLIT_UNROLL 2
#endif
#ifndef CRL_SPLIT_MAX_LIVE
#define CRL_SPLIT_MAX_LIVE 1
#endif
constexpr int kSplitMaxLive = CRL_SPLIT_MAX_LIVE;        // ... for warps that hold at most this many trajectories
constexpr bool kSplitRollout = CRL_SPLIT_ROLLOUT != 0;   // two-phase line search for a warp's lone trajectory (coop_recurrence)
constexpr int kHelpMax = CRL_HELP_MAX;   // after the first pass, at most this many still-searching lanes are helped by the others
constexpr bool kHelpLs = kHelpMax > 0;

template <class R>
struct ProblemDev {
    const R* params;
    long long params_stride_b;
    int params_stride_t;
    long long B;
    int T;
    Bounds<R> ub;
    __device__ ParamRef<R> prm(long long b) const { return ParamRef<R>{params + b * params_stride_b, params_stride_t}; }
};

// ------------------------------------------------------------------------------ stage kernels
template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) rollout_kernel(ProblemDev<R> pb, const R* x0, const R* u0, long long u0_sb,
                                                         R* X, double* J) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    View<R, Mdl::D, 1> Xv{X + b * pb.T * Mdl::D};
    const double Jb = rollout_open<Mdl, R, 1>(x0 + b * Mdl::N, 1, u0 ? u0 + b * u0_sb : nullptr, Mdl::M, pb.prm(b),
                                              pb.ub, pb.T, Xv);
    if (J) J[b] = Jb;
}

template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) cost_kernel(ProblemDev<R> pb, const R* X, double* J) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    const ParamRef<R> prm = pb.prm(b);
    const R* Xb = X + b * pb.T * Mdl::D;
    double acc = 0.0;
    for (int t = 0; t < pb.T; ++t) {
        R xu[Mdl::D];
#pragma unroll
        for (int i = 0; i < Mdl::D; ++i) xu[i] = Xb[(size_t)t * Mdl::D + i];
        acc = acc + step_cost<Mdl, R>(xu, prm, t);
    }
    J[b] = acc;
}

template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) derivs_kernel(ProblemDev<R> pb, const R* X, R* F, R* c, R* C) {
    constexpr int N = Mdl::N, D = Mdl::D, P = Mdl::P;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (b, t)
    if (idx >= pb.B * pb.T) return;
    const long long b = idx / pb.T;
    const int t = (int)(idx % pb.T);
    const ParamRef<R> prm = pb.prm(b);
    R xu[D], p[P], Fl[N * D], cl[D], Cl[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i) xu[i] = X[idx * D + i];
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = prm.get(t, i);
    derivs_dense<Mdl, R>(xu, p, Fl, cl, Cl);
    if (F) for (int i = 0; i < N * D; ++i) F[idx * N * D + i] = Fl[i];
    if (c) for (int i = 0; i < D; ++i) c[idx * D + i] = cl[i];
    if (C) for (int i = 0; i < D * D; ++i) C[idx * D * D + i] = Cl[i];
}

template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) backward_kernel(ProblemDev<R> pb, const R* X, R* K, R* k) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    View<R, Mdl::D, 1> Xv{const_cast<R*>(X) + b * pb.T * Mdl::D};
    View<R, Mdl::M * Mdl::N, 1> Kv{K + b * pb.T * Mdl::M * Mdl::N};
    View<R, Mdl::M, 1> kv{k + b * pb.T * Mdl::M};
    backward_fused<Mdl, R, Mdl::N, 1, 1>(Xv, pb.prm(b), pb.T, Kv, kv);
}

template <int N, int M, class R>
__global__ void __launch_bounds__(kBlock) backward_dense_kernel(long long B, int T, const R* F, const R* c, const R* C,
                                                                R* K, R* k) {
    constexpr int D = N + M;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    backward_dense<N, M, R>(F + b * T * N * D, c + b * T * D, C + b * T * D * D, T, K + b * T * M * N, k + b * T * M);
}

template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) update_traj_kernel(ProblemDev<R> pb, const R* X, const R* K, const R* k,
                                                             double alpha, R* Xn, double* Jn) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    View<R, Mdl::D, 1> Xo{const_cast<R*>(X) + b * pb.T * Mdl::D};
    View<R, Mdl::D, 1> Xw{Xn + b * pb.T * Mdl::D};
    View<R, Mdl::M * Mdl::N, 1> Kv{const_cast<R*>(K) + b * pb.T * Mdl::M * Mdl::N};
    View<R, Mdl::M, 1> kv{const_cast<R*>(k) + b * pb.T * Mdl::M};
    const double J = rollout_closed<Mdl, R, Mdl::N, 1, 1, 1>(Xo, Kv, kv, (R)alpha, pb.prm(b), pb.ub, pb.T, Xw);
    if (Jn) Jn[b] = J;
}

struct OptsDev {
    int max_iter, check_stop, max_ls, ls_method, stop_method;
    double criterion, gamma;
};

template <class Mdl, class R>
__global__ void __launch_bounds__(kBlock) forward_pass_kernel(ProblemDev<R> pb, OptsDev op, const R* X, const R* K,
                                                              const R* k, const double* J_last, R* Xn, double* J_new,
                                                              int* alpha_index, uint8_t* stop) {
    constexpr int D = Mdl::D;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pb.B) return;
    View<R, D, 1> Xo{const_cast<R*>(X) + b * pb.T * D};
    View<R, D, 1> Xw{Xn + b * pb.T * D};
    View<R, Mdl::M * Mdl::N, 1> Kv{const_cast<R*>(K) + b * pb.T * Mdl::M * Mdl::N};
    View<R, Mdl::M, 1> kv{const_cast<R*>(k) + b * pb.T * Mdl::M};
    const ParamRef<R> prm = pb.prm(b);
    const double Jl = J_last[b];
    double Jn = Jl, alpha = 1.0;
    int taken = -1;
    const int tries = (op.ls_method == CRL_LS_NONE) ? 1 : op.max_ls;
    for (int a = 0; a < tries && taken < 0; ++a) {
        const double Jt = rollout_closed<Mdl, R, Mdl::N, 1, 1, 1>(Xo, Kv, kv, (R)alpha, prm, pb.ub, pb.T, Xw);
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

// ------------------------------------------------------------------------------ smem staging (TMA)
// Bulk asynchronous copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) stream the next time
// step's contiguous tile blocks into shared memory while the current step computes; completion is
// tracked by an mbarrier per stage.  One elected lane issues the copies, every lane of the warp
// that takes part waits on the barrier and then reads its own column (conflict-free LDS).
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_reinit(unsigned long long* bar) {
    const unsigned a = smem_addr(bar);
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = smem_addr(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// Per-warp staging area: two stages of one time step's X / K / k blocks + two mbarriers.
template <class Mdl, class R>
struct StageLayout {
    static constexpr int XB = Mdl::D * 32, KB = Mdl::M * Mdl::NS * 32, kB = Mdl::M * 32;   // reals per block
    static constexpr int STAGE = XB + KB + kB;
    static constexpr int kWarpBytes = 2 * STAGE * (int)sizeof(R) + 128;                     // + barriers (padded)
};

// Closed-loop rollout inputs of the warp's own tiles (KN = NS).
template <class Mdl, class R>
struct TileFeed {
    using L = StageLayout<Mdl, R>;
    static constexpr int kKN = Mdl::NS;
    const R* gX;                  // column 0, t = 0 of the old trajectory tile
    const R* gK;
    const R* gk;
    R* sm;                        // this warp's two stages
    unsigned long long* bar;      // this warp's two barriers
    unsigned mask;
    int lane, col, leader, nsteps;

    // `mask_` = the lanes that take part, computed by the caller where the warp is still converged;
    // col_ = the tile column (trajectory) this lane reads: its own, or another lane's when it helps
    // with that lane's line search
    __device__ TileFeed(unsigned mask_, const R* gX_, const R* gK_, const R* gk_, R* sm_, unsigned long long* bar_,
                        int col_ = -1)
        : gX(gX_), gK(gK_), gk(gk_), sm(sm_), bar(bar_) {
        mask = mask_;
        lane = threadIdx.x & 31;
        col = col_ < 0 ? lane : col_;
        leader = __ffs(mask) - 1;
        nsteps = 0;
    }
    __device__ void issue(int t) {
        const int st = t & 1;
        R* dst = sm + st * L::STAGE;
        mbar_expect_tx(bar + st, (unsigned)(L::STAGE * sizeof(R)));
        bulk_g2s(dst, gX + (size_t)t * L::XB, L::XB * sizeof(R), bar + st);
        bulk_g2s(dst + L::XB, gK + (size_t)t * L::KB, L::KB * sizeof(R), bar + st);
        bulk_g2s(dst + L::XB + L::KB, gk + (size_t)t * L::kB, L::kB * sizeof(R), bar + st);
    }
    __device__ void begin(int n) {
        nsteps = n;
        __syncwarp(mask);
        if (lane == leader) {
            mbar_reinit(bar);
            mbar_reinit(bar + 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (n > 0) issue(0);
        }
        __syncwarp(mask);
    }
    __device__ void end() { __syncwarp(mask); }
    __device__ void row0(R* xu) const {
#pragma unroll
        for (int i = 0; i < Mdl::D; ++i) xu[i] = gX[i * 32 + col];
    }
    __device__ void get(int t, R* xo, R* uo, R* Kc, R* kc) {
        __syncwarp(mask);                                   // everyone is done with the stage about to be refilled
        if (lane == leader && t + 1 < nsteps) issue(t + 1);
        mbar_wait(bar + (t & 1), (unsigned)((t >> 1) & 1));
        const R* s0 = sm + (t & 1) * L::STAGE + col;
#pragma unroll
        for (int j = 0; j < Mdl::NS; ++j) xo[j] = s0[j * 32];
#pragma unroll
        for (int a = 0; a < Mdl::M; ++a) uo[a] = s0[(Mdl::N + a) * 32];
#pragma unroll
        for (int j = 0; j < Mdl::M * Mdl::NS; ++j) Kc[j] = s0[L::XB + j * 32];
#pragma unroll
        for (int a = 0; a < Mdl::M; ++a) kc[a] = s0[L::XB + L::KB + a * 32];
    }
};

// Rows of the current trajectory tile, last to first, for the backward pass.
template <class Mdl, class R>
struct TileRowFeed {
    using L = StageLayout<Mdl, R>;
    const R* gX;
    R* sm;
    unsigned long long* bar;
    unsigned mask;
    int lane, leader, nrows;
    __device__ TileRowFeed(unsigned mask_, const R* gX_, R* sm_, unsigned long long* bar_)
        : gX(gX_), sm(sm_), bar(bar_) {
        mask = mask_;
        lane = threadIdx.x & 31;
        leader = __ffs(mask) - 1;
        nrows = 0;
    }
    __device__ void issue(int i) {
        const int st = i & 1;
        mbar_expect_tx(bar + st, (unsigned)(L::XB * sizeof(R)));
        bulk_g2s(sm + st * L::STAGE, gX + (size_t)(nrows - 1 - i) * L::XB, L::XB * sizeof(R), bar + st);
    }
    __device__ void begin(int n) {
        nrows = n;
        __syncwarp(mask);
        if (lane == leader) {
            mbar_reinit(bar);
            mbar_reinit(bar + 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (n > 0) issue(0);
        }
        __syncwarp(mask);
    }
    __device__ void end() { __syncwarp(mask); }
    __device__ void get(int i, R* xu) {
        __syncwarp(mask);
        if (lane == leader && i + 1 < nrows) issue(i + 1);
        mbar_wait(bar + (i & 1), (unsigned)((i >> 1) & 1));
        const R* s0 = sm + (i & 1) * L::STAGE + lane;
#pragma unroll
        for (int c = 0; c < Mdl::D; ++c) xu[c] = s0[c * 32];
    }
};

// ------------------------------------------------------------------------------ persistent solve
// One straggler handed from lane mode to the cooperative finishers (ilqr_coop.cuh).
struct QueueEntry {
    long long traj;                // index in the batch
    const void* src;               // element (t = 0, c = 0) of its current trajectory: a tile column, stride 32
    double J_last;
    int it;                        // iterations done so far
    int ready;                     // published (written last, after a fence)
};
static_assert(sizeof(QueueEntry) == 32, "queue entries are 32 bytes");

template <class R>
struct SolveArgs {
    ProblemDev<R> pb;
    OptsDev op;
    const R* x0;
    const R* u0;
    long long u0_sb;
    const double* J_init;
    R* X_out;
    double* J_out;
    int* iters_out;
    uint8_t* conv_out;
    double* J_trace;
    int8_t* alpha_trace;
    R* ws;                         // [slots][slot_elems]
    long long slot_elems;          // reals per warp slot
    unsigned long long* ctl;       // control block (zeroed per launch): see kCtl*
    QueueEntry* queue;             // straggler queue, one entry per lane of the grid
    int coop_threshold;            // a drained warp with <= this many live trajectories hands them over; < 0 = never
    // Stage schedule (crl_solve_stages; barrier models): trajectory b runs one inner loop per parameter row
    // params[b][0 .. n_stage) and keeps its trajectory in place between them.  n_stage <= 1 = a plain solve.
    int n_stage;
    int* stage_iters;              // [B][n_stage] iterations of every inner loop
};

// Stage schedules pack (stage, iteration) into the queue entry's / the engine's iteration word.
constexpr int kStageShift = 24;
constexpr int kStageItMask = (1 << kStageShift) - 1;

// An inner loop of trajectory `traj` ended after `it` iterations in stage `stage`.  Returns true if the
// trajectory goes on with the next stage (log_barrier_ilqr.py:121-146: same trajectory, next parameter row,
// iteration counter back to 0, stored cost reset to +inf after a converged loop); otherwise writes the total
// iteration count to *total.
template <class R>
__device__ __forceinline__ bool stage_advance(const SolveArgs<R>& a, long long traj, int stage, int it, int* total) {
    __stcg(a.stage_iters + traj * a.n_stage + stage, it);
    if (stage + 1 < a.n_stage) return true;
    int sum = 0;
    for (int s = 0; s < a.n_stage; ++s) sum += __ldcg(a.stage_iters + traj * a.n_stage + s);
    *total = sum;
    return false;
}

// Control block words.
constexpr int kCtlNext = 0;        // next unassigned trajectory of the batch
constexpr int kCtlTail = 1;        // straggler queue: entries reserved by producers
constexpr int kCtlHead = 2;        // straggler queue: entries claimed by finishers
constexpr int kCtlDone = 3;        // warps that have left lane mode (and published their stragglers)
constexpr int kCtlFin = 4;         // straggler-queue trajectories finished
constexpr size_t kCtlBytes = 256;

// Optional phase timers (development builds, -DCRL_PHASE_TIMERS): lane 0 of every warp adds the
// SM cycles it spent per phase to control-block words kCtlTimers + phase.
[[maybe_unused]] constexpr int kCtlTimers = 8;
enum Phase : int { PH_REFILL = 0, PH_BACKWARD, PH_LINESEARCH, PH_RETIRE, PH_COOP_WAIT, PH_COOP_GATHER, PH_COOP_BACKWARD,
                   PH_COOP_LINESEARCH, PH_COOP_OUTPUT, PH_COUNT };
#ifdef CRL_PHASE_TIMERS
struct PhaseTimer {
    long long acc[PH_COUNT] = {};
    long long t0;
    __device__ static long long now() {
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
        return t;
    }
    __device__ PhaseTimer() : t0(now()) {}
    __device__ void lap(int phase) {
        const long long t = now();
        acc[phase] += t - t0;
        t0 = t;
    }
    __device__ void flush(unsigned long long* ctl, int lane) {
        if (lane == 0)
            for (int i = 0; i < PH_COUNT; ++i) atomicAdd(ctl + kCtlTimers + i, (unsigned long long)acc[i]);
    }
};
#else
struct PhaseTimer {
    __device__ void lap(int) {}
    __device__ void flush(unsigned long long*, int) {}
};
#endif

template <class Mdl>
__host__ __device__ constexpr long long lane_tiles_reals(int T) {
    return (long long)T * 32 * (2 * Mdl::D + Mdl::M * Mdl::NS + Mdl::M);
}

// Warp-cooperative copy of one trajectory between layouts: element e of the destination comes from src_at(e).  The loads
// of kCopyBatch elements per lane are issued before any of them is stored: written as `dst[e] = src[...]` the compiler must
// assume the two arrays alias and waits for every load (a DRAM access from a tile that left the L2 long ago) before the
// next one - 35 serialised round trips per trajectory (measured: the retire phase was 14 % of lane mode).
constexpr int kCopyBatch = 8;
template <class R, class SrcAt>
__device__ __forceinline__ void warp_copy_elems(R* dst, int n, int lane, SrcAt src_at) {
    for (int e0 = lane; e0 < n; e0 += 32 * kCopyBatch) {
        R v[kCopyBatch];
#pragma unroll
        for (int j = 0; j < kCopyBatch; ++j) {
            const int e = e0 + 32 * j;
            if (e < n) v[j] = *src_at(e);
        }
#pragma unroll
        for (int j = 0; j < kCopyBatch; ++j) {
            const int e = e0 + 32 * j;
            if (e < n) dst[e] = v[j];
        }
    }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// ------------------------------------------------------------------------------ cooperative engine
// One warp iterates NT trajectories at a time, every lane working on every iteration
// (ilqr_coop.cuh): the NT backward passes are stepped together (coop_backward<NT>), the line
// search gives each trajectory GL = 32 / NT lanes, one step size per lane, every candidate
// rollout stored in its own scratch buffer so that the accepted one simply becomes the current
// trajectory.  Finished trajectories are replaced from a Source:
//   BatchSource      fresh trajectories of the batch (coop_solve_kernel: the whole solve is done
//                    this way - lowest latency per iteration, used when the batch is too small
//                    to amortise the lane-per-trajectory kernel's long tail);
//   StragglerSource  the straggler queue of the lane-per-trajectory kernel (its tail).
// Scratch of one warp: xscr = per trajectory two line-search buffers [T][GL][DP] (the GL candidates of
// a pass side by side; DP = D padded to whole 32-byte sectors, so a candidate's row is a run of sectors
// of its own and the current trajectory - one candidate of one of the buffers, row stride GL * DP - is
// read without touching the other candidates' bytes; an element-interleaved [T][D][GL] layout moved a
// sector per element, 4-8x the bytes); gscr = per trajectory K [T][M*NS], k [T][M], G table [T][2J].
// Candidates kCoopStore.. of a pass (step sizes accepted in < 1 % of the iterations) are evaluated for
// their cost only; if one of them wins, the first lane of the group replays it into candidate 0.

// row length of a candidate in the line-search buffers: D padded to a multiple of 4 reals (32 bytes in FP64)
template <class Mdl>
__host__ __device__ constexpr int coop_dp() { return (Mdl::D + 3) & ~3; }

// closed-loop rollout for one step size from a stride-XS trajectory, stored to a stride-XS column;
// returns its cost (a separate function so that it gets its own register allocation)
template <class Mdl, class R, int RS>
__device__ __noinline__ double coop_rollout(const R* X, const R* K, const R* k, R alpha, ParamRef<R> prm, Bounds<R> ub,
                                            int T, R* Xnew, bool store) {
    RowPrefetchFeed<Mdl, R, RS> feed(X, K, k);
    // stores the padding behind each row as well: every 32-byte sector of a candidate row is written whole
    return rollout_closed_f<Mdl, R>(feed, alpha, prm, ub, T, RowView<R, RS, Mdl::D, coop_dp<Mdl>()>{Xnew, store});
}

// ---- two-phase rollout of a LONE trajectory (kinematic chains) ----------------------------------------------
// In the kinematic-chain models only the joint states and the actions are on the closed-loop rollout's recurrence
// (dx -> u -> s'); the end-effector outputs (sin / cos of the cumulative angles) and the cost terms - three quarters
// of a step's instructions, all of the barrier cost's logarithms - depend on finished rows only.  A warp that holds a
// single trajectory (the end of the tail, where the per-iteration latency is the critical path of the whole solve)
// therefore splits the line search: (1) every candidate lane runs the bare recurrence and stores joint states and
// actions; (2) all 32 lanes share the (candidate, time step) pairs: outputs of row t from the angles of row t-1, then
// the cost of row t; (3) every candidate lane sums its per-step costs in time order.  Each value is produced by the
// operation sequence of rollout_closed_f, so the result is bit-identical to the one-phase rollout.
template <class Mdl, class R, int RS>
__device__ __noinline__ void coop_recurrence(const R* X, const R* K, const R* k, R alpha, Bounds<R> ub, int T, R* Xnew) {
    constexpr int N = Mdl::N, M = Mdl::M, NS = Mdl::NS, J = Mdl::M;
    RowPrefetchFeed<Mdl, R, RS> feed(X, K, k);
    R s[NS], u[M];
    feed.begin(T - 1);
    CRL_UNROLL for (int j = 0; j < NS; ++j) s[j] = X[j];
    CRL_UNROLL for (int y = 0; y < N - NS; ++y) Xnew[NS + y] = X[NS + y];             // row 0 is copied whole
    CRL_UNROLL for (int a = 0; a < M; ++a) u[a] = X[N + a];
    for (int t = 0; t < T - 1; ++t) {
        R xo_c[NS], uo_c[M], K_c[M * NS], k_c[M];
        feed.get(t, xo_c, uo_c, K_c, k_c);
        R dx[NS];
        CRL_UNROLL for (int j = 0; j < NS; ++j) dx[j] = s[j] - xo_c[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            R acc = K_c[a * NS] * dx[0];
            CRL_UNROLL for (int j = 1; j < NS; ++j) acc = fmadd(K_c[a * NS + j], dx[j], acc);
            R du = acc + alpha * k_c[a];
            u[a] = clamp_(uo_c[a] + du, ub.lo, ub.hi);
        }
        R* row = Xnew + (size_t)t * RS;
        CRL_UNROLL for (int j = 0; j < NS; ++j) row[j] = s[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) row[N + a] = u[a];
        R sn[NS];
        CRL_UNROLL for (int j = 0; j < J; ++j) {                                     // step_trig's joint part
            sn[2 * j] = s[2 * j] + R(kH) * s[2 * j + 1];
            sn[2 * j + 1] = s[2 * j + 1] + R(kH) * u[j];
        }
        CRL_UNROLL for (int j = 0; j < NS; ++j) s[j] = sn[j];
    }
    if (T > 1) { CRL_UNROLL for (int a = 0; a < M; ++a) u[a] = R(0); }                // last action stays 0
    R* row = Xnew + (size_t)(T - 1) * RS;
    CRL_UNROLL for (int j = 0; j < NS; ++j) row[j] = s[j];
    CRL_UNROLL for (int a = 0; a < M; ++a) row[N + a] = u[a];
}

// phases 2 and 3 for the candidates `cand_mask` (bit c = candidate c of the buffer Xbuf = [T][GL][DP]); returns this
// lane's candidate cost (lanes c < GL of `group_lane0`'s group; 0 for the others).  Lcost: GL * T doubles of scratch.
template <class Mdl, class R, int GL, int DP>
__device__ __noinline__ double coop_outputs_and_costs(R* Xbuf, unsigned cand_mask, ParamRef<R> prm, int T, double* Lcost,
                                                      int lane, int my_cand) {
    constexpr int N = Mdl::N, D = Mdl::D, NS = Mdl::NS, RS = GL * DP;
    const int n_cand = __popc(cand_mask);
    const int n_items = n_cand * T;
    // kU (candidate, time step) pairs per lane and round, all their loads issued before any of them is evaluated
    // (the rows sit in L2 / DRAM and the stores below would otherwise fence the next pair's loads)
    constexpr int kU = CRL_SPLIT_UNROLL;
    for (int i0 = lane; i0 < n_items; i0 += 32 * kU) {
        R xu[kU][D], pang[kU][NS];
        R* row[kU];
        int tt[kU], cc[kU];
        bool on[kU];
        CRL_UNROLL for (int r = 0; r < kU; ++r) {
            const int i = i0 + 32 * r;
            on[r] = i < n_items;
            const int ii = on[r] ? i : lane;
            cc[r] = (int)__fns(cand_mask, 0, ii / T + 1);
            tt[r] = ii % T;
            row[r] = Xbuf + (size_t)tt[r] * RS + cc[r] * DP;
            CRL_UNROLL for (int j = 0; j < NS; ++j) xu[r][j] = __ldcg(row[r] + j);
            CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) xu[r][N + a] = __ldcg(row[r] + N + a);
            const R* prev = tt[r] > 0 ? row[r] - RS : row[r];
            CRL_UNROLL for (int j = 0; j < NS; ++j) pang[r][j] = __ldcg(prev + j);
            CRL_UNROLL for (int y = NS; y < N; ++y) xu[r][y] = __ldcg(row[r] + y);          // valid for t = 0 only
        }
        CRL_UNROLL for (int r = 0; r < kU; ++r) {
            if (!on[r]) continue;
            if (tt[r] > 0) {
                // outputs of row t = step_trig's output part at row t-1
                R pxu[D], ang[Mdl::NANG], sn[Mdl::NANG], cs[Mdl::NANG], xn[N];
                CRL_UNROLL for (int j = 0; j < NS; ++j) pxu[j] = pang[r][j];
                CRL_UNROLL for (int j = NS; j < D; ++j) pxu[j] = R(0);
                Mdl::template angles<R>(pxu, ang);
                sincos_batch<Mdl::NANG, R>(ang, sn, cs);
                Mdl::template step_trig<R>(pxu, sn, cs, xn);
                CRL_UNROLL for (int y = NS; y < N; ++y) { xu[r][y] = xn[y]; row[r][y] = xn[y]; }
            }
            Lcost[cc[r] * T + tt[r]] = step_cost<Mdl, R>(xu[r], prm, tt[r]);
        }
    }
    __syncwarp();
    double Jc = 0.0;
    if (my_cand >= 0 && ((cand_mask >> my_cand) & 1u)) {
        const double* L = Lcost + (size_t)my_cand * T;
        for (int t = 0; t < T; ++t) Jc = Jc + L[t];
    }
    return Jc;
}

// NT == 8: the pair-group pass (4 lanes per trajectory, kinematic chains); NT == 4: the column-group pass
// (8 lanes per trajectory); otherwise the element-per-lane pass
// the element-per-lane pass exists for the kinematic chains only
template <class Mdl> struct HasElementPass { static constexpr bool value = true; };
template <> struct HasElementPass<RoboticArm> { static constexpr bool value = false; };
template <class Base> struct HasElementPass<Barrier<Base>> { static constexpr bool value = false; };   // plain costs only

template <class Mdl, class R, int NT>
__device__ __noinline__ void coop_backward_call(const CoopTraj<R> (&tr)[NT], int T, R* S, int lane) {
    if constexpr (NT == 8) {
        static_assert(HasElementPass<Mdl>::value, "the pair-group pass (NT = 8) is written for the kinematic chains");
        pairgroup_backward<Mdl, R>(tr, T, S, lane);
    } else if constexpr (NT == ColGroup<Mdl>::NG) {
        if constexpr (IsChain<Mdl>::value) colgroup_backward<Mdl, R>(tr, T, S, lane);
        else colgroup_backward_arm<Mdl, R>(tr, T, S, lane);
    } else {
        static_assert(HasElementPass<Mdl>::value, "this model only has the column-group pass (NT = 4)");
        coop_backward<Mdl, R, NT>(tr, T, S, lane);
    }
}

// shared memory (reals) one warp of the engine needs
template <class Mdl, int NT>
__host__ __device__ constexpr int coop_smem_reals() {
    return NT == 8 ? (PairGroup<Mdl>::SMEM > CoopLayout<Mdl>::SZ ? PairGroup<Mdl>::SMEM : CoopLayout<Mdl>::SZ)
           : NT == ColGroup<Mdl>::NG
               ? (ColGroup<Mdl>::SMEM > CoopLayout<Mdl>::SZ ? ColGroup<Mdl>::SMEM : CoopLayout<Mdl>::SZ)
               : NT * CoopLayout<Mdl>::SZ;
}

// Trajectories a warp iterates together in the cooperative engine: 4 (column-group backward pass, 8
// line-search lanes per trajectory).  8 (pair-group pass, 4 line-search lanes; kinematic chains only) was
// measured too: as the whole solve it is 4 % faster in fixed-work runs (backward -31 %, but 1.8 line-search
// passes per iteration instead of 1), as the tail of the lane kernel 15 % slower - the tail is bound by the
// latency of an iteration, and an iteration of eight takes twice as long as an iteration of four.
#ifdef CRL_COOP_NT
template <class Mdl> constexpr int kCoopNTFor = CRL_COOP_NT;
#else
template <class Mdl> constexpr int kCoopNTFor = 4;
#endif

template <int NT>
struct CoopGeom {
    static constexpr int GL = (32 / NT) > 16 ? 16 : (32 / NT);   // line-search lanes (= candidates per pass) per trajectory
};
// reals between consecutive time steps of one candidate
template <class Mdl, int NT>
__host__ __device__ constexpr int coop_rs() { return CoopGeom<NT>::GL * coop_dp<Mdl>(); }
// candidates of a pass that are stored (the rest: cost only)
template <int NT>
__host__ __device__ constexpr int coop_store() { return CoopGeom<NT>::GL == 8 ? 6 : CoopGeom<NT>::GL; }
template <class Mdl, int NT>
__host__ __device__ constexpr long long coop_xscr_reals(int T) {
    return (long long)NT * 2 * T * coop_rs<Mdl, NT>();
}
template <class Mdl>
__host__ __device__ constexpr long long coop_gscr_reals_per_traj(int T) {
    // K, k, the table of variable Jacobian entries, and (two-phase line search) the per-step costs of 8 candidates
    return (long long)T * (Mdl::M * Mdl::NS + Mdl::M + Mdl::kCoopTab + 8);
}
template <class Mdl, int NT>
__host__ __device__ constexpr long long coop_scratch_reals(int T) {
    return coop_xscr_reals<Mdl, NT>(T) + NT * coop_gscr_reals_per_traj<Mdl>(T);
}

// A warp slot of the lane kernel: its tiles, followed (models with a cooperative path) by the scratch of
// the engine it becomes in the tail.
template <class Mdl>
__host__ __device__ constexpr long long slot_reals(int T) {
    return lane_tiles_reals<Mdl>(T) + (CoopSupport<Mdl>::value ? coop_scratch_reals<Mdl, kCoopNTFor<Mdl>>(T) : 0);
}

// Source of fresh trajectories: the batch itself.
template <class Mdl, class R>
struct BatchSource {
    const SolveArgs<R>& a;
    bool exhausted = false;
    __device__ explicit BatchSource(const SolveArgs<R>& a_) : a(a_) {}
    // Fills up to `want` of the empty slots (mask `empty`, bit q = slot q); slot q's trajectory is
    // rolled out into dst[q].  Returns the mask of the slots filled; traj/it/J are set for them.
    template <int NT>
    __device__ unsigned take(unsigned empty, R* const (&dst)[NT], long long (&traj)[NT], int (&it)[NT], double (&J)[NT],
                             int lane) {
        constexpr int RS = coop_rs<Mdl, NT>();
        if (exhausted || empty == 0u) return 0u;
        const int want = __popc(empty);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.ctl + kCtlNext, (unsigned long long)want);
        base = __shfl_sync(kFull, base, 0);
        if (base + want >= (unsigned long long)a.pb.B) exhausted = true;
        unsigned got = 0u;
        int n = 0;
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            if (!((empty >> q) & 1u)) continue;
            const long long idx = (long long)base + n;
            ++n;
            if (idx >= a.pb.B) continue;
            got |= 1u << q;
            traj[q] = idx;
            it[q] = 0;
            J[q] = a.J_init ? a.J_init[idx] : (double)INFINITY;
        }
        // initial open-loop rollouts, one lane per new trajectory
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            if (lane == q && ((got >> q) & 1u)) {
                const long long idx = traj[q];
                rollout_open_v<Mdl, R>(a.x0 + idx * Mdl::N, 1, a.u0 ? a.u0 + idx * a.u0_sb : nullptr, Mdl::M, a.pb.prm(idx),
                                       a.pb.ub, a.pb.T, RowView<R, RS>{dst[q]});
            }
        }
        __syncwarp();
        return got;
    }
    __device__ bool drained(int) const { return exhausted; }      // nothing more will ever come
    __device__ void finished(int, int) {}
};

// Source of the lane kernel's tail: the straggler queue.
template <class Mdl, class R>
struct StragglerSource {
    const SolveArgs<R>& a;
    unsigned total_warps;
    long long claim[8] = {-1, -1, -1, -1, -1, -1, -1, -1};            // queue index claimed for slot q, not yet published
    __device__ StragglerSource(const SolveArgs<R>& a_, unsigned tw) : a(a_), total_warps(tw) {}
    bool closed = false;
    unsigned long long q_tail = 0, q_head = 0;
    template <int NT>
    __device__ unsigned take(unsigned empty, R* const (&dst)[NT], long long (&traj)[NT], int (&it)[NT], double (&J)[NT],
                             int lane) {
        static_assert(NT <= 8, "claim slots");
        unsigned long long q_done = 0, q_fin = 0;
        if (lane == 0) {
            q_done = ld_volatile_u64(a.ctl + kCtlDone);
            __threadfence();                                   // the tail is final once every warp has reported
            q_tail = ld_volatile_u64(a.ctl + kCtlTail);
            q_head = ld_volatile_u64(a.ctl + kCtlHead);
            q_fin = ld_volatile_u64(a.ctl + kCtlFin);
        }
        q_done = __shfl_sync(kFull, q_done, 0);
        q_tail = __shfl_sync(kFull, q_tail, 0);
        q_head = __shfl_sync(kFull, q_head, 0);
        q_fin = __shfl_sync(kFull, q_fin, 0);
        closed = q_done >= (unsigned long long)total_warps;
        // fair share of the unfinished queue trajectories per warp that serves the queue
        const unsigned long long unfinished = q_tail > q_fin ? q_tail - q_fin : 0ull;
        const unsigned long long servers = q_done > 0 ? q_done : 1ull;
        unsigned long long share = (unfinished + servers - 1) / servers;
        const int n_star = share > (unsigned long long)NT ? NT : (share < 1 ? 1 : (int)share);
        int n_busy = NT - __popc(empty), n_clm = 0;
        CRL_UNROLL for (int q = 0; q < NT; ++q) n_clm += claim[q] >= 0 ? 1 : 0;
        int want = 0;
        {
            int free_slots = 0;
            CRL_UNROLL for (int q = 0; q < NT; ++q) free_slots += (((empty >> q) & 1u) && claim[q] < 0) ? 1 : 0;
            want = n_star - n_busy - n_clm;
            if (want > free_slots) want = free_slots;
            const unsigned long long avail = q_tail > q_head ? q_tail - q_head : 0ull;
            if ((unsigned long long)(want > 0 ? want : 0) > avail) want = (int)avail;
        }
        if (want > 0) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.ctl + kCtlHead, (unsigned long long)want);
            base = __shfl_sync(kFull, base, 0);
            int n = 0;
            CRL_UNROLL for (int q = 0; q < NT; ++q)
                if (((empty >> q) & 1u) && claim[q] < 0 && n < want) claim[q] = (long long)base + n++;
        }
        unsigned got = 0u;
        const unsigned long long qcap = (unsigned long long)total_warps * 32ull;
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            if (claim[q] < 0) continue;
            int rdy = 0;
            if (lane == 0 && (unsigned long long)claim[q] < qcap)
                rdy = *reinterpret_cast<const volatile int*>(&a.queue[claim[q]].ready);
            rdy = __shfl_sync(kFull, rdy, 0);
            if (rdy == 0) {
                if (closed && (unsigned long long)claim[q] >= q_tail) claim[q] = -1;   // the queue ended before this index
                continue;
            }
            __threadfence();
            const QueueEntry* qe = a.queue + claim[q];
            traj[q] = __ldcg(&qe->traj);
            const R* src = (const R*)(uintptr_t)__ldcg(reinterpret_cast<const unsigned long long*>(&qe->src));
            J[q] = __ldcg(&qe->J_last);
            it[q] = __ldcg(&qe->it);
            const int n = a.pb.T * Mdl::D;
            R* const dq = dst[q];
            for (int e0 = lane; e0 < n; e0 += 32 * kCopyBatch) {        // loads first (see warp_copy_elems)
                R v[kCopyBatch];
#pragma unroll
                for (int j = 0; j < kCopyBatch; ++j)
                    if (e0 + 32 * j < n) v[j] = __ldcg(src + (size_t)(e0 + 32 * j) * 32);
#pragma unroll
                for (int j = 0; j < kCopyBatch; ++j) {
                    const int e = e0 + 32 * j;
                    if (e < n) dq[(size_t)(e / Mdl::D) * coop_rs<Mdl, NT>() + e % Mdl::D] = v[j];
                }
            }
            claim[q] = -1;
            got |= 1u << q;
        }
        __syncwarp();
        return got;
    }
    __device__ bool drained(int) const {
        bool pending = false;
        CRL_UNROLL for (int q = 0; q < 8; ++q) pending = pending || claim[q] >= 0;
        return closed && !pending && q_head >= q_tail;
    }
    __device__ void finished(int n, int lane) {
        if (n > 0 && lane == 0) atomicAdd(a.ctl + kCtlFin, (unsigned long long)n);
    }
};

template <class Mdl, class R, int NT, class Source>
__device__ void coop_engine(const SolveArgs<R>& a, Source& src, R* xscr, R* gscr, R* S, int lane, PhaseTimer& tm) {
    constexpr int M = Mdl::M, D = Mdl::D, NS = Mdl::NS, GL = CoopGeom<NT>::GL;
    constexpr int DP = coop_dp<Mdl>(), RS = coop_rs<Mdl, NT>(), NST = coop_store<NT>();
    const int T = a.pb.T;
    const size_t TD = (size_t)T * D, TK = (size_t)T * M * NS, Tk = (size_t)T * M;
    const size_t TGS = (size_t)coop_gscr_reals_per_traj<Mdl>(T);
    const size_t BUF = (size_t)T * RS;                        // one line-search buffer [T][GL][DP]
    const int tries = (a.op.ls_method == CRL_LS_NONE) ? 1 : a.op.max_ls;
    const bool feas = a.op.ls_method == CRL_LS_FEASIBILITY;
    // slot state, identical in every lane: the current trajectory of slot q is candidate s_col[q] of its buffer s_buf[q]
    long long s_traj[NT];
    int s_it[NT], s_buf[NT], s_col[NT], s_stage[NT];
    double s_J[NT];
    CRL_UNROLL for (int q = 0; q < NT; ++q) { s_traj[q] = -1; s_it[q] = 0; s_buf[q] = 0; s_col[q] = 0; s_J[q] = 0.0; s_stage[q] = 0; }
    // line-search role of this lane: candidate c_ls of trajectory q_ls
    const bool ls_lane = lane < NT * GL;
    const int q_ls = ls_lane ? lane / GL : 0, c_ls = ls_lane ? lane % GL : 0;

    while (true) {
        // ---- refill empty slots (a new trajectory goes to column 0 of buffer 0 of its slot)
        unsigned empty = 0u;
        CRL_UNROLL for (int q = 0; q < NT; ++q) empty |= (s_traj[q] < 0 ? 1u : 0u) << q;
        if (empty != 0u) {
            R* dst[NT];
            CRL_UNROLL for (int q = 0; q < NT; ++q) dst[q] = xscr + (size_t)(2 * q) * BUF;
            const unsigned got = src.template take<NT>(empty, dst, s_traj, s_it, s_J, lane);
            CRL_UNROLL for (int q = 0; q < NT; ++q)
                if ((got >> q) & 1u) {
                    s_buf[q] = 0;
                    s_col[q] = 0;
                    s_stage[q] = 0;
                    if constexpr (Mdl::kBarrier) {           // stage schedules: the source's iteration word carries the stage
                        s_stage[q] = s_it[q] >> kStageShift;
                        s_it[q] &= kStageItMask;
                    }
                }
            empty &= ~got;
        }
        tm.lap(PH_COOP_WAIT);
        if (empty == (1u << NT) - 1u) {
            if (src.drained(lane)) return;
            __nanosleep(1000);
            continue;
        }
        // ---- backward passes of all slots, stepped together
        CoopTraj<R> tr[NT];
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            const bool on = s_traj[q] >= 0;
            R* g = gscr + (size_t)q * TGS;
            tr[q].X = xscr + (size_t)(2 * q + s_buf[q]) * BUF + s_col[q] * DP;
            tr[q].rs = RS;
            tr[q].prm = a.pb.prm(on ? s_traj[q] : 0);
            if constexpr (Mdl::kBarrier) tr[q].prm.ptr += s_stage[q] * Mdl::P;
            tr[q].K = g;
            tr[q].k = g + TK;
            tr[q].G = g + TK + Tk;
            tr[q].on = on;
            tr[q].first = s_it[q] == 0;
        }
        {
            // a warp that holds a single trajectory (the end of the tail) steps it alone: lowest latency
            int n_on = 0, q_on = 0;
            CRL_UNROLL for (int q = 0; q < NT; ++q)
                if (tr[q].on) { ++n_on; q_on = q; }
            bool alone = false;
            if constexpr (NT > 1 && HasElementPass<Mdl>::value) {
                if (n_on == 1) {
                    CoopTraj<R> one[1];
                    one[0] = tr[0];
                    CRL_UNROLL for (int q = 1; q < NT; ++q)
                        if (q == q_on) one[0] = tr[q];
                    coop_backward_call<Mdl, R, 1>(one, T, S, lane);
                    alone = true;
                }
            }
            if (!alone) coop_backward_call<Mdl, R, NT>(tr, T, S, lane);
        }
        __syncwarp();
        tm.lap(PH_COOP_BACKWARD);
        // ---- line search: GL step sizes per trajectory and pass, candidates side by side in the slot's other buffer
        // trajectories this warp holds (warp-uniform): the two-phase line search is chosen by it
        int n_live = 0;
        CRL_UNROLL for (int q = 0; q < NT; ++q) n_live += s_traj[q] >= 0 ? 1 : 0;
        bool accepted[NT];
        int next[NT], taken[NT], newcol[NT];
        double J_new[NT];
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            accepted[q] = false; next[q] = 0; taken[q] = -1; newcol[q] = 0; J_new[q] = s_J[q];
        }
        while (true) {
            bool any = false;
            CRL_UNROLL for (int q = 0; q < NT; ++q) any = any || (s_traj[q] >= 0 && !accepted[q] && next[q] < tries);
            if (!any) break;
            // this lane's candidate
            int my_next = 0, my_buf = 0;
            bool my_pend = false;
            double my_Jl = 0.0;
            CoopTraj<R> mt = tr[0];
            CRL_UNROLL for (int q = 0; q < NT; ++q)
                if (q == q_ls) {
                    my_next = next[q]; my_buf = s_buf[q]; my_Jl = s_J[q]; mt = tr[q];
                    my_pend = s_traj[q] >= 0 && !accepted[q] && next[q] < tries;
                }
            const int idx = my_next + c_ls;
            double Jt = 0.0;
            bool ok = false;
            bool split = false;
            if constexpr (kSplitRollout && Mdl::kKinematic && Mdl::kBarrier && sizeof(R) == 8 && GL == 8)
                split = n_live <= kSplitMaxLive;
            if (split) {
                // two-phase rollout (see coop_recurrence): every lane of the warp takes part in phase 2
                const bool mine = ls_lane && my_pend && idx < tries;
                if (mine) {
                    double al = 1.0;
                    for (int i = 0; i < idx; ++i) al = al * a.op.gamma;
                    coop_recurrence<Mdl, R, RS>(mt.X, mt.K, mt.k, (R)al, a.pb.ub, T,
                                                xscr + (size_t)(2 * q_ls + (my_buf ^ 1)) * BUF + c_ls * DP);
                }
                const unsigned mm = __ballot_sync(kFull, mine);
                __syncwarp();
                CRL_UNROLL for (int q = 0; q < NT; ++q) {
                    const unsigned cm = (mm >> (q * GL)) & ((1u << GL) - 1u);
                    if (cm == 0u) continue;                                          // warp-uniform
                    double* const Lcost = reinterpret_cast<double*>(gscr + (size_t)q * TGS + TK + Tk + (size_t)T * Mdl::kCoopTab);
                    const double Jq = coop_outputs_and_costs<Mdl, R, GL, DP>(xscr + (size_t)(2 * q + (s_buf[q] ^ 1)) * BUF, cm,
                                                                             tr[q].prm, T, Lcost, lane,
                                                                             (mine && q_ls == q) ? c_ls : -1);
                    if (mine && q_ls == q) Jt = Jq;
                }
                __syncwarp();
                if (mine) ok = (a.op.ls_method == CRL_LS_NONE) || ls_accept(Jt, my_Jl, feas);
            } else if (ls_lane && my_pend && idx < tries) {
                double al = 1.0;
                for (int i = 0; i < idx; ++i) al = al * a.op.gamma;
                Jt = coop_rollout<Mdl, R, RS>(mt.X, mt.K, mt.k, (R)al, mt.prm, a.pb.ub, T,
                                              xscr + (size_t)(2 * q_ls + (my_buf ^ 1)) * BUF + c_ls * DP, c_ls < NST);
                ok = (a.op.ls_method == CRL_LS_NONE) || ls_accept(Jt, my_Jl, feas);
            }
            const unsigned okm = __ballot_sync(kFull, ok);
            CRL_UNROLL for (int q = 0; q < NT; ++q) {
                const bool pend = s_traj[q] >= 0 && !accepted[q] && next[q] < tries;
                const unsigned gm = (okm >> (q * GL)) & ((1u << GL) - 1u);
                const int w = gm != 0u ? __ffs(gm) - 1 : 0;      // lowest lane of the group = largest step size
                const double Jw = __shfl_sync(kFull, Jt, q * GL + w);
                if (pend) {
                    if (gm != 0u) {
                        accepted[q] = true;
                        taken[q] = next[q] + w;
                        J_new[q] = Jw;
                        newcol[q] = w;
                    } else {
                        next[q] += GL;
                    }
                }
            }
        }
        if constexpr (NST < GL) {
            // a winner among the cost-only candidates: the first lane of its group replays it into candidate 0
            bool need = false;
            CRL_UNROLL for (int q = 0; q < NT; ++q) need = need || (accepted[q] && newcol[q] >= NST);
            if (need) {
                int my_taken = -1, my_buf = 0;
                CoopTraj<R> mt = tr[0];
                CRL_UNROLL for (int q = 0; q < NT; ++q)
                    if (q == q_ls && accepted[q] && newcol[q] >= NST) { my_taken = taken[q]; my_buf = s_buf[q]; mt = tr[q]; }
                if (ls_lane && c_ls == 0 && my_taken >= 0) {
                    double al = 1.0;
                    for (int i = 0; i < my_taken; ++i) al = al * a.op.gamma;
                    coop_rollout<Mdl, R, RS>(mt.X, mt.K, mt.k, (R)al, mt.prm, a.pb.ub, T,
                                             xscr + (size_t)(2 * q_ls + (my_buf ^ 1)) * BUF, true);
                }
                CRL_UNROLL for (int q = 0; q < NT; ++q)
                    if (accepted[q] && newcol[q] >= NST) newcol[q] = 0;
            }
        }
        __syncwarp();
        tm.lap(PH_COOP_LINESEARCH);
        // ---- stop test, bookkeeping, results
        int n_fin = 0;
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            if (s_traj[q] < 0) continue;
            const bool stop = stop_test(J_new[q], s_J[q], a.op.criterion, a.op.stop_method == CRL_STOP_RELATIVE);
            s_J[q] = J_new[q];
            if (accepted[q]) { s_buf[q] ^= 1; s_col[q] = newcol[q]; }
            if (lane == 0) {
                if (a.J_trace) a.J_trace[s_traj[q] * a.op.max_iter + s_it[q]] = J_new[q];
                if (a.alpha_trace) a.alpha_trace[s_traj[q] * a.op.max_iter + s_it[q]] = (int8_t)taken[q];
            }
            ++s_it[q];
            if ((stop && a.op.check_stop) || s_it[q] >= a.op.max_iter) {
                int total = s_it[q];
                if constexpr (Mdl::kBarrier) {
                    if (a.n_stage > 1) {
                        int go_on = 0;
                        if (lane == 0) go_on = stage_advance<R>(a, s_traj[q], s_stage[q], s_it[q], &total) ? 1 : 0;
                        if (__shfl_sync(kFull, go_on, 0) != 0) {
                            ++s_stage[q];
                            s_it[q] = 0;
                            if (stop && a.op.check_stop) s_J[q] = (double)INFINITY;
                            continue;
                        }
                    }
                }
                if (lane == 0) {
                    a.J_out[s_traj[q]] = s_J[q];
                    a.iters_out[s_traj[q]] = total;
                    a.conv_out[s_traj[q]] = (stop && a.op.check_stop) ? 1 : 0;
                }
                if (a.X_out) {
                    const R* from = xscr + (size_t)(2 * q + s_buf[q]) * BUF + s_col[q] * DP;
                    R* to = a.X_out + s_traj[q] * TD;
                    warp_copy_elems<R>(to, T * D, lane, [from](int e) { return from + (size_t)(e / D) * RS + e % D; });
                    __syncwarp();                               // the buffers are reused by the slot's next trajectory
                }
                s_traj[q] = -1;
                ++n_fin;
            }
        }
        src.finished(n_fin, lane);
        tm.lap(PH_COOP_OUTPUT);
    }
}

// Publishes the live trajectories of a warp that leaves lane mode on the straggler queue (mask
// `live`; each lane's current trajectory is its column `my_col` of a tile) and reports the warp.
template <class R>
__device__ void publish_stragglers(const SolveArgs<R>& a, bool live, long long traj, int it, double J_last, const R* my_col,
                                   int lane) {
    const unsigned live_mask = __ballot_sync(kFull, live);
    if (live_mask != 0u) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.ctl + kCtlTail, (unsigned long long)__popc(live_mask));
        base = __shfl_sync(kFull, base, 0);
        if (live) {
            QueueEntry* e = a.queue + base + __popc(live_mask & ((1u << lane) - 1u));
            e->traj = traj;
            e->src = my_col;
            e->J_last = J_last;
            e->it = it;
            __threadfence();
            *reinterpret_cast<volatile int*>(&e->ready) = 1;
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAdd(a.ctl + kCtlDone, 1ull);
    }
}


// The whole solve by the cooperative engine: persistent grid, WARPS warps per CTA, every warp
// takes NT trajectories of the batch at a time.
template <class Mdl, class R, int WARPS, int MINB, int NT>
__global__ void __launch_bounds__(WARPS * 32, MINB) coop_solve_kernel(SolveArgs<R> a) {
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    R* const S = reinterpret_cast<R*>(coop_smem) + (size_t)(threadIdx.x >> 5) * coop_smem_reals<Mdl, NT>();
    R* const slot = a.ws + warp * a.slot_elems;
    R* const xscr = slot;
    R* const gscr = slot + coop_xscr_reals<Mdl, NT>(a.pb.T);
    PhaseTimer tm;
    BatchSource<Mdl, R> src(a);
    coop_engine<Mdl, R, NT>(a, src, xscr, gscr, S, lane, tm);
    tm.flush(a.ctl, lane);
}

// WARPS = warps per CTA, MINB = resident CTAs per SM to compile for.  SYNC = true phase-locks the
// warps of a CTA (one barrier at the top of an iteration and one after the backward pass): the
// unrolled backward / line-search loops are ~16 KB of code each, and warps of one SM running
// different loops at the same time overflow the SM's instruction cache (measured: 75 % hit rate,
// GPC instruction cache at 81 % of its request throughput, 20-34 % "no instruction" stalls).
template <class Mdl, class R, int WARPS, int MINB, bool SYNC>
__global__ void __launch_bounds__(WARPS * 32, MINB) solve_kernel(SolveArgs<R> a) {
    constexpr int N = Mdl::N, M = Mdl::M, D = Mdl::D, NS = Mdl::NS;
    constexpr bool kCoop = CoopSupport<Mdl>::value;
    const int lane = threadIdx.x & 31;
    const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int T = a.pb.T;
    R* wsb = a.ws + slot * a.slot_elems + lane;
    const View<R, D, 32> Xbuf0{wsb}, Xbuf1{wsb + (size_t)T * D * 32};
    const View<R, M * NS, 32> Kv{wsb + (size_t)2 * T * D * 32};
    const View<R, M, 32> kv{wsb + (size_t)2 * T * D * 32 + (size_t)T * M * NS * 32};

    // per-warp staging area in dynamic shared memory
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SL = StageLayout<Mdl, R>;
    unsigned char* const my_sm = smem_raw + (threadIdx.x >> 5) * SL::kWarpBytes;
    R* const stage_sm = reinterpret_cast<R*>(my_sm);
    unsigned long long* const stage_bar = reinterpret_cast<unsigned long long*>(my_sm + 2 * SL::STAGE * sizeof(R));
    if (lane == 0) {
        mbar_init(stage_bar);
        mbar_init(stage_bar + 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    int cur = 0;                 // warp-uniform: which X buffer holds the current trajectories
    bool exhausted = false;      // warp-uniform: the global queue is empty
    long long traj = -1;         // this lane's trajectory, -1 = idle
    int it = 0;
    int pred = 0;                // step index this lane's trajectory accepted last (line-search guess)
    double J_last = 0.0;
    ParamRef<R> prm{a.pb.params, 0};
    PhaseTimer tm;

    while (true) {
        // ---- refill idle lanes from the global queue
        const unsigned need = __ballot_sync(kFull, traj < 0);
        if (need != 0u && !exhausted) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.ctl + kCtlNext, (unsigned long long)__popc(need));
            base = __shfl_sync(kFull, base, 0);
            if (base + __popc(need) >= (unsigned long long)a.pb.B) exhausted = true;
            if (traj < 0) {
                const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (idx < a.pb.B) {
                    traj = idx;
                    it = 0;
                    pred = 0;
                    J_last = a.J_init ? a.J_init[idx] : (double)INFINITY;
                    prm = a.pb.prm(idx);
                    const View<R, D, 32> Xc = cur ? Xbuf1 : Xbuf0;
                    rollout_open<Mdl, R, 32>(a.x0 + idx * N, 1, a.u0 ? a.u0 + idx * a.u0_sb : nullptr, M, prm, a.pb.ub, T,
                                             Xc);
                }
            }
        }
        const bool active = traj >= 0;
        if (kCoop && a.coop_threshold >= 0 && !exhausted) {
            // a warp with no idle lane does not visit the queue: look at the counter instead
            unsigned long long nxt = 0;
            if (lane == 0) nxt = ld_volatile_u64(a.ctl + kCtlNext);
            nxt = __shfl_sync(kFull, nxt, 0);
            if (nxt >= (unsigned long long)a.pb.B) exhausted = true;
        }
        if (SYNC) {
            const int live = __syncthreads_count(active);   // CTA-uniform; also aligns the warps
            if (live == 0) break;
            if (kCoop && a.coop_threshold >= 0) {
                // queue empty and few trajectories left in this CTA: hand them to the finishers
                if (__syncthreads_and(exhausted) && live <= a.coop_threshold * WARPS) break;
            }
        } else {
            const unsigned live = __ballot_sync(kFull, active);
            if (live == 0u) break;
            if (kCoop && a.coop_threshold >= 0 && exhausted && __popc(live) <= a.coop_threshold) break;
        }

        const View<R, D, 32> Xc = cur ? Xbuf1 : Xbuf0;
        const View<R, D, 32> Xn = cur ? Xbuf0 : Xbuf1;
        tm.lap(PH_REFILL);

        // ---- backward pass: K, k for every active lane
        const unsigned act_mask = __ballot_sync(kFull, active);
        if (active) {
            TileRowFeed<Mdl, R> rows(act_mask, Xc.base - lane, stage_sm, stage_bar);
            backward_fused_f<Mdl, R, NS>(rows, prm, T, Kv, kv, it == 0);
        }
#ifndef CRL_NO_MID_BARRIER
        if (SYNC) __syncthreads();
#endif
        tm.lap(PH_BACKWARD);

        // ---- line search: first improving alpha = 1, g, g^2, ...
        bool accepted = false;
        double J_new = J_last;
        int taken = -1;
        const int tries = (a.op.ls_method == CRL_LS_NONE) ? 1 : a.op.max_ls;
        const bool feas = a.op.ls_method == CRL_LS_FEASIBILITY;
        const unsigned act = __ballot_sync(kFull, active);
        if (!kCoop && exhausted && __popc(act) <= kCoopMax) {
            // Straggler mode (queue empty, few trajectories left in this warp): the line search of
            // one trajectory at a time is spread over the idle lanes, one step size per lane, so a
            // whole line search costs one rollout of latency instead of up to max_line_search.
            // Candidate arithmetic is rollout_closed's, hence results do not depend on the mode.
            __syncwarp();                                   // K, k of the owners are read by other lanes below
            const unsigned idle = ~act;
            const int n_idle = __popc(idle);
            const int my_rank = __popc(idle & ((1u << lane) - 1u));
            const bool is_idle = (idle >> lane) & 1u;
            R* const ws0 = wsb - lane;                      // column 0 of this warp's tiles
            for (unsigned m = act; m != 0u; m &= m - 1u) {
                const int owner = __ffs(m) - 1;
                const double Jl_o = __shfl_sync(kFull, J_last, owner);
                const unsigned long long pp = __shfl_sync(kFull, (unsigned long long)(uintptr_t)prm.ptr, owner);
                const ParamRef<R> prm_o{(const R*)(uintptr_t)pp, __shfl_sync(kFull, prm.stride_t, owner)};
                const View<R, D, 32> Xc_o{ws0 + (cur ? (size_t)T * D * 32 : 0) + owner};
                const View<R, M * NS, 32> Kv_o{ws0 + (size_t)2 * T * D * 32 + owner};
                const View<R, M, 32> kv_o{ws0 + (size_t)2 * T * D * 32 + (size_t)T * M * NS * 32 + owner};
                int win_lane = -1, win_idx = -1;
                double J_win = Jl_o;
                for (int a0 = 0; a0 < tries && win_lane < 0; a0 += n_idle) {
                    const int idx = a0 + my_rank;
                    const bool work = is_idle && idx < tries;
                    double Jt = 0.0;
                    bool ok = false;
                    if (work) {
                        double al = 1.0;
                        for (int i = 0; i < idx; ++i) al = al * a.op.gamma;
                        Jt = rollout_closed<Mdl, R, NS, 32, 32, 32>(Xc_o, Kv_o, kv_o, (R)al, prm_o, a.pb.ub, T, Xn);
                        ok = (a.op.ls_method == CRL_LS_NONE) || ls_accept(Jt, Jl_o, feas);
                    }
                    const unsigned okm = __ballot_sync(kFull, ok);
                    if (okm != 0u) {
                        win_lane = __ffs(okm) - 1;               // lowest idle lane = lowest step index
                        win_idx = a0 + __popc(idle & ((1u << win_lane) - 1u));
                        J_win = __shfl_sync(kFull, Jt, win_lane);
                    }
                }
                __syncwarp();
                if (win_lane >= 0) {
                    // move the winning candidate from the helper's column to the owner's column of X_new
                    R* const xn0 = Xn.base - lane;
                    for (int e = lane; e < T * D; e += 32) xn0[(size_t)e * 32 + owner] = xn0[(size_t)e * 32 + win_lane];
                }
                __syncwarp();
                if (lane == owner && win_lane >= 0) {
                    accepted = true;
                    J_new = J_win;
                    taken = win_idx;
                }
            }
        } else {
            // Candidates are evaluated kLsChunk at a time in one pass over X, K, k (rollout_multi);
            // each lane keeps the candidate it expects to win (`pred`, the index accepted last
            // time) and re-materialises the winner with one more rollout only if it guessed wrong.
            double alpha = 1.0, alpha_taken = 1.0;
            bool replay = false;
            int a0 = 0;
            while (a0 < tries) {
                const bool pending = active && !accepted;
                const unsigned pend_mask = __ballot_sync(kFull, pending);
                if (pend_mask == 0u) break;
                const int n_pend = __popc(pend_mask);
                if (kHelpLs && a0 > 0 && n_pend <= kHelpMax) {
                    // Few lanes are still searching (after the first pass 80 % have their step): instead of
                    // another multi-alpha pass of the whole warp, the 32 lanes share the pending
                    // trajectories - G = 32 / n_pend lanes per trajectory, one step size each, costs only,
                    // reading the owner's column of the staged tile.  The accepted step is stored by the
                    // owner in the replay pass below.  Same arithmetic per candidate as everywhere else.
                    const int G = 32 / n_pend;
                    const int r = lane / G, j = lane - r * G;
                    const bool has = r < n_pend;
                    const int owner = has ? (int)__fns(pend_mask, 0, r + 1) : 0;
                    const int idx = a0 + j;
                    const bool work = has && idx < tries;
                    const double Jl_o = __shfl_sync(kFull, J_last, owner);
                    const unsigned long long pp = __shfl_sync(kFull, (unsigned long long)(uintptr_t)prm.ptr, owner);
                    const ParamRef<R> prm_o{(const R*)(uintptr_t)pp, __shfl_sync(kFull, prm.stride_t, owner)};
                    const unsigned work_mask = __ballot_sync(kFull, work);
                    double Jt = 0.0;
                    bool ok = false;
                    if (work) {
                        double al = alpha;
                        for (int i = 0; i < j; ++i) al = al * a.op.gamma;
                        TileFeed<Mdl, R> feed(work_mask, Xc.base - lane, Kv.base - lane, kv.base - lane, stage_sm, stage_bar, owner);
                        const R alr[1] = {(R)al};
                        double Jc[1];
                        rollout_multi_f<Mdl, R, 1, TileFeed<Mdl, R>, 32, false>(feed, alr, prm_o, a.pb.ub, T, 0, Xn, Jc);
                        Jt = Jc[0];
                        ok = ls_accept(Jt, Jl_o, feas);
                    }
                    const unsigned okm = __ballot_sync(kFull, ok);
                    int w = -1, src_lane = lane;
                    if (pending) {
                        const int rme = __popc(pend_mask & ((1u << lane) - 1u));
                        const unsigned gm = (okm >> (rme * G)) & (G >= 32 ? kFull : ((1u << G) - 1u));
                        if (gm != 0u) {
                            w = __ffs(gm) - 1;                  // lowest lane of the group = largest step size
                            src_lane = rme * G + w;
                        }
                    }
                    const double Jw = __shfl_sync(kFull, Jt, src_lane);
                    if (pending && w >= 0) {
                        accepted = true;
                        J_new = Jw;
                        taken = a0 + w;
                        double al = alpha;
                        for (int i = 0; i < w; ++i) al = al * a.op.gamma;
                        alpha_taken = al;
                        replay = true;
                    }
                    for (int i = 0; i < G; ++i) alpha = alpha * a.op.gamma;
                    a0 += G;
                    continue;
                }
                // a warp whose pending lanes all expect alpha = 1 to be accepted tries it alone first
                const bool wide = (tries - a0 >= 2) && (a0 > 0 || __any_sync(kFull, pending && pred > 0));
                if (wide) {
                    R al[kLsChunk];
                    double ald[kLsChunk];
    #pragma unroll
                    for (int c = 0; c < kLsChunk; ++c) {
                        ald[c] = alpha;
                        al[c] = (R)alpha;
                        alpha = alpha * a.op.gamma;
                    }
                    if (pending) {
                        int keep = pred - a0;
                        keep = keep < 0 ? 0 : (keep > kLsChunk - 1 ? kLsChunk - 1 : keep);
                        double Jc[kLsChunk];
                        TileFeed<Mdl, R> feed(pend_mask, Xc.base - lane, Kv.base - lane, kv.base - lane, stage_sm, stage_bar);
                        rollout_multi_f<Mdl, R, kLsChunk, TileFeed<Mdl, R>, 32, true>(feed, al, prm, a.pb.ub, T, keep, Xn, Jc);
    #pragma unroll
                        for (int c = 0; c < kLsChunk; ++c) {
                            if (!accepted && a0 + c < tries && ls_accept(Jc[c], J_last, feas)) {
                                accepted = true;
                                J_new = Jc[c];
                                taken = a0 + c;
                                alpha_taken = ald[c];
                                replay = (c != keep);
                            }
                        }
                    }
                    a0 += kLsChunk;
                } else {
                    if (pending) {
                        TileFeed<Mdl, R> feed(pend_mask, Xc.base - lane, Kv.base - lane, kv.base - lane, stage_sm, stage_bar);
                        const double Jt = rollout_closed_f<Mdl, R>(feed, (R)alpha, prm, a.pb.ub, T, Xn);
                        if (a.op.ls_method == CRL_LS_NONE || ls_accept(Jt, J_last, feas)) {
                            accepted = true;
                            J_new = Jt;
                            taken = a0;
                        }
                    }
                    alpha = alpha * a.op.gamma;
                    a0 += 1;
                }
            }
            const unsigned replay_mask = __ballot_sync(kFull, replay);
            if (replay_mask != 0u) {
                if (replay) {
                    TileFeed<Mdl, R> feed(replay_mask, Xc.base - lane, Kv.base - lane, kv.base - lane, stage_sm, stage_bar);
                    rollout_closed_f<Mdl, R>(feed, (R)alpha_taken, prm, a.pb.ub, T, Xn);
                }
            }
        }
        if (accepted) pred = taken;
        tm.lap(PH_LINESEARCH);

        // ---- stop test, bookkeeping, retirement
        bool finished = false;
        if (active) {
            const bool stop = stop_test(J_new, J_last, a.op.criterion, a.op.stop_method == CRL_STOP_RELATIVE);
            J_last = J_new;
            if (a.J_trace) a.J_trace[traj * a.op.max_iter + it] = J_new;
            if (a.alpha_trace) a.alpha_trace[traj * a.op.max_iter + it] = (int8_t)taken;
            ++it;
            finished = (stop && a.op.check_stop) || it >= a.op.max_iter;
            int total = it;
            if constexpr (Mdl::kBarrier) {
                if (finished && a.n_stage > 1) {
                    const int stage = (int)((prm.ptr - a.pb.prm(traj).ptr) / Mdl::P);
                    if (stage_advance<R>(a, traj, stage, it, &total)) {
                        prm.ptr += Mdl::P;
                        it = 0;
                        if (stop && a.op.check_stop) J_last = (double)INFINITY;
                        finished = false;
                    }
                }
            }
            if (finished) {
                a.J_out[traj] = J_last;
                a.iters_out[traj] = total;
                a.conv_out[traj] = (stop && a.op.check_stop) ? 1 : 0;
            }
        }
        if (a.X_out) {
            // Final trajectories leave the tile through the whole warp: element e of a finished lane's
            // column goes to X_out[traj][e], 32 consecutive elements per store (a lane writing its own
            // 8.8 KB row by row holds the warp - and, phase-locked, the CTA - for tens of microseconds).
            for (unsigned m = __ballot_sync(kFull, finished); m != 0u; m &= m - 1u) {
                const int owner = __ffs(m) - 1;
                const long long tj = __shfl_sync(kFull, traj, owner);
                const bool acc_o = __shfl_sync(kFull, (int)accepted, owner) != 0;
                const R* col = (acc_o ? Xn.base : Xc.base) - lane + owner;
                R* dst = a.X_out + tj * T * D;
                warp_copy_elems<R>(dst, T * D, lane, [col](int e) { return col + (size_t)e * 32; });
            }
        }
        // lanes that go on without an accepted step carry their trajectory into the other
        // buffer so that `cur` can flip for the whole warp
        const bool carry = active && !finished && !accepted;
        if (__any_sync(kFull, carry)) {
            if (carry) {
#pragma unroll 8
                for (int t = 0; t < T; ++t) {
                    const R* src = Xc.row(t);
                    R* dst = Xn.row(t);
#pragma unroll
                    for (int i = 0; i < D; ++i) Xn.st(dst, i, Xc.ld(src, i));
                }
            }
        }
        if (finished) traj = -1;
        cur ^= 1;
        tm.lap(PH_RETIRE);
    }

    if constexpr (kCoop) {
      if (a.coop_threshold >= 0) {
        // Every warp reports to the straggler queue exactly once (with its live trajectories, which
        // sit in tile `cur`) and then serves it; the other X tile and the gain tiles are its scratch.
        const View<R, D, 32> Xc = cur ? Xbuf1 : Xbuf0;
        R* const xfree = a.ws + slot * a.slot_elems + lane_tiles_reals<Mdl>(T);   // engine scratch behind the tiles
        R* const gfree = xfree + coop_xscr_reals<Mdl, kCoopNTFor<Mdl>>(T);
        int it_word = it;
        if constexpr (Mdl::kBarrier) {
            if (traj >= 0 && a.n_stage > 1) it_word |= (int)((prm.ptr - a.pb.prm(traj).ptr) / Mdl::P) << kStageShift;
        }
        publish_stragglers<R>(a, traj >= 0, traj, it_word, J_last, Xc.base, lane);
        tm.lap(PH_RETIRE);
        StragglerSource<Mdl, R> src(a, gridDim.x * WARPS);
        coop_engine<Mdl, R, kCoopNTFor<Mdl>>(a, src, xfree, gfree, stage_sm, lane, tm);
      }
    }
    tm.flush(a.ctl, lane);
}

// ------------------------------------------------------------------------------ host side
static int cuda_status(cudaError_t e) { return e == cudaSuccess ? CRL_OK : CRL_ERR_CUDA - (int)e; }

template <class R>
static ProblemDev<R> to_dev(const crl_problem_t* p) {
    ProblemDev<R> d;
    d.params = (const R*)p->params;
    d.params_stride_b = p->params_stride_b;
    d.params_stride_t = p->params_stride_t;
    d.B = p->batch;
    d.T = p->horizon;
    d.ub.lo = (R)p->u_lo;
    d.ub.hi = (R)p->u_hi;
    return d;
}

static OptsDev to_dev(const crl_solver_opts_t* o) {
    OptsDev d;
    d.max_iter = o->max_iter;
    d.check_stop = o->is_check_stop;
    d.max_ls = o->max_line_search;
    d.ls_method = o->line_search_method;
    d.stop_method = o->stopping_method;
    d.criterion = o->stopping_criterion;
    d.gamma = o->gamma;
    return d;
}

static int check_problem(const crl_problem_t* p) {
    if (!p) return CRL_ERR_BAD_ARGUMENT;
    if (gen_is_model(p->model)) {                    // generated models: FP64, parameter rows of the model's own width
        int pw = 0;
        gen_dims(p->model, nullptr, nullptr, &pw);
        if (p->dtype != CRL_F64) return CRL_ERR_UNSUPPORTED_MODEL;
        if (p->batch < 0 || p->horizon < 1 || (p->batch > 0 && !p->params)) return CRL_ERR_BAD_ARGUMENT;
        if ((p->params_stride_t != 0 && p->params_stride_t != pw) || p->params_stride_b < 0) return CRL_ERR_BAD_ARGUMENT;
        if (p->bounds_host) {
            int n = 0, m = 0;
            double lo[8], hi[8];
            gen_dims(p->model, &n, &m, nullptr);
            gen_bounds(p->model, lo, hi);
            for (int i = 0; i < n + m; ++i) {
                const double bl = p->bounds_host[2 * i], bh = p->bounds_host[2 * i + 1];
                if (!(bl <= bh)) return CRL_ERR_BAD_ARGUMENT;
                if (gen_is_barrier(p->model) && (bl != lo[i] || bh != hi[i])) return CRL_ERR_UNSUPPORTED_MODEL;
            }
        }
        return CRL_OK;
    }
    if (p->model < CRL_TWO_LINK || p->model > CRL_ROBOTIC_ARM_BARRIER) return CRL_ERR_UNSUPPORTED_MODEL;
    if (p->dtype != CRL_F64 && p->dtype != CRL_F32) return CRL_ERR_BAD_ARGUMENT;
    if (p->model >= CRL_TWO_LINK_BARRIER && p->dtype != CRL_F64) return CRL_ERR_UNSUPPORTED_MODEL;   // barrier models: FP64 only
    if (p->batch < 0 || p->horizon < 1) return CRL_ERR_BAD_ARGUMENT;
    if (p->batch > 0 && !p->params) return CRL_ERR_BAD_ARGUMENT;
    if (p->params_stride_t != 0 && p->params_stride_t != (p->model >= CRL_TWO_LINK_BARRIER ? 4 : 2)) return CRL_ERR_BAD_ARGUMENT;
    if (p->params_stride_b < 0) return CRL_ERR_BAD_ARGUMENT;
    if (!(p->u_lo <= p->u_hi)) return CRL_ERR_BAD_ARGUMENT;
    return CRL_OK;
}

static int check_opts(const crl_solver_opts_t* o) {
    if (!o) return CRL_ERR_BAD_ARGUMENT;
    if (o->max_iter < 0 || o->max_line_search < 0 || o->reserved2 != 0) return CRL_ERR_BAD_ARGUMENT;
    if (o->reserved != 0 && o->reserved != 2 && o->reserved != 12 && o->reserved != 13 && (o->reserved < 20 || o->reserved > 23))
        return CRL_ERR_BAD_ARGUMENT;             // a configuration that is not compiled in is rejected when the grid is sized
    if (o->line_search_method < CRL_LS_VANILLA || o->line_search_method > CRL_LS_NONE) return CRL_ERR_BAD_ARGUMENT;
    if (o->stopping_method < CRL_STOP_RELATIVE || o->stopping_method > CRL_STOP_VANILLA) return CRL_ERR_BAD_ARGUMENT;
    return CRL_OK;
}

// Makes the device that owns `ptr` current for the duration of a call (the library links its
// own CUDA runtime, so it cannot rely on the caller's cudaSetDevice) and rejects host pointers.
struct DeviceScope {
    int prev = -1, status = CRL_OK;
    bool changed = false;
    explicit DeviceScope(const void* ptr) {
        cudaPointerAttributes at;
        cudaError_t e = cudaPointerGetAttributes(&at, ptr);
        if (e != cudaSuccess) {
            cudaGetLastError();
            status = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CRL_ERR_NO_DEVICE : cuda_status(e);
            return;
        }
        if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
            status = CRL_ERR_BAD_ARGUMENT;      // host memory: this library only takes device pointers
            return;
        }
        if (cudaGetDevice(&prev) == cudaSuccess && prev != at.device) {
            status = cuda_status(cudaSetDevice(at.device));
            changed = (status == CRL_OK);
        }
    }
    ~DeviceScope() {
        if (changed) cudaSetDevice(prev);
    }
};

static size_t queue_bytes(size_t slots) { return (slots * 32 * sizeof(QueueEntry) + 255) & ~(size_t)255; }

// Per-warp live-trajectory count at or below which a drained warp hands its trajectories to the
// cooperative finishers: opts->coop_threshold (0 = default, < 0 = off), off if the model has no
// cooperative path or its shared-memory tables do not fit the warp's staging area at this horizon.
template <class Mdl, class R>
static int coop_threshold_for(const crl_solver_opts_t* o, int T) {
    if (!CoopSupport<Mdl>::value || o->coop_threshold < 0) return -1;
    (void)T;
    static_assert(!CoopSupport<Mdl>::value || coop_smem_reals<Mdl, kCoopNTFor<Mdl>>() <= 2 * StageLayout<Mdl, R>::STAGE,
                  "the cooperative engine's shared-memory area must fit the warp's staging area");
    const int th = o->coop_threshold == 0 ? 12 : o->coop_threshold;
    return th > 32 ? 32 : th;
}

static unsigned blocks_for(long long work) { return (unsigned)((work + kBlock - 1) / kBlock); }

// model x dtype dispatch: the body (variadic, may contain commas) sees `Mdl` and `R`
#ifdef CRL_DEV_FAST   // development builds: ThreeLink / float64 only (compiles in a fraction of the time)
#define CRL_DISPATCH(model, dtype, ...)                                                  \
    do {                                                                                 \
        if ((dtype) == CRL_F64 && (model) == CRL_THREE_LINK) {                           \
            using R = double;                                                            \
            using Mdl = ThreeLink;                                                       \
            __VA_ARGS__;                                                                 \
        } else if ((dtype) == CRL_F64 && (model) == CRL_THREE_LINK_BARRIER) {            \
            using R = double;                                                            \
            using Mdl = Barrier<ThreeLink>;                                              \
            __VA_ARGS__;                                                                 \
        }                                                                                \
    } while (0)
#else
#define CRL_DISPATCH(model, dtype, ...)                                                  \
    do {                                                                                 \
        if ((dtype) == CRL_F64) {                                                        \
            using R = double;                                                            \
            switch (model) {                                                             \
                case CRL_TWO_LINK: { using Mdl = TwoLink; __VA_ARGS__; } break;                 \
                case CRL_THREE_LINK: { using Mdl = ThreeLink; __VA_ARGS__; } break;             \
                case CRL_ROBOTIC_ARM: { using Mdl = RoboticArm; __VA_ARGS__; } break;           \
                case CRL_TWO_LINK_BARRIER: { using Mdl = Barrier<TwoLink>; __VA_ARGS__; } break;        \
                case CRL_THREE_LINK_BARRIER: { using Mdl = Barrier<ThreeLink>; __VA_ARGS__; } break;    \
                case CRL_ROBOTIC_ARM_BARRIER: { using Mdl = Barrier<RoboticArm>; __VA_ARGS__; } break;  \
            }                                                                            \
        } else {                                                                         \
            using R = float;                                                             \
            switch (model) {                                                             \
                case CRL_TWO_LINK: { using Mdl = TwoLink; __VA_ARGS__; } break;                 \
                case CRL_THREE_LINK: { using Mdl = ThreeLink; __VA_ARGS__; } break;             \
                case CRL_ROBOTIC_ARM: { using Mdl = RoboticArm; __VA_ARGS__; } break;           \
            }                                                                            \
        }                                                                                \
    } while (0)
#endif

// Launch configurations of the solve.  opts->reserved selects one (0 = default):
//   2      : lane-per-trajectory kernel, 4 warps per CTA, 2 resident CTAs per SM, warps free-running
//   12     : the same, warps of a CTA phase-locked (SYNC) - the default
//   13     : 8 warps per CTA, 1 resident CTA per SM, phase-locked (all eight warps of an SM run the same loop body at the
//            same time: two warps per scheduler share the instruction stream).  EXPERIMENTAL, not the default: measured
//            1 - 5 % faster than 12 on the benchmark batches (bit-identical there), but a lane-only solve (coop_threshold
//            < 0) of 70 problems at T = 2 did not terminate under it on a workspace holding stale data (found at the end
//            of round 2, profiles/r02/hang_probe_variant13.txt; configuration 12 passes the same probe) - unresolved.
//   20, 21 : the cooperative engine as the whole solve (coop_solve_kernel, 4 trajectories per warp), 2 / 3
//            resident CTAs per SM; 22, 23 (development builds, -DCRL_DEV_FAST): 2 / 8 trajectories per warp
//            (element-per-lane / pair-group backward pass)
// Every configuration produces the same results bit for bit.
template <class R> struct SolveCfg;
template <> struct SolveCfg<double> { static constexpr int kDefault = 12; };
template <> struct SolveCfg<float> { static constexpr int kDefault = 12; };

// A requested grid (opts->grid_blocks) is honoured up to the number of CTAs the device can hold at once.
static int clamp_grid(int requested, int resident) { return requested > 0 && requested < resident ? requested : resident; }

// Persistent kernels whose CTAs wait for each other (straggler hand-over) are launched cooperatively: the runtime
// then guarantees that all CTAs are resident together - an oversized grid fails loudly (cudaErrorCooperativeLaunchTooLarge)
// and a launch that finds part of the device busy with another kernel waits instead of starting a partial grid that
// would spin for CTAs that can never be scheduled.
template <class Args>
static cudaError_t launch_resident(const void* kernel, int grid_blocks, int threads, int smem, cudaStream_t s, const Args& a) {
    Args copy = a;
    void* params[] = {&copy};
    return cudaLaunchCooperativeKernel(kernel, dim3((unsigned)grid_blocks), dim3((unsigned)threads), params, (size_t)smem, s);
}

template <class Mdl, class R, int WARPS, int MINB, bool SYNC>
struct SolveVariant {
    static constexpr int kWarps = WARPS;
    static long long slot_elems(int T) { return slot_reals<Mdl>(T); }
    static constexpr int kSmem = WARPS * StageLayout<Mdl, R>::kWarpBytes;
    static int grid(int requested) {
        if (cudaFuncSetAttribute(solve_kernel<Mdl, R, WARPS, MINB, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmem) != cudaSuccess)
            return -1;
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_kernel<Mdl, R, WARPS, MINB, SYNC>, WARPS * 32,
                                                          kSmem) != cudaSuccess)
            return -1;
        if (per_sm < 1) per_sm = 1;
        // the tail's finishers wait for every CTA of the grid to report: the grid must be co-resident, so a requested
        // size is clamped to what fits the device (and the launch below is cooperative, which enforces it)
        return clamp_grid(requested, sms * per_sm);
    }
    static cudaError_t launch(int grid_blocks, cudaStream_t s, const SolveArgs<R>& a) {
        return launch_resident((const void*)solve_kernel<Mdl, R, WARPS, MINB, SYNC>, grid_blocks, WARPS * 32, kSmem, s, a);
    }
};

// The cooperative engine as the whole solve (coop_solve_kernel), NT trajectories per warp.
template <class Mdl, class R, int WARPS, int MINB, int NT>
struct CoopVariant {
    static constexpr int kWarps = WARPS;
    static long long slot_elems(int T) { return coop_scratch_reals<Mdl, NT>(T); }
    static constexpr int kSmem = WARPS * coop_smem_reals<Mdl, NT>() * (int)sizeof(R);
    static int grid(int requested) {
        if (cudaFuncSetAttribute(coop_solve_kernel<Mdl, R, WARPS, MINB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmem) != cudaSuccess)
            return -1;
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_solve_kernel<Mdl, R, WARPS, MINB, NT>, WARPS * 32,
                                                          kSmem) != cudaSuccess)
            return -1;
        if (per_sm < 1) per_sm = 1;
        return clamp_grid(requested, sms * per_sm);
    }
    static cudaError_t launch(int grid_blocks, cudaStream_t s, const SolveArgs<R>& a) {
        return launch_resident((const void*)coop_solve_kernel<Mdl, R, WARPS, MINB, NT>, grid_blocks, WARPS * 32, kSmem, s, a);
    }
};

#define CRL_COOP_VARIANT(W, B, NT, ...)                                            \
    if constexpr (CoopSupport<Mdl>::value) { using V = CoopVariant<Mdl, R, W, B, NT>; __VA_ARGS__; }

#ifdef CRL_DEV_FAST
#define CRL_SOLVE_VARIANT(v, ...)                                                  \
    switch (v) {                                                                   \
        case 2: if constexpr (!Mdl::kBarrier) { using V = SolveVariant<Mdl, R, 4, 2, false>; __VA_ARGS__; } break;  \
        case 12: { using V = SolveVariant<Mdl, R, 4, 2, true>; __VA_ARGS__; } break;  \
        case 13: { using V = SolveVariant<Mdl, R, 8, 1, true>; __VA_ARGS__; } break;  \
        case 20: { CRL_COOP_VARIANT(4, 2, kCoopNTFor<Mdl>, __VA_ARGS__) } break;         \
        case 21: { CRL_COOP_VARIANT(4, 3, kCoopNTFor<Mdl>, __VA_ARGS__) } break;         \
        case 22: if constexpr (!Mdl::kBarrier) { CRL_COOP_VARIANT(4, 2, 2, __VA_ARGS__) } break;   \
        case 23: if constexpr (!Mdl::kBarrier) { CRL_COOP_VARIANT(4, 2, 8, __VA_ARGS__) } break;   \
        default: break;                                                            \
    }
#else
#define CRL_SOLVE_VARIANT(v, ...)                                                  \
    switch (v) {                                                                   \
        case 2: if constexpr (!Mdl::kBarrier) { using V = SolveVariant<Mdl, R, 4, 2, false>; __VA_ARGS__; } break;  \
        case 12: { using V = SolveVariant<Mdl, R, 4, 2, true>; __VA_ARGS__; } break;  \
        case 13: { using V = SolveVariant<Mdl, R, 8, 1, true>; __VA_ARGS__; } break;  \
        case 20: { CRL_COOP_VARIANT(4, 2, kCoopNTFor<Mdl>, __VA_ARGS__) } break;         \
        case 21: { CRL_COOP_VARIANT(4, 3, kCoopNTFor<Mdl>, __VA_ARGS__) } break;         \
        default: break;                                                            \
    }
#endif

// grid size and warps per CTA of the configuration `variant` (0 = default); grid < 0 on error
template <class Mdl, class R>
static void solve_config(int requested, int variant, int T, int* grid, int* warps, long long* slot_elems) {
    const int v = variant > 0 ? variant : SolveCfg<R>::kDefault;
    *grid = -1;
    *warps = 4;
    *slot_elems = 0;
    CRL_SOLVE_VARIANT(v, (*grid = V::grid(requested), *warps = V::kWarps, *slot_elems = V::slot_elems(T)));
}

// Configuration used when the caller leaves the choice open (opts->reserved == 0): the lane-per-trajectory
// kernel, unless the batch does not even fill its resident lanes once (kAutoCoopBatch = 148 SMs x 8 warps x
// 32 lanes): such a solve is all tail, and the cooperative engine as the whole solve is 5 - 25 % faster
// (measured, ThreeLink to convergence: 82 vs 105 ms at 1,024 trajectories, 152 vs 192 ms at 16,384, 210 vs
// 231 ms at 32,768; 333 vs 318 ms at 65,536) and 6x faster per iteration for a single problem.
constexpr long long kAutoCoopBatch = 148LL * 8 * 32;
template <class Mdl, class R>
static int effective_variant(const crl_problem_t* prob, const crl_solver_opts_t* opts) {
    if (opts->reserved > 0) return opts->reserved;
    if (CoopSupport<Mdl>::value && opts->coop_threshold >= 0 && prob->batch <= kAutoCoopBatch) return 20;
    return SolveCfg<R>::kDefault;
}

template <class Mdl, class R>
static cudaError_t solve_launch(int variant, int grid, cudaStream_t s, const SolveArgs<R>& a) {
    const int v = variant > 0 ? variant : SolveCfg<R>::kDefault;
    cudaError_t e = cudaErrorInvalidValue;
    CRL_SOLVE_VARIANT(v, e = V::launch(grid, s, a));
    return e;
}

}  // namespace crl

using namespace crl;

extern "C" {

int crl_abi_version(void) { return CRL_ABI_VERSION; }

const char* crl_status_string(int status) {
    switch (status) {
        case CRL_OK: return "ok";
        case CRL_ERR_BAD_ARGUMENT: return "bad argument";
        case CRL_ERR_UNSUPPORTED_MODEL: return "unsupported model (no device model compiled for this scenario)";
        case CRL_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case CRL_ERR_NO_DEVICE: return "no CUDA device";
        default: break;
    }
    if (status <= CRL_ERR_CUDA) return cudaGetErrorString((cudaError_t)(CRL_ERR_CUDA - status));
    return "unknown status";
}

int crl_model_dims(int model, int* n, int* m, int* p) {
    int N, M, P;
    switch (model) {
        case CRL_TWO_LINK: N = TwoLink::N; M = TwoLink::M; P = TwoLink::P; break;
        case CRL_THREE_LINK: N = ThreeLink::N; M = ThreeLink::M; P = ThreeLink::P; break;
        case CRL_ROBOTIC_ARM: N = RoboticArm::N; M = RoboticArm::M; P = RoboticArm::P; break;
        case CRL_TWO_LINK_BARRIER: N = TwoLink::N; M = TwoLink::M; P = Barrier<TwoLink>::P; break;
        case CRL_THREE_LINK_BARRIER: N = ThreeLink::N; M = ThreeLink::M; P = Barrier<ThreeLink>::P; break;
        case CRL_ROBOTIC_ARM_BARRIER: N = RoboticArm::N; M = RoboticArm::M; P = Barrier<RoboticArm>::P; break;
        default: return gen_dims(model, n, m, p) == 0 ? CRL_OK : CRL_ERR_UNSUPPORTED_MODEL;
    }
    if (n) *n = N;
    if (m) *m = M;
    if (p) *p = P;
    return CRL_OK;
}

int crl_model_default_bounds(int model, double* u_lo, double* u_hi) {
    double lo, hi;
    switch (model) {
        case CRL_TWO_LINK: case CRL_TWO_LINK_BARRIER: lo = TwoLink::u_lo(0); hi = TwoLink::u_hi(0); break;
        case CRL_THREE_LINK: case CRL_THREE_LINK_BARRIER: lo = ThreeLink::u_lo(0); hi = ThreeLink::u_hi(0); break;
        case CRL_ROBOTIC_ARM: case CRL_ROBOTIC_ARM_BARRIER: lo = RoboticArm::u_lo(0); hi = RoboticArm::u_hi(0); break;
        default: return CRL_ERR_UNSUPPORTED_MODEL;
    }
    if (u_lo) *u_lo = lo;
    if (u_hi) *u_hi = hi;
    return CRL_OK;
}

int crl_model_bounds(int model, double* lo, double* hi) {
    if (gen_is_model(model)) return gen_bounds(model, lo, hi) == 0 ? CRL_OK : CRL_ERR_UNSUPPORTED_MODEL;
    int n = 0, m = 0;
    double ulo = 0.0, uhi = 0.0;
    int st = crl_model_dims(model, &n, &m, nullptr);
    if (st) return st;
    st = crl_model_default_bounds(model, &ulo, &uhi);
    if (st) return st;
    for (int i = 0; i < n + m; ++i) {
        if (lo) lo[i] = i < n ? -(double)INFINITY : ulo;
        if (hi) hi[i] = i < n ? (double)INFINITY : uhi;
    }
    return CRL_OK;
}

int crl_rollout(const crl_problem_t* prob, const void* x0, const void* u0, int64_t u0_stride_b, void* X, double* J,
                crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!x0 || !X) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(x0);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_rollout(prob, (const double*)x0, (const double*)u0, u0_stride_b, (double*)X, J, s));
    CRL_DISPATCH(prob->model, prob->dtype, (rollout_kernel<Mdl, R><<<blocks_for(prob->batch), kBlock, 0, s>>>(
                                               to_dev<R>(prob), (const R*)x0, (const R*)u0, u0_stride_b, (R*)X, J)));
    return cuda_status(cudaGetLastError());
}

int crl_cost(const crl_problem_t* prob, const void* X, double* J, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!X || !J) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(X);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_cost(prob, (const double*)X, J, s));
    CRL_DISPATCH(prob->model, prob->dtype,
                 (cost_kernel<Mdl, R><<<blocks_for(prob->batch), kBlock, 0, s>>>(to_dev<R>(prob), (const R*)X, J)));
    return cuda_status(cudaGetLastError());
}

int crl_derivs(const crl_problem_t* prob, const void* X, void* F, void* c, void* C, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!X) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(X);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_derivs(prob, (const double*)X, (double*)F, (double*)c, (double*)C, s));
    CRL_DISPATCH(prob->model, prob->dtype,
                 (derivs_kernel<Mdl, R><<<blocks_for(prob->batch * prob->horizon), kBlock, 0, s>>>(
                     to_dev<R>(prob), (const R*)X, (R*)F, (R*)c, (R*)C)));
    return cuda_status(cudaGetLastError());
}

int crl_backward(const crl_problem_t* prob, const void* X, void* K, void* k, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!X || !K || !k) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(X);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_backward(prob, (const double*)X, (double*)K, (double*)k, s));
    CRL_DISPATCH(prob->model, prob->dtype, (backward_kernel<Mdl, R><<<blocks_for(prob->batch), kBlock, 0, s>>>(
                                               to_dev<R>(prob), (const R*)X, (R*)K, (R*)k)));
    return cuda_status(cudaGetLastError());
}

int crl_backward_dense(int32_t n, int32_t m, int32_t dtype, int64_t batch, int32_t horizon, const void* F, const void* c,
                       const void* C, void* K, void* k, crl_stream_t stream) {
    if (batch < 0 || horizon < 1 || (dtype != CRL_F64 && dtype != CRL_F32)) return CRL_ERR_BAD_ARGUMENT;
    if (!((n == 6 && m == 2) || (n == 8 && m == 3) || (n == 4 && m == 1) || (n == 5 && m == 1) || (n == 4 && m == 2)))
        return CRL_ERR_UNSUPPORTED_MODEL;
    if (n <= 5 && dtype != CRL_F64) return CRL_ERR_UNSUPPORTED_MODEL;          // generated models: FP64 only
    if (batch == 0) return CRL_OK;
    if (!F || !c || !C || !K || !k) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(F);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = blocks_for(batch);
#define CRL_BD(NN, MM, RR)                                                                                           \
    backward_dense_kernel<NN, MM, RR><<<g, kBlock, 0, s>>>(batch, horizon, (const RR*)F, (const RR*)c, (const RR*)C, \
                                                           (RR*)K, (RR*)k)
    if (n == 4 && m == 1) CRL_BD(4, 1, double);
    else if (n == 5) CRL_BD(5, 1, double);
    else if (n == 4) CRL_BD(4, 2, double);
    else if (n == 6 && dtype == CRL_F64) CRL_BD(6, 2, double);
    else if (n == 6) CRL_BD(6, 2, float);
    else if (dtype == CRL_F64) CRL_BD(8, 3, double);
    else CRL_BD(8, 3, float);
#undef CRL_BD
    return cuda_status(cudaGetLastError());
}

int crl_update_traj(const crl_problem_t* prob, const void* X, const void* K, const void* k, double alpha, void* X_new,
                    double* J_new, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!X || !K || !k || !X_new) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(X);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_update(prob, (const double*)X, (const double*)K, (const double*)k, alpha, (double*)X_new, J_new, s));
    CRL_DISPATCH(prob->model, prob->dtype,
                 (update_traj_kernel<Mdl, R><<<blocks_for(prob->batch), kBlock, 0, s>>>(
                     to_dev<R>(prob), (const R*)X, (const R*)K, (const R*)k, alpha, (R*)X_new, J_new)));
    return cuda_status(cudaGetLastError());
}

int crl_forward_pass(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* X, const void* K,
                     const void* k, const double* J_last, void* X_new, double* J_new, int32_t* alpha_index,
                     uint8_t* stop, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    st = check_opts(opts);
    if (st) return st;
    if (prob->batch == 0) return CRL_OK;
    if (!X || !K || !k || !J_last || !X_new || !J_new || !alpha_index || !stop) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(X);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model)) return cuda_status((cudaError_t)gen_launch_forward(prob, opts, (const double*)X, (const double*)K, (const double*)k, J_last, (double*)X_new, J_new, alpha_index, stop, s));
    CRL_DISPATCH(prob->model, prob->dtype,
                 (forward_pass_kernel<Mdl, R><<<blocks_for(prob->batch), kBlock, 0, s>>>(
                     to_dev<R>(prob), to_dev(opts), (const R*)X, (const R*)K, (const R*)k, J_last, (R*)X_new, J_new,
                     alpha_index, stop)));
    return cuda_status(cudaGetLastError());
}

size_t crl_solve_workspace_bytes(const crl_problem_t* prob, const crl_solver_opts_t* opts) {
    if (check_problem(prob) || check_opts(opts)) return 0;
    if (gen_is_model(prob->model)) return gen_solve_workspace_bytes(prob);
    long long reals = 0;
    int grid = 0, warps = 0;
    CRL_DISPATCH(prob->model, prob->dtype,
                 (solve_config<Mdl, R>(opts->grid_blocks, effective_variant<Mdl, R>(prob, opts), prob->horizon, &grid, &warps,
                                       &reals)));
    if (grid <= 0) return 0;
    const size_t elem = prob->dtype == CRL_F64 ? 8 : 4;
    const size_t slots = (size_t)grid * warps;
    // control block, straggler queue (one entry per lane of the grid), warp slots
    return kCtlBytes + queue_bytes(slots) + slots * (size_t)reals * elem;
}

static int solve_impl(const crl_problem_t* prob, const crl_solver_opts_t* opts, int n_stage, int32_t* stage_iters,
                      const void* x0, const void* u0, int64_t u0_stride_b, const double* J_init, void* X, double* J,
                      int32_t* iters, uint8_t* converged, double* J_trace, int8_t* alpha_trace, void* workspace,
                      size_t workspace_bytes, crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    st = check_opts(opts);
    if (st) return st;
    if (n_stage > 1) {
        // one parameter row per stage, constant over the horizon; barrier models only; no per-iteration traces
        const bool barrier = prob->model == CRL_TWO_LINK_BARRIER || prob->model == CRL_THREE_LINK_BARRIER ||
                             prob->model == CRL_ROBOTIC_ARM_BARRIER || gen_is_barrier(prob->model);
        if (!barrier) return CRL_ERR_UNSUPPORTED_MODEL;
        if (n_stage >= 128 || opts->max_iter > kStageItMask || J_trace || alpha_trace) return CRL_ERR_BAD_ARGUMENT;
        if (prob->params_stride_t != 0 && !gen_is_barrier(prob->model)) return CRL_ERR_BAD_ARGUMENT;
        if (prob->batch > 0 && !stage_iters) return CRL_ERR_BAD_ARGUMENT;
    }
    if (prob->batch == 0) return CRL_OK;
    if (!x0 || !J || !iters || !converged || !workspace) return CRL_ERR_BAD_ARGUMENT;
    if (opts->max_iter < 1) return CRL_ERR_BAD_ARGUMENT;
    if (((uintptr_t)workspace & 255u) != 0) return CRL_ERR_BAD_ARGUMENT;
    const size_t need = crl_solve_workspace_bytes(prob, opts);
    if (need == 0) return CRL_ERR_NO_DEVICE;
    if (workspace_bytes < need) return CRL_ERR_WORKSPACE_TOO_SMALL;
    DeviceScope scope(x0);
    if (scope.status) return scope.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (gen_is_model(prob->model))
        return cuda_status((cudaError_t)gen_launch_solve(prob, opts, (const double*)x0, (const double*)u0, u0_stride_b, J_init,
                                                         (double*)X, J, iters, converged, J_trace, alpha_trace, workspace, s,
                                                         n_stage, stage_iters));
    size_t qbytes = 0;
    int grid = 0, warps = 0;
    long long slot_elems = 0;
    CRL_DISPATCH(prob->model, prob->dtype,
                 (solve_config<Mdl, R>(opts->grid_blocks, effective_variant<Mdl, R>(prob, opts), prob->horizon, &grid, &warps,
                                       &slot_elems)));
    if (grid <= 0) return CRL_ERR_BAD_ARGUMENT;
    qbytes = queue_bytes((size_t)grid * warps);
    cudaError_t e = cudaMemsetAsync(workspace, 0, kCtlBytes + qbytes, s);   // counters and the queue's ready flags
    if (e != cudaSuccess) return cuda_status(e);
    CRL_DISPATCH(prob->model, prob->dtype, {
        SolveArgs<R> a;
        a.pb = to_dev<R>(prob);
        a.op = to_dev(opts);
        a.x0 = (const R*)x0;
        a.u0 = (const R*)u0;
        a.u0_sb = u0_stride_b;
        a.J_init = J_init;
        a.X_out = (R*)X;
        a.J_out = J;
        a.iters_out = iters;
        a.conv_out = converged;
        a.J_trace = J_trace;
        a.alpha_trace = alpha_trace;
        a.ctl = (unsigned long long*)workspace;
        a.queue = (QueueEntry*)((char*)workspace + kCtlBytes);
        a.ws = (R*)((char*)workspace + kCtlBytes + qbytes);
        a.coop_threshold = coop_threshold_for<Mdl, R>(opts, prob->horizon);
        a.slot_elems = slot_elems;
        a.n_stage = n_stage;
        a.stage_iters = stage_iters;
        e = solve_launch<Mdl, R>(effective_variant<Mdl, R>(prob, opts), grid, s, a);
    });
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cuda_status(e);
    }
    return cuda_status(cudaGetLastError());
}

int crl_solve(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* x0, const void* u0,
              int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters, uint8_t* converged,
              double* J_trace, int8_t* alpha_trace, void* workspace, size_t workspace_bytes, crl_stream_t stream) {
    return solve_impl(prob, opts, 1, nullptr, x0, u0, u0_stride_b, J_init, X, J, iters, converged, J_trace, alpha_trace,
                      workspace, workspace_bytes, stream);
}

int crl_solve_stages(const crl_problem_t* prob, const crl_solver_opts_t* opts, int32_t n_stages, const void* x0,
                     const void* u0, int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters,
                     uint8_t* converged, int32_t* stage_iters, void* workspace, size_t workspace_bytes,
                     crl_stream_t stream) {
    if (n_stages < 1) return CRL_ERR_BAD_ARGUMENT;
    return solve_impl(prob, opts, n_stages, stage_iters, x0, u0, u0_stride_b, J_init, X, J, iters, converged, nullptr,
                      nullptr, workspace, workspace_bytes, stream);
}

}  // extern "C"
