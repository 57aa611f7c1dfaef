// Jacobian of a learned dynamics model (three-layer MLP) at many points at once - the F matrices of NNiLQR.
//
// Reference: CuriousRL/algorithm/ilqr_solver/nn_ilqr.py:265-274 (`eval_grad_dynamic_model`): per time step the
// network is evaluated on n copies of the row and back-propagated with the identity (n backward passes through
// cuBLAS/autograd).  The networks it is used with (`SmallNetwork`, `LargeNetwork`, nn_ilqr.py:66-100) are
//     y = W3 relu(BN2(W2 relu(BN1(W1 x + b1)) + b2)) + b3
// and in eval mode BatchNorm is an affine map, folded into (W, b) by the caller.  With m1 = [z1 > 0], m2 = [z2 > 0]
//     dy/dx = (W3 diag(m2)) W2 diag(m1) W1.
// For S points the dominant product is  P = G2 W2  with  G2[(s, r), :] = W3[r, :] * m2[s, :]  - an (S n) x H2 x H1
// GEMM (S = trajectories x T) - the one dense contraction of the whole product, so it runs on the 5th-generation
// tensor cores: tcgen05.mma (kind::tf32, FP32 accumulators in tensor memory) issued by one thread per CTA.
//
//   mlp_forward_kernel   tile = 128 points.  A = relu(W1 x + b1) is formed on the fly per 32-wide K block (the first
//                        layer has K = n + m <= 16: CUDA cores), z2 = A W2^T accumulates in TMEM over all of H1 for the
//                        whole width H2 (<= 512 columns), the epilogue adds b2, records the sign bits m2 and applies W3.
//                        Writes y [S][n] and the bit masks m1 [S][H1/32], m2 [S][H2/32].
//   mlp_jacobian_kernel  tile = 128 rows (s, r).  A block = W3 rows masked by m2 (generated in shared memory, never in
//                        HBM), B block = W2^T slice (cp.async, one block ahead); 256-column chunks of P accumulate in one
//                        of two TMEM accumulators over H2; the epilogue reads a chunk (tcgen05.ld), masks it by m1 and
//                        contracts it with W1 (K = 256, N = n + m: CUDA cores, W1 in shared memory) into per-thread
//                        Jacobian rows, spread over the blocks of the next chunk.  Warp-specialised: 8 producer /
//                        epilogue warps and one MMA-issuer warp, coupled by mbarriers only (full / free per stage, acc /
//                        accfree per accumulator), three shared-memory stages.
// Operands are staged in the no-swizzle K-major canonical layout (a block is `float4 blk[8 K-chunks][rows]`: 8-row x
// 16-byte core matrices, leading byte offset = rows * 16, stride byte offset = 128) and made visible to the tensor core
// with fence.proxy.async; completion comes back through tcgen05.commit on mbarriers.
// Measured on a B200 (profiles/r02/nn_jacobian.md): LargeNetwork (11 -> 800 -> 400 -> 8), 409,600 points: 16.7 ms
// (forward 2.8 ms + Jacobian 13.9 ms) = 150 TFLOP/s; the FP32 CUDA-core path does 1.8 TFLOP/s.  The Jacobian kernel is
// bound by its producers (about 3,000 cycles of staging + epilogue per block against 512 cycles of MMA), see the notes.
//
// `crl_mlp_jacobian(..., impl = 0)` is a plain FP32 CUDA-core evaluation of the same formulas (one thread per output
// element, five small kernels): the checker of the tensor-core path in tests and the path for a handful of points.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/curiousrl_b200.h"

namespace crlnn {

constexpr int kThreads = 128;
constexpr int kKB = 32;                 // K elements per staged block (8 chunks of 4 tf32)
constexpr int kMaxD = 16;               // n + m
constexpr int kMaxOut = 16;             // n

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, unsigned cols) {       // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, unsigned cols) {           // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate, M = 128, K = 8 per instruction
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major, no-swizzle operand whose 16-byte K-chunks are `lbo` bytes apart and
// whose 8-row groups are 128 bytes apart (cute::UMMA::SmemDescriptor: start >> 4 at bit 0, leading byte offset >> 4 at
// bit 16, stride byte offset >> 4 at bit 32, version 1 at bit 46, layout type 0 = SWIZZLE_NONE at bit 61).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (bit 4), A = B = tf32 (value 2 at bits 7 and 10), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ bool mask_bit(const uint32_t* words, int i) { return (words[i >> 5] >> (i & 31)) & 1u; }

struct MlpDims {
    int d, h1, h2, n;          // h1, h2 multiples of 32 (zero-padded by the caller)
    long long S;
};

// ------------------------------------------------------------------------------------------------ forward
// dynamic shared memory: A blocks 2 x [8][128] float4 | B blocks 2 x [8][h2] float4 | W1 [h1][d] | b1 [h1] | W3 [n][h2] |
// b2 [h2] | barriers
__global__ void __launch_bounds__(kThreads, 1)
mlp_forward_kernel(MlpDims dm, const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                   const float* __restrict__ b3, float* __restrict__ y, uint32_t* __restrict__ m1, uint32_t* __restrict__ m2) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int d = dm.d, h1 = dm.h1, h2 = dm.h2, n = dm.n;
    float4* Ablk = reinterpret_cast<float4*>(smem);                          // [2][8][128]
    float4* Bblk = Ablk + 2 * 8 * 128;                                       // [2][8][h2]
    float* W1s = reinterpret_cast<float*>(Bblk + 2 * 8 * h2);                // [h1][d]
    float* b1s = W1s + h1 * d;
    float* W3s = b1s + h1;                                                   // [n][h2]
    float* b2s = W3s + n * h2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2s + h2);                  // free[0], free[1], acc
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const unsigned tmem_cols = h2 <= 32 ? 32 : h2 <= 64 ? 64 : h2 <= 128 ? 128 : h2 <= 256 ? 256 : 512;
    const int halves = h2 > 256 ? 2 : 1, nh = h2 / halves;

    for (int i = tid; i < h1 * d; i += kThreads) W1s[i] = w1[i];
    for (int i = tid; i < h1; i += kThreads) b1s[i] = b1[i];
    for (int i = tid; i < n * h2; i += kThreads) W3s[i] = w3[i];
    for (int i = tid; i < h2; i += kThreads) b2s[i] = b2[i];
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = instr_desc(nh);
    const int nkb = h1 / kKB;
    const long long tiles = (dm.S + 127) / 128;
    unsigned blk = 0, tile_no = 0;                                           // CTA-uniform counters of staged blocks / tiles

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_no) {
        const long long s = tile * 128 + tid;
        const bool valid = s < dm.S;
        float xr[kMaxD];
#pragma unroll
        for (int j = 0; j < kMaxD; ++j) xr[j] = (valid && j < d) ? x[s * d + j] : 0.0f;
        for (int kb = 0; kb < nkb; ++kb, ++blk) {
            const int buf = blk & 1;
            if (blk >= 2) mbar_wait(bars + buf, ((blk >> 1) - 1) & 1);       // the MMAs that read this buffer are done
            // A block: relu(W1 x + b1) for hidden units kb*32 .. +31 of this thread's point; sign bits -> m1
            float4* Ab = Ablk + buf * (8 * 128);
            uint32_t bits = 0u;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int h = kb * kKB + c * 4 + e;
                    float z = b1s[h];
                    const float* wr = W1s + h * d;
#pragma unroll
                    for (int j = 0; j < kMaxD; ++j)
                        if (j < d) z = fmaf(wr[j], xr[j], z);
                    const bool on = z > 0.0f;
                    bits |= (on ? 1u : 0u) << (c * 4 + e);
                    v[e] = on ? z : 0.0f;
                }
                Ab[c * 128 + tid] = make_float4(v[0], v[1], v[2], v[3]);
            }
            if (valid) m1[s * (h1 / 32) + kb] = bits;
            // B block: rows = hidden units of layer 2, K = this block's 32 units of layer 1 (W2 is [h2][h1], K-major)
            float4* Bb = Bblk + buf * (8 * h2);
            for (int r = tid; r < h2; r += kThreads) {
                const float4* src = reinterpret_cast<const float4*>(w2 + (size_t)r * h1 + kb * kKB);
#pragma unroll
                for (int c = 0; c < 8; ++c) Bb[c * h2 + r] = __ldg(src + c);
            }
            proxy_fence();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a0 = smem_u32(Ab), b0 = smem_u32(Bb);
#pragma unroll
                for (int st = 0; st < kKB / 8; ++st) {
                    const uint64_t da = smem_desc(a0 + st * 2 * (128 * 16), 128 * 16);
                    for (int hh = 0; hh < halves; ++hh) {
                        const uint64_t db = smem_desc(b0 + st * 2 * (h2 * 16) + hh * nh * 16, h2 * 16);
                        tc_mma_tf32(tmem + hh * nh, da, db, idesc, kb > 0 || st > 0);
                    }
                }
                tc_commit(bars + buf);
                if (kb == nkb - 1) tc_commit(bars + 2);
            }
        }
        // epilogue: z2 + b2 -> sign bits m2, a2 = relu, y = W3 a2 + b3
        mbar_wait(bars + 2, tile_no & 1);
        tc_fence_after();
        float yo[kMaxOut];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) yo[o] = 0.0f;
        for (int g = 0; g < h2 / 32; ++g) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, v);
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int h = g * 32 + i;
                const float z = v[i] + b2s[h];
                const bool on = z > 0.0f;
                bits |= (on ? 1u : 0u) << i;
                const float a2 = on ? z : 0.0f;
#pragma unroll
                for (int o = 0; o < kMaxOut; ++o)
                    if (o < n) yo[o] = fmaf(W3s[o * h2 + h], a2, yo[o]);
            }
            if (valid) m2[s * (h2 / 32) + g] = bits;
        }
        if (valid) {
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o)
                if (o < n) y[s * n + o] = yo[o] + b3[o];
        }
        tc_fence_before();
        __syncthreads();                                                     // TMEM is overwritten by the next tile
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ Jacobian
// One flat loop over the CTA's blocks b = (tile, chunk, K block): chunks are 256 hidden units of layer 1 wide (two TMEM
// accumulators of 256 columns, used alternately), K blocks 32 hidden units of layer 2 deep.  Three shared-memory stages:
// the B block of b + 1 is fetched with cp.async while A(b) is generated and the MMAs of b - 1 run; a stage is reused when
// the commit of the MMAs that read it has arrived on its mbarrier.  The epilogue of chunk c (TMEM -> registers, mask by m1,
// contraction with W1) is spread, one 32-column group per block, over the first blocks of chunk c + 1, so the tensor
// core keeps working through it; only the last chunk of a tile is drained at once.
// dynamic shared memory: A stages 3 x [8][128] float4 | B stages 3 x [8][256] float4 | W3 [n][h2 + 4] | W1 [h1][dp] |
// masks 2 x (m2 [TS][h2/32], m1 [TS][h1/32]) | barriers
constexpr int kJStages = 3;
constexpr int kJChunk = 256;
constexpr int kJThreads = 256;          // two threads per tile row

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kJThreads) : "memory"); }

// Flat position of a block inside the CTA's work: (tile counter, chunk, K block) and stage / use counters, advanced
// without divisions.
struct JPos {
    int ti, ch, kb, stage;
    unsigned use;               // how often `stage` has been used before this block
    __device__ void next(int nchunks, int nkb) {
        if (++kb == nkb) { kb = 0; if (++ch == nchunks) { ch = 0; ++ti; } }
        if (++stage == kJStages) { stage = 0; ++use; }
    }
};

// Warps 0-7 (two threads per tile row) are PRODUCERS and EPILOGUE: they stage blocks and arrive on the stage's `full`
// barrier; warp 8 is the MMA ISSUER: one lane waits for `full`, issues the block's tcgen05.mma's and commits them to the
// stage's `free` barrier (and, after a chunk's last block, to the accumulator's `acc` barrier).  The producers never wait
// for an MMA to be issued - only for the stage they refill to have been read and for a finished chunk to drain - so
// staging, the tensor core and the epilogue overlap.  `accfree` tells the issuer that a drained accumulator may be
// overwritten.
__global__ void __launch_bounds__(kJThreads + 32, 1)
mlp_jacobian_kernel(MlpDims dm, const float* __restrict__ w1, const float* __restrict__ w2t, const float* __restrict__ w3,
                    const uint32_t* __restrict__ m1, const uint32_t* __restrict__ m2, float* __restrict__ jac) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = (tid >> 7) & 1;                        // two threads per tile row: K-chunks / column groups split
    const int d = dm.d, h1 = dm.h1, h2 = dm.h2, n = dm.n;
    const int dp = (d + 3) & ~3;                                             // row pitch of W1 in shared memory
    const int TS = 128 / n;                                                  // points per tile
    const int w1w = h1 / 32, w2w = h2 / 32, mw = TS * (w1w + w2w);
    const int w3p = h2 + 4;                                                  // row pitch of W3: the n rows a warp reads together fall into different banks
    float4* Ablk = reinterpret_cast<float4*>(smem);                          // [3][8][128]
    float4* Bblk = Ablk + kJStages * 8 * 128;                                // [3][8][256]
    float* W3s = reinterpret_cast<float*>(Bblk + kJStages * 8 * kJChunk);    // [n][h2 + 4]
    float* W1s = W3s + n * w3p;                                              // [h1][dp]
    uint32_t* masks = reinterpret_cast<uint32_t*>(W1s + h1 * dp);            // [2][ m2 [TS][w2w] | m1 [TS][w1w] ]
    uint64_t* bars = reinterpret_cast<uint64_t*>(masks + 2 * mw + ((2 * mw) & 1));
    uint64_t* bar_full = bars;                                               // [3] stage staged          (8 producer warps arrive)
    uint64_t* bar_free = bars + kJStages;                                    // [3] stage read            (tcgen05.commit)
    uint64_t* bar_acc = bars + 2 * kJStages;                                 // [2] chunk accumulated     (tcgen05.commit)
    uint64_t* bar_accfree = bars + 2 * kJStages + 2;                         // [2] accumulator drained   (8 producer warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kJStages + 4);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);                    // [128][kMaxD] partial rows of the second half

    if (tid < kJThreads) {
        for (int i = tid; i < n * h2; i += kJThreads) W3s[(i / h2) * w3p + i % h2] = w3[i];
        for (int i = tid; i < h1 * dp; i += kJThreads) {
            const int h = i / dp, j = i - h * dp;
            W1s[i] = j < d ? w1[h * d + j] : 0.0f;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < kJStages; ++i) { mbar_init(bar_full + i, kJThreads / 32); mbar_init(bar_free + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc + i, 1); mbar_init(bar_accfree + i, kJThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 2 * kJChunk);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nkb = h2 / kKB;
    const int nchunks = (h1 + kJChunk - 1) / kJChunk;
    const long long tiles = (dm.S + TS - 1) / TS;
    const int my_tiles = tiles > blockIdx.x ? (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const long long total = (long long)my_tiles * nchunks * nkb;

    if (warp == kJThreads / 32) {
        // ================================================================ MMA issuer (one lane)
        if ((tid & 31) == 0) {
            JPos p{0, 0, 0, 0, 0u};
            unsigned acc_uses[2] = {0u, 0u};                                 // chunks started per accumulator
            for (long long b = 0; b < total; ++b, p.next(nchunks, nkb)) {
                const int a = p.ch & 1;
                if (p.kb == 0) {
                    if (acc_uses[a] >= 1) mbar_wait(bar_accfree + a, (acc_uses[a] - 1) & 1u);   // the previous chunk on it is drained
                    ++acc_uses[a];
                }
                mbar_wait(bar_full + p.stage, p.use & 1u);
                tc_fence_after();
                const int n0 = p.ch * kJChunk, nc = h1 - n0 < kJChunk ? h1 - n0 : kJChunk;
                const uint32_t idesc = instr_desc(nc);
                const uint32_t a0 = smem_u32(Ablk + p.stage * (8 * 128)), b0 = smem_u32(Bblk + p.stage * (8 * kJChunk));
#pragma unroll
                for (int st = 0; st < kKB / 8; ++st)
                    tc_mma_tf32(tmem + a * kJChunk, smem_desc(a0 + st * 2 * (128 * 16), 128 * 16),
                                smem_desc(b0 + st * 2 * (kJChunk * 16), kJChunk * 16), idesc, p.kb > 0 || st > 0);
                tc_commit(bar_free + p.stage);
                if (p.kb == nkb - 1) tc_commit(bar_acc + a);
            }
        }
    } else {
        // ================================================================ producers + epilogue
        const int ls = row / n, out = row - ls * n;                          // this thread's row: point ls, output `out`
        const float* my_w3 = W3s + out * w3p;
        auto load_masks = [&](long long tile, int slot) {
            const long long s0 = tile * TS;
            uint32_t* dst = masks + slot * mw;
            for (int i = tid; i < TS * w2w; i += kJThreads) {
                const long long sp = s0 + i / w2w;
                dst[i] = sp < dm.S ? m2[sp * w2w + i % w2w] : 0u;
            }
            for (int i = tid; i < TS * w1w; i += kJThreads) {
                const long long sp = s0 + i / w1w;
                dst[TS * w2w + i] = sp < dm.S ? m1[sp * w1w + i % w1w] : 0u;
            }
        };
        // B block of position q: rows = hidden units of the chunk, 32 K values each, fetched asynchronously
        auto issue_b = [&](const JPos& q) {
            if (q.use >= 1) mbar_wait(bar_free + q.stage, (q.use - 1) & 1u);   // the MMAs that read the stage are done
            const int n0 = q.ch * kJChunk, nc = h1 - n0 < kJChunk ? h1 - n0 : kJChunk;
            float4* Bb = Bblk + q.stage * (8 * kJChunk);
            for (int br = tid; br < nc; br += kJThreads) {
                const float* src = w2t + (size_t)(n0 + br) * h2 + q.kb * kKB;
#pragma unroll
                for (int c = 0; c < 8; ++c) cp_async16(Bb + c * kJChunk + br, src + c * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float acc[kMaxD];
        unsigned acc_done[2] = {0u, 0u};                                     // chunks drained per accumulator
        int pend_chunk = -1, pend_group = 0, pend_groups = 0;               // epilogue in progress: chunk, next group, groups
        // one 32-column group of the epilogue of chunk `ch`: P masked by m1, contracted with W1
        auto epilogue_group = [&](int ch, int g, const uint32_t* my_m1) {
            const int n0 = ch * kJChunk;
            float v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (ch & 1) * kJChunk + g * 32, v);
            const uint32_t mb = my_m1[(n0 >> 5) + g];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float pv = (mb >> i) & 1u ? v[i] : 0.0f;
                const float4* wr = reinterpret_cast<const float4*>(W1s + (n0 + g * 32 + i) * dp);
#pragma unroll
                for (int q = 0; q < kMaxD / 4; ++q)
                    if (q * 4 < d) {
                        const float4 w = wr[q];
                        acc[q * 4 + 0] = fmaf(pv, w.x, acc[q * 4 + 0]);
                        acc[q * 4 + 1] = fmaf(pv, w.y, acc[q * 4 + 1]);
                        acc[q * 4 + 2] = fmaf(pv, w.z, acc[q * 4 + 2]);
                        acc[q * 4 + 3] = fmaf(pv, w.w, acc[q * 4 + 3]);
                    }
            }
        };
        // wait until chunk `ch` is accumulated / tell the issuer that it is drained
        auto drain_begin = [&](int ch) {
            mbar_wait(bar_acc + (ch & 1), acc_done[ch & 1] & 1u);
            ++acc_done[ch & 1];
            tc_fence_after();
        };
        auto drain_end = [&](int ch) {
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_accfree + (ch & 1));
        };

        JPos p{0, 0, 0, 0, 0u}, q{0, 0, 0, 0, 0u};                           // current block / the block whose B is fetched
        if (total > 0) {
            load_masks(blockIdx.x, 0);
            issue_b(q);
            q.next(nchunks, nkb);
        }
        producers_sync();
        for (long long b = 0; b < total; ++b, p.next(nchunks, nkb)) {
            const long long tile = blockIdx.x + (long long)p.ti * gridDim.x;
            const int slot = p.ti & 1;
            const bool valid = ls < TS && tile * TS + ls < dm.S;
            const uint32_t* my_m2 = masks + slot * mw + (valid ? ls : 0) * w2w;
            const uint32_t* my_m1 = masks + slot * mw + TS * w2w + (valid ? ls : 0) * w1w;
            if (p.ch == 0 && p.kb == 0) {
#pragma unroll
                for (int j = 0; j < kMaxD; ++j) acc[j] = 0.0f;
                if (p.ti + 1 < my_tiles) load_masks(tile + gridDim.x, slot ^ 1);   // read from the next tile's first block on
            }
            // A block: row (s, r) of G2 = W3[r, :] masked by m2[s, :], 32 hidden units of layer 2 (the stage is free: its
            // previous use was waited for when this block's B was issued).  Formed BEFORE the next block's cp.async's are
            // issued: shared-memory loads queue behind them in the load/store unit.
            float4* Ab = Ablk + p.stage * (8 * 128);
            const uint32_t bits = valid ? my_m2[p.kb] : 0u;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = half * 4 + cc;
                const float4 w = *reinterpret_cast<const float4*>(my_w3 + p.kb * kKB + c * 4);
                Ab[c * 128 + row] = make_float4((bits >> (c * 4 + 0)) & 1u ? w.x : 0.0f, (bits >> (c * 4 + 1)) & 1u ? w.y : 0.0f,
                                                (bits >> (c * 4 + 2)) & 1u ? w.z : 0.0f, (bits >> (c * 4 + 3)) & 1u ? w.w : 0.0f);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");             // B of THIS block (issued one block ago)
            proxy_fence();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_full + p.stage);
            if (b + 1 < total) {                                              // fetch the next block's B while this one is multiplied
                issue_b(q);
                q.next(nchunks, nkb);
            }
            // the epilogue of the previous chunk, one pair of groups per block (all that is left at the chunk's last block)
            if (pend_chunk >= 0) {
                if (pend_group == 0) drain_begin(pend_chunk);
                do {
                    if (pend_group + half < pend_groups) epilogue_group(pend_chunk, pend_group + half, my_m1);
                    pend_group += 2;
                } while (pend_group < pend_groups && p.kb == nkb - 1);
                if (pend_group >= pend_groups) {
                    drain_end(pend_chunk);
                    pend_chunk = -1;
                }
            }
            if (p.kb == nkb - 1) {
                const int n0 = p.ch * kJChunk, nc = h1 - n0 < kJChunk ? h1 - n0 : kJChunk;
                if (p.ch == nchunks - 1) {
                    // last chunk of the tile: drain it now and write the rows out
                    drain_begin(p.ch);
                    for (int g = half; g < nc / 32; g += 2) epilogue_group(p.ch, g, my_m1);
                    drain_end(p.ch);
                    if (half == 1) {
#pragma unroll
                        for (int j = 0; j < kMaxD; ++j) red[row * kMaxD + j] = acc[j];
                    }
                    producers_sync();
                    if (half == 0 && valid) {
                        float* dst = jac + ((tile * TS + ls) * n + out) * d;
#pragma unroll
                        for (int j = 0; j < kMaxD; ++j)
                            if (j < d) dst[j] = acc[j] + red[row * kMaxD + j];
                    }
                    producers_sync();                                         // `red` and the other mask slot are rewritten next
                } else {
                    pend_chunk = p.ch;
                    pend_group = 0;
                    pend_groups = nc / 32;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * kJChunk);
}

// ------------------------------------------------------------------------------------------------ FP32 CUDA-core path
__global__ void naive_layer1(MlpDims dm, const float* x, const float* w1, const float* b1, float* a1, uint32_t* m1) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (s, word)
    const int words = dm.h1 / 32;
    if (i >= dm.S * words) return;
    const long long s = i / words;
    const int wd = (int)(i % words);
    uint32_t bits = 0u;
    for (int e = 0; e < 32; ++e) {
        const int h = wd * 32 + e;
        float z = b1[h];
        for (int j = 0; j < dm.d; ++j) z = fmaf(w1[h * dm.d + j], x[s * dm.d + j], z);
        const bool on = z > 0.0f;
        bits |= (on ? 1u : 0u) << e;
        a1[s * dm.h1 + h] = on ? z : 0.0f;
    }
    m1[i] = bits;
}
__global__ void naive_layer2(MlpDims dm, const float* a1, const float* w2, const float* b2, float* a2, uint32_t* m2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (s, word)
    const int words = dm.h2 / 32;
    if (i >= dm.S * words) return;
    const long long s = i / words;
    const int wd = (int)(i % words);
    uint32_t bits = 0u;
    for (int e = 0; e < 32; ++e) {
        const int h = wd * 32 + e;
        float z = 0.0f;
        for (int k = 0; k < dm.h1; ++k) z = fmaf(w2[(size_t)h * dm.h1 + k], a1[s * dm.h1 + k], z);
        z += b2[h];
        const bool on = z > 0.0f;
        bits |= (on ? 1u : 0u) << e;
        a2[s * dm.h2 + h] = on ? z : 0.0f;
    }
    m2[i] = bits;
}
__global__ void naive_layer3(MlpDims dm, const float* a2, const float* w3, const float* b3, float* y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (s, out)
    if (i >= dm.S * dm.n) return;
    const long long s = i / dm.n;
    const int o = (int)(i % dm.n);
    float z = 0.0f;
    for (int k = 0; k < dm.h2; ++k) z = fmaf(w3[o * dm.h2 + k], a2[s * dm.h2 + k], z);
    y[i] = z + b3[o];
}
__global__ void naive_back2(MlpDims dm, const float* w2t, const float* w3, const uint32_t* m1, const uint32_t* m2, float* g1) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (s, out, h1)
    if (i >= dm.S * dm.n * dm.h1) return;
    const int h = (int)(i % dm.h1);
    const long long so = i / dm.h1;
    const int o = (int)(so % dm.n);
    const long long s = so / dm.n;
    float p = 0.0f;
    if (mask_bit(m1 + s * (dm.h1 / 32), h)) {
        const uint32_t* mm = m2 + s * (dm.h2 / 32);
        for (int k = 0; k < dm.h2; ++k)
            if (mask_bit(mm, k)) p = fmaf(w3[o * dm.h2 + k], w2t[(size_t)h * dm.h2 + k], p);
    }
    g1[i] = p;
}
__global__ void naive_back1(MlpDims dm, const float* g1, const float* w1, float* jac) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (s, out, j)
    if (i >= dm.S * dm.n * dm.d) return;
    const int j = (int)(i % dm.d);
    const long long so = i / dm.d;
    float z = 0.0f;
    for (int h = 0; h < dm.h1; ++h) z = fmaf(g1[so * dm.h1 + h], w1[h * dm.d + j], z);
    jac[i] = z;
}

// Linear filter along the time axis: out[b][t][w] = sum_t' G[t][t'] in[b][t'][w], FP64, terms added in ascending t'
// (zero taps skipped).  G is the matrix of scipy.ndimage.gaussian_filter1d(., sigma, axis=0) (reflecting boundary), built
// by the caller once per horizon; NNiLQR smooths the learned Jacobians F along T with it (nn_ilqr.py:379-380).
__global__ void smooth_time_kernel(long long B, int T, int W, const double* __restrict__ G, const double* __restrict__ in,
                                   double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;        // (b, t, w)
    if (i >= B * T * W) return;
    const int w = (int)(i % W);
    const long long bt = i / W;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const double* g = G + (size_t)t * T;
    const double* src = in + (size_t)b * T * W + w;
    double acc = 0.0;
    for (int u = 0; u < T; ++u) {
        const double c = g[u];
        if (c != 0.0) acc = fma(c, src[(size_t)u * W], acc);
    }
    out[i] = acc;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace crlnn

using namespace crlnn;

extern "C" {

int crl_smooth_time(int64_t batch, int32_t horizon, int32_t width, const double* G, const double* in, double* out,
                    crl_stream_t stream) {
    if (batch < 0 || horizon < 1 || width < 1) return CRL_ERR_BAD_ARGUMENT;
    if (batch == 0) return CRL_OK;
    if (!G || !in || !out || in == out) return CRL_ERR_BAD_ARGUMENT;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, in) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return CRL_ERR_BAD_ARGUMENT;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(at.device) != cudaSuccess) return CRL_ERR_CUDA;
    const long long work = (long long)batch * horizon * width;
    smooth_time_kernel<<<(unsigned)((work + 127) / 128), 128, 0, (cudaStream_t)stream>>>(batch, horizon, width, G, in, out);
    const cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    return e == cudaSuccess ? CRL_OK : CRL_ERR_CUDA - (int)e;
}

size_t crl_mlp_jacobian_workspace_bytes(int32_t d_in, int32_t h1, int32_t h2, int32_t n_out, int64_t n_points, int32_t impl) {
    (void)d_in;
    if (h1 <= 0 || h2 <= 0 || n_out <= 0 || n_points < 0) return 0;
    size_t b = align256((size_t)n_points * (h1 / 32) * 4) + align256((size_t)n_points * (h2 / 32) * 4);
    if (impl == 0)
        b += align256((size_t)n_points * h1 * 4) + align256((size_t)n_points * h2 * 4) +
             align256((size_t)n_points * n_out * h1 * 4);
    return b + 256;
}

int crl_mlp_jacobian(int32_t d_in, int32_t h1, int32_t h2, int32_t n_out, int64_t n_points, const float* x, const float* w1,
                     const float* b1, const float* w2, const float* w2t, const float* b2, const float* w3, const float* b3,
                     float* y, float* jac, void* workspace, size_t workspace_bytes, int32_t impl, crl_stream_t stream) {
    if (d_in < 1 || d_in > kMaxD || n_out < 1 || n_out > kMaxOut || h1 < 32 || h2 < 32 || (h1 % 32) || (h2 % 32) || h2 > 512 ||
        n_points < 0 || (impl != 0 && impl != 1))
        return CRL_ERR_BAD_ARGUMENT;
    if (n_points == 0) return CRL_OK;
    if (!x || !w1 || !b1 || !w2 || !w2t || !b2 || !w3 || !b3 || !y || !jac || !workspace) return CRL_ERR_BAD_ARGUMENT;
    if (((uintptr_t)workspace & 255u) != 0) return CRL_ERR_BAD_ARGUMENT;
    if (workspace_bytes < crl_mlp_jacobian_workspace_bytes(d_in, h1, h2, n_out, n_points, impl)) return CRL_ERR_WORKSPACE_TOO_SMALL;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, x) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return CRL_ERR_BAD_ARGUMENT;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(at.device) != cudaSuccess) return CRL_ERR_CUDA;
    cudaStream_t s = (cudaStream_t)stream;
    MlpDims dm{d_in, h1, h2, n_out, (long long)n_points};
    char* wsp = (char*)workspace;
    uint32_t* m1 = (uint32_t*)wsp;
    wsp += align256((size_t)n_points * (h1 / 32) * 4);
    uint32_t* m2 = (uint32_t*)wsp;
    wsp += align256((size_t)n_points * (h2 / 32) * 4);
    cudaError_t e = cudaSuccess;
    if (impl == 0) {
        float* a1 = (float*)wsp;
        wsp += align256((size_t)n_points * h1 * 4);
        float* a2 = (float*)wsp;
        wsp += align256((size_t)n_points * h2 * 4);
        float* g1 = (float*)wsp;
        auto blocks = [](long long work) { return (unsigned)((work + 127) / 128); };
        naive_layer1<<<blocks(n_points * (h1 / 32)), 128, 0, s>>>(dm, x, w1, b1, a1, m1);
        naive_layer2<<<blocks(n_points * (h2 / 32)), 128, 0, s>>>(dm, a1, w2, b2, a2, m2);
        naive_layer3<<<blocks(n_points * n_out), 128, 0, s>>>(dm, a2, w3, b3, y);
        naive_back2<<<blocks(n_points * n_out * h1), 128, 0, s>>>(dm, w2t, w3, m1, m2, g1);
        naive_back1<<<blocks(n_points * n_out * d_in), 128, 0, s>>>(dm, g1, w1, jac);
        e = cudaGetLastError();
    } else {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, at.device);
        if (sms < 1) sms = 1;
        const size_t f_smem = (size_t)2 * 8 * 128 * 16 + (size_t)2 * 8 * h2 * 16 + ((size_t)h1 * d_in + h1 + (size_t)n_out * h2 + h2) * 4 +
                              3 * 8 + 16;
        const int TS = 128 / n_out;
        const int dp = (d_in + 3) & ~3;
        const size_t j_smem = (size_t)kJStages * 8 * (128 + kJChunk) * 16 + ((size_t)n_out * (h2 + 4) + (size_t)h1 * dp) * 4 +
                              ((size_t)2 * TS * (h1 / 32 + h2 / 32) + 1) * 4 + (2 * kJStages + 4) * 8 + 16 + (size_t)128 * kMaxD * 4;
        if (f_smem > 227 * 1024 || j_smem > 227 * 1024) {
            cudaSetDevice(prev);
            return CRL_ERR_BAD_ARGUMENT;
        }
        e = cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_jacobian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)j_smem);
        if (e == cudaSuccess) {
            const long long f_tiles = (n_points + 127) / 128, j_tiles = (n_points + TS - 1) / TS;
            mlp_forward_kernel<<<(unsigned)(f_tiles < sms ? f_tiles : sms), kThreads, f_smem, s>>>(dm, x, w1, b1, w2, b2, w3, b3, y, m1,
                                                                                                  m2);
            mlp_jacobian_kernel<<<(unsigned)(j_tiles < sms ? j_tiles : sms), kJThreads + 32, j_smem, s>>>(dm, w1, w2t, w3, m1, m2, jac);
            e = cudaGetLastError();
        }
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return CRL_ERR_CUDA - (int)e;
    }
    return CRL_OK;
}

}  // extern "C"
