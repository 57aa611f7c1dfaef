// Shared plumbing of the batched-iLQR kernels: host/device qualifiers, explicit-rounding
// arithmetic helpers and strided trajectory views.
//
// The numerical core (ilqr_models.cuh, ilqr_core.cuh) is written as __host__ __device__
// templates so that the exact code the kernels run can also be compiled by g++ into a
// test-only host library (tests/hostsim) and checked against the oracle without a GPU.
// The shipped library contains no host execution path.
//
// Floating-point contraction is OFF for the whole build (nvcc -fmad=false, g++
// -ffp-contract=off): every fused multiply-add in this code is written as fma() on
// purpose, where the reference's BLAS kernels accumulate with FMA; everything else
// rounds after each operation like the reference's scalar (numba/LLVM) code.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CRL_HD __host__ __device__ __forceinline__
#define CRL_UNROLL _Pragma("unroll")
#define CRL_KEEP_ROLLED _Pragma("unroll 1")
#else
#define CRL_HD inline __attribute__((always_inline))
#define CRL_UNROLL _Pragma("GCC unroll 16")
#define CRL_KEEP_ROLLED _Pragma("GCC unroll 1")
#endif

namespace crl {

// ---- scalar helpers -------------------------------------------------------------------
CRL_HD double fmadd(double a, double b, double c) { return ::fma(a, b, c); }
CRL_HD float fmadd(float a, float b, float c) { return ::fmaf(a, b, c); }

CRL_HD double fabs_(double x);
CRL_HD float fabs_(float x);

// ---- branch-free sin/cos ----------------------------------------------------------------
// The library sincos() carries a data-dependent slow path (Payne-Hanek) behind a branch, which
// fences every call into its own basic block: the 3 ... 12 evaluations a time step needs then
// run as serial dependent chains.  sincos_batch evaluates N angles with straight-line code
// (3-step Cody-Waite reduction by pi/2 with FMA, fdlibm/cephes minimax kernels on |r| <= pi/4,
// quadrant fix-up by selects) so the compiler can interleave them, and falls back to the
// library - once, for the whole batch - only if some |x| exceeds the range where the
// reduction is accurate (or is NaN/Inf).  Max error ~1 ulp (double), ~2 ulp (float): the
// same order as the CUDA / glibc routines it replaces; identical code runs in the host
// test build, so results there match the GPU bit for bit.
CRL_HD int lo_bits(double t) {
#if defined(__CUDA_ARCH__)
    return __double2loint(t);
#else
    long long b;
    __builtin_memcpy(&b, &t, 8);
    return (int)(b & 0xffffffffLL);
#endif
}
CRL_HD int lo_bits(float t) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(t);
#else
    int b;
    __builtin_memcpy(&b, &t, 4);
    return b;
#endif
}

CRL_HD void sincos_core(double x, double* s, double* c) {
    const double t = fmadd(x, 0x1.45f306dc9c883p-1, 6755399441055744.0);   // x*2/pi + 1.5*2^52: rounds to nearest int
    const int q = lo_bits(t);
    const double fn = t - 6755399441055744.0;
    double r = fmadd(fn, -0x1.921fb54442d18p+0, x);
    r = fmadd(fn, -0x1.1a62633145c07p-54, r);
    r = fmadd(fn, 0x1.f1976b7ed8fbcp-110, r);
    const double z = r * r;
    double ps = fmadd(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fmadd(z, ps, 2.75573137070700676789e-06);
    ps = fmadd(z, ps, -1.98412698298579493134e-04);
    ps = fmadd(z, ps, 8.33333333332248946124e-03);
    ps = fmadd(z, ps, -1.66666666666666324348e-01);
    const double sr = fmadd(z * r, ps, r);
    double pc = fmadd(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fmadd(z, pc, -2.75573143513906633035e-07);
    pc = fmadd(z, pc, 2.48015872894767294178e-05);
    pc = fmadd(z, pc, -1.38888888888741095749e-03);
    pc = fmadd(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    const double cr = w + (((1.0 - w) - hz) + z * z * pc);
    const double ss = (q & 1) ? cr : sr;
    const double cc = (q & 1) ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}
CRL_HD void sincos_core(float x, float* s, float* c) {
    const float t = fmadd(x, 0.636619747f, 12582912.0f);                    // 1.5*2^23
    const int q = lo_bits(t);
    const float fn = t - 12582912.0f;
    float r = fmadd(fn, -1.57079637f, x);
    r = fmadd(fn, 4.37113883e-08f, r);
    r = fmadd(fn, 1.71512451e-15f, r);
    const float z = r * r;
    float ps = fmadd(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmadd(z, ps, -1.6666654611e-1f);
    const float sr = fmadd(z * r, ps, r);
    float pc = fmadd(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmadd(z, pc, 4.166664568298827e-2f);
    const float cr = fmadd(z * z, pc, fmadd(z, -0.5f, 1.0f));
    const float ss = (q & 1) ? cr : sr;
    const float cc = (q & 1) ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}
CRL_HD double trig_limit(double) { return 1.0e5; }
CRL_HD float trig_limit(float) { return 4096.0f; }
CRL_HD void sincos_lib(double x, double* s, double* c) { ::sincos(x, s, c); }
CRL_HD void sincos_lib(float x, float* s, float* c) { ::sincosf(x, s, c); }

template <int NANG, class R>
CRL_HD void sincos_batch(const R* x, R* s, R* c) {
    bool in_range = true;
    CRL_UNROLL for (int i = 0; i < NANG; ++i) {
        sincos_core(x[i], &s[i], &c[i]);
        in_range = in_range && (fabs_(x[i]) <= trig_limit(R(0)));            // false for NaN too
    }
    if (!in_range) {
        CRL_UNROLL for (int i = 0; i < NANG; ++i)
            if (!(fabs_(x[i]) <= trig_limit(R(0)))) sincos_lib(x[i], &s[i], &c[i]);
    }
}

CRL_HD double fabs_(double x) { return ::fabs(x); }
CRL_HD float fabs_(float x) { return ::fabsf(x); }

// min(max(lo, v), hi) with the argument order and comparison direction of the reference's
// `min(max(c[0], x), c[1])` (ilqr_dynamic_model.py:108,128,131; numba reproduces CPython's
// "replace the running best only if strictly better" rule, so a NaN x becomes lo).
template <class R>
CRL_HD R clamp_(R v, R lo, R hi) {
    R a = (v > lo) ? v : lo;      // max(lo, v)
    return (hi < a) ? hi : a;     // min(a, hi)
}

// ---- strided views --------------------------------------------------------------------
// One trajectory-sized array seen by one lane.  Element (t, c) lives at
// base[(t * W + c) * SC]:  SC = 1  -> row-major [T][W] owned by this trajectory (API layout),
//                          SC = 32 -> warp tile [T][W][32 lanes], base already offset by lane
//                                     (workspace layout: every warp access is one 32-lane row).
template <class R, int W, int SC>
struct View {
    R* base;
    CRL_HD R* row(int t) const { return base + (size_t)t * (W * SC); }
    CRL_HD static R ld(const R* row, int c) { return row[c * SC]; }
    CRL_HD static void st(R* row, int c, R v) { row[c * SC] = v; }
};

// One trajectory whose rows are contiguous but RS reals apart: element (t, c) at base[t * RS + c].
// The cooperative engine's line-search buffers are [T][candidates][DP] (DP = row length padded to a
// multiple of 32 bytes), so a candidate's row is a run of whole 32-byte sectors: reading the current
// trajectory touches only its own bytes (an element-interleaved layout costs a sector per element).
// `on` = false turns the stores off (a candidate that is evaluated for its cost only).
// PAD0 / PAD1 (optional): storing element PAD0 - 1 also zeroes elements PAD0 .. PAD1 - 1, i.e. the padding behind a row,
// so that every 32-byte sector of the row is written whole (a partly written sector costs the memory system a read of
// the rest when the line leaves the L2; measured on the benchmark solve: 320 -> 306 ms, profiles/r02/ab_pad_rows.txt).
template <class R, int RS, int PAD0 = 0, int PAD1 = 0>
struct RowView {
    R* base;
    bool on = true;
    CRL_HD R* row(int t) const { return base + (size_t)t * RS; }
    CRL_HD static R ld(const R* row, int c) { return row[c]; }
    CRL_HD void st(R* row, int c, R v) const {
        if (on) {
            row[c] = v;
            if (PAD1 > PAD0 && c == PAD0 - 1) {
                CRL_UNROLL for (int p = PAD0; p < PAD1; ++p) row[p] = R(0);
            }
        }
    }
};

}  // namespace crl
