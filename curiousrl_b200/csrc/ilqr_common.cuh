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
#else
#define CRL_HD inline __attribute__((always_inline))
#define CRL_UNROLL _Pragma("GCC unroll 16")
#endif

namespace crl {

// ---- scalar helpers -------------------------------------------------------------------
CRL_HD double fmadd(double a, double b, double c) { return ::fma(a, b, c); }
CRL_HD float fmadd(float a, float b, float c) { return ::fmaf(a, b, c); }

CRL_HD void sincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
CRL_HD void sincos_(float x, float* s, float* c) { ::sincosf(x, s, c); }

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

}  // namespace crl
