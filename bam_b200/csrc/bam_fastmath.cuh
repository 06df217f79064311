// bam_fastmath.cuh -- table-driven double-precision log / exp for the FP64-bound kernels.
//
// Why: the hot loops are bound by the FP64 pipe (64 lanes / SM).  libdevice's log() costs 29 FP64-pipe
// instructions (+ ~30 others, 22 of them UMOVs that re-materialise constants), exp() 16, pow() ~90.  With a
// 128-entry reciprocal table (log) and a 64-entry 2^(j/64) table (exp) staged in shared memory, the polynomial
// degrees drop to 6 and 5:  tlog = 10 FP64 instructions, texp = 12.
//
// Accuracy (validated on the host against long-double libm over 2e7 points, see tests/test_fastmath.py and
// DESIGN.md): tlog |abs error| <= 4e-15 for |log x| <= 20 (<= 5e-16 for |log x| <= 3), texp relative error
// <= 2.4e-16.  Both are far inside the 1e-12 parity budget.  Domain: tlog x > 0 normal, or x == 0 (-> -inf);
// texp any finite x, +-inf; results that would be denormal are flushed to 0.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#define BAM_LOGTAB_N 128
#define BAM_EXPTAB_N 64

__device__ __forceinline__ double tlog(double x, const double2* __restrict__ lt) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const double2 t = lt[(hi >> 13) & (BAM_LOGTAB_N - 1)];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);  // mantissa in [1,2)
    const double r = fma(m, t.x, -1.0);                                       // |r| < 2^-8
    double q = fma(-1.0 / 6.0, r, 0.2);                                       // log1p(r), degree 6
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double pr = fma(q, r * r, r);
    const double s = fma((double)e, 0.693147180559945309417232, t.y);
    const double res = s + pr;
    return (x == 0.0) ? -CUDART_INF_F : res;
}

__device__ __forceinline__ double texp(double x, const double* __restrict__ et) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer in the low word
    const double t = fma(x, 92.332482616893657, MAGIC);      // x * 64/ln2
    const int ki = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -0.010830424667801708, x);            // ln2/64 split hi (28 bits) / lo
    r = fma(kd, -2.8447437476806321e-11, r);
    const double T = et[ki & (BAM_EXPTAB_N - 1)];
    double q = fma(1.0 / 120.0, r, 1.0 / 24.0);              // e^r - 1, degree 5, |r| <= ln2/128
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    double res = fma(T, q * r, T);
    res = __hiloint2double(__double2hiint(res) + ((ki >> 6) << 20), __double2loint(res));  // * 2^(ki/64)
    if (x < -708.0) res = 0.0;
    if (!(x <= 709.0)) res = x + CUDART_INF_F;                  // +big -> inf, NaN -> NaN
    return res;
}
