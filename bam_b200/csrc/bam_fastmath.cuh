// bam_fastmath.cuh -- table-driven double-precision log / exp for the FP64-bound kernels.
//
// Why: the hot loops are bound by the FP64 pipe (64 lanes / SM) and, next, by instruction issue.  libdevice's
// log() costs 29 FP64-pipe instructions (+ ~30 others, 22 of them UMOVs that re-materialise constants), exp() 16,
// pow() ~90.  With a 256-entry reciprocal table (log) and a 256-entry 2^(j/256) table (exp) staged in shared
// memory the polynomial degrees drop to 5 and 4:  tlog = 8 FP64 instructions, texp = 9, no FP64 compares.
// Polynomial coefficients come from __constant__ memory so that they are DFMA operands (c[bank][offset]) instead
// of immediates that ptxas materialises with UMOV pairs inside the loop.
//
// Accuracy (host emulation against long-double libm over 2e7 points: tests/test_fastmath.py):
// tlog |abs error| <= 4e-15 for |log x| <= 20 (<= 5e-16 for |log x| <= 3); texp relative error <= 3e-16.
// Domain: tlog x > 0 and normal, or x == 0 (-> -inf).  texp: any x; |x| >= 708 takes a (rare) slow branch.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "bam_device.cuh"  // BAM_LOGTAB_N, BAM_EXPTAB_N

// [0..3] log1p: 1/5, -1/4, 1/3, -1/2   [4] ln2   [5] 256/ln2   [6] -ln2/256 hi   [7] -ln2/256 lo
// [8..10] exp: 1/24, 1/6, 1/2          [11] 1.5*2^52
__constant__ double kFM[12] = {0.2,
                               -0.25,
                               1.0 / 3.0,
                               -0.5,
                               0.693147180559945309417232,
                               369.32993046757463,
                               -0.0027076061523985118,
                               -2.1663774597568432e-11,
                               1.0 / 24.0,
                               1.0 / 6.0,
                               0.5,
                               6755399441055744.0};

// The tables live in shared memory and are addressed with 32-bit shared-window addresses (ld.shared): pointer
// arithmetic on generic pointers made ptxas rebuild the shared base (S2UR/ULEA) in every loop iteration.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_f64x2(unsigned a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

// log(x), x > 0 and normal (no zero / denormal / inf handling): 8 FP64 + ~6 integer instructions + 1 LDS.128.
// `lt` = shared address of the BAM_LOGTAB_N x {1/c_i, -log(1/c_i)} table.
__device__ __forceinline__ double tlog_pos(double x, unsigned lt) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double2 t = lds_f64x2(lt + ((hi >> 8) & 0xff0));                   // entry (hi >> 12) & 255, 16 B each
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);  // mantissa in [1,2)
    const double r = fma(m, t.x, -1.0);                                       // |r| < 2^-9
    double q = fma(kFM[0], r, kFM[1]);                                        // log1p(r), degree 5
    q = fma(q, r, kFM[2]);
    q = fma(q, r, kFM[3]);
    const double pr = fma(q, r * r, r);
    const double s = fma((double)((hi >> 20) - 1023), kFM[4], t.y);
    return s + pr;
}

// log(x), x >= 0: as tlog_pos, plus log(0) = -inf through an integer test (keeps the FP64 pipe free of compares)
__device__ __forceinline__ double tlog(double x, unsigned lt) {
    double res = tlog_pos(x, lt);
    if ((__double2hiint(x) | __double2loint(x)) == 0) res = -CUDART_INF_F;
    return res;
}

// exp(x) for |x| < 708 WITHOUT the range test (the caller tracks max |x| and falls back to texp)
__device__ __forceinline__ double texp_inrange(double x, unsigned et) {
    const double t = fma(x, kFM[5], kFM[11]);
    const int ki = __double2loint(t);
    const double kd = t - kFM[11];
    double r = fma(kd, kFM[6], x);
    r = fma(kd, kFM[7], r);
    const double T = lds_f64(et + ((ki << 3) & 0x7f8));
    double q = fma(kFM[8], r, kFM[9]);
    q = fma(q, r, kFM[10]);
    q = fma(q, r, 1.0);
    const double res = fma(T, q * r, T);
    return __hiloint2double(__double2hiint(res) + ((ki << 12) & 0xfff00000), __double2loint(res));
}
#define BAM_EXP_OOR 0x40862000u /* high word of 708.0: |x| at or above needs the checked texp */

// exp(x); `et` = shared address of the BAM_EXPTAB_N x 2^(j/N) table
__device__ __forceinline__ double texp(double x, unsigned et) {
    const double t = fma(x, kFM[5], kFM[11]);  // x * 256/ln2, rounded to an integer in the low word
    const int ki = __double2loint(t);
    const double kd = t - kFM[11];
    double r = fma(kd, kFM[6], x);             // ln2/256 split hi (26 bits) / lo
    r = fma(kd, kFM[7], r);
    const double T = lds_f64(et + ((ki << 3) & 0x7f8));
    double q = fma(kFM[8], r, kFM[9]);         // e^r - 1, degree 4, |r| <= ln2/512
    q = fma(q, r, kFM[10]);
    q = fma(q, r, 1.0);
    double res = fma(T, q * r, T);
    res = __hiloint2double(__double2hiint(res) + ((ki << 12) & 0xfff00000), __double2loint(res));  // * 2^(ki >> 8)
    if (((unsigned)__double2hiint(x) & 0x7fffffffu) >= 0x40862000u) {  // |x| >= 708, inf or NaN: rare
        if (x < 0.0) res = 0.0;
        else res = x + CUDART_INF_F;
    }
    return res;
}

// 1/v for a positive normal v: MUFU.RCP64H seed (>= 20 bits) + one cubic step  r0 (1 + e + e^2), e = 1 - v r0
__device__ __forceinline__ double trcp(double v) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(v));
    const double e = fma(-v, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}
