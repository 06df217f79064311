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

// Coefficients whose low 32 bits are zero are encoded by ptxas as 32-bit immediates of DFMA / DMUL: they need neither a
// uniform register nor a constant load.  -1/2, -1/4 are exact; for 1/5 (log1p), 1/24 (exp), 1/120 and 1/720 (sin / cos
// of the residual angle) the value rounded to 20 mantissa bits is enough: they multiply r^5, r^4, r^5 and r^6 with
// |r| < 2^-9, ln2/512 and pi/256, so the coefficient error (5e-7 relative) contributes < 1e-19 (tests/test_fastmath.py).
#define FM_C5 0x1.9999ap-3    /* 1/5   */
#define FM_C24 0x1.55555p-5   /* 1/24  */
#define FM_C120 0x1.11111p-7  /* 1/120 */
#define FM_C720 0x1.6c16cp-10 /* 1/720 */

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
    double q = fma(FM_C5, r, -0.25);                                          // log1p(r), degree 5
    q = fma(q, r, kFM[2]);
    q = fma(q, r, -0.5);
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
    double q = fma(FM_C24, r, kFM[9]);
    q = fma(q, r, 0.5);
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
    double q = fma(FM_C24, r, kFM[9]);         // e^r - 1, degree 4, |r| <= ln2/512
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    double res = fma(T, q * r, T);
    res = __hiloint2double(__double2hiint(res) + ((ki << 12) & 0xfff00000), __double2loint(res));  // * 2^(ki >> 8)
    if (((unsigned)__double2hiint(x) & 0x7fffffffu) >= 0x40862000u) {  // |x| >= 708, inf or NaN: rare
        if (x < 0.0) res = 0.0;
        else res = x + CUDART_INF_F;
    }
    return res;
}

// log N(y; mu, sqrt(v)) given the variance v > 0, with the table-driven log (libdevice log() costs ~60 instructions,
// 22 of them constant materialisations); v outside the normal range (denormal, inf, NaN) takes the libdevice path
__device__ __forceinline__ double gauss_logpdf_var_t(double y, double mu, double v, unsigned lt) {
    const double r = y - mu;
    const unsigned hi = (unsigned)__double2hiint(v);
    const double lg = (hi - 0x00100000u < 0x7fe00000u) ? tlog_pos(v, lt) : log(v);
    const double t = fma(r * r, fast_rcp_pos(v), lg);
    return fma(-0.5, t, -BAM_LOG_SQRT_2PI);
}

// 1/v for a positive normal v: MUFU.RCP64H seed (>= 20 bits) + one cubic step  r0 (1 + e + e^2), e = 1 - v r0
__device__ __forceinline__ double trcp(double v) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(v));
    const double e = fma(-v, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}

// ---------------------------------------------------------------------------------------------------------------
// Structural-error noise of the prediction path: Box-Muller pairs, FOUR normals from one Philox4x32-10 block.
//
// Specification (shared with oracle/bam_oracle.c:oracle_normal_pred and bam_device.cuh:philox_normal_pred): for
// replication `rep`, prediction input t and output y,
//     (w0,w1,w2,w3) = Philox4x32-10(key = seed; counter = (rep_lo, rep_hi, t>>2, y)),   p = (t>>1)&1,
//     u1 = (w[2p] + 0.5) 2^-32,  u2 = (w[2p+1] + 0.5) 2^-32,
//     z(t) = sqrt(-2 ln u1) * (t even ? cos(2 pi u2) : sin(2 pi u2)).
// 32-bit uniforms (|z| <= 6.7) are the resolution class of the generators the reference itself draws from; four
// consecutive inputs share one Philox block, two share one log and one square root.
// FP64-pipe cost per normal: ~2 (uniforms) + 4 (tlog) + 3.5 (sqrt) + 11.5 (sincos) + 0.5 = ~22 (libdevice: ~110).
// ---------------------------------------------------------------------------------------------------------------

// [0..5] sin kernel S1..S6, [6..11] cos kernel C1..C6 (fdlibm k_sin.c / k_cos.c minimax sets, |x| <= pi/4),
// [12] pi, [13] 1.5*2^52 (rint), [14] 2^52 - 0.5
__constant__ double kFN[16] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10,
                               4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                               -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11,
                               3.14159265358979323846,      6755399441055744.0,          4503599627370495.5,
                               0.0};

// w + 0.5 as a double (exact): magic-number add instead of an integer -> double conversion
__device__ __forceinline__ double u32_plus_half(uint32_t w) { return __hiloint2double(0x43300000, (int)w) - kFN[14]; }

// sqrt(x) for a positive normal x: MUFU.RSQ64H seed (>= 20 bits), one Goldschmidt step (>= 40 bits), one
// correction g += (x - g^2) h: result within 1 ulp.  7 FP64 instructions (IEEE sqrt(): ~15 with its slow path).
__device__ __forceinline__ double tsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    const double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}

// sin(pi a), cos(pi a) for a in [0, 2]
__device__ __forceinline__ void tsincospi(double a, double& s, double& c) {
    const double t = fma(a, 2.0, kFN[13]);
    const int q = __double2loint(t);           // rint(2a) in 0..4
    const double kd = t - kFN[13];
    const double x = fma(kd, -0.5, a) * kFN[12];  // pi * (a - q/2), |x| <= pi/4
    const double z = x * x;
    double ps = fma(kFN[5], z, kFN[4]);
    ps = fma(ps, z, kFN[3]);
    ps = fma(ps, z, kFN[2]);
    ps = fma(ps, z, kFN[1]);
    ps = fma(ps, z, kFN[0]);
    const double sn = fma(x * z, ps, x);
    double pc = fma(kFN[11], z, kFN[10]);
    pc = fma(pc, z, kFN[9]);
    pc = fma(pc, z, kFN[8]);
    pc = fma(pc, z, kFN[7]);
    pc = fma(pc, z, kFN[6]);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    // quadrant: angle = q*pi/2 + x
    const double s0 = (q & 1) ? cs : sn, c0 = (q & 1) ? sn : cs;
    const unsigned sneg = (unsigned)q & 2u, cneg = (unsigned)(q + 1) & 2u;
    s = __hiloint2double((int)((unsigned)__double2hiint(s0) ^ (sneg << 30)), __double2loint(s0));
    c = __hiloint2double((int)((unsigned)__double2hiint(c0) ^ (cneg << 30)), __double2loint(c0));
}

// sin / cos of the Box-Muller angle 2 pi (w + 0.5) 2^-32 from the 32-bit word itself: the top 8 bits select the bin
// centre a_j = 2 pi (j + 0.5) / 256 of a (cos, sin) table in shared memory (`st`), the low 24 bits give the residual
// |r| <= pi/256, where sin r and cos r - 1 need 2 and 3 Taylor terms (truncation < 1e-17).  15 FP64 instructions and six
// constants (1/2, 1/6, 1/24 are shared with texp) against 23 and thirteen for the quadrant-reduced minimax kernels:
// with those, the constants of the propagation loop no longer fit the uniform register file and every use cost an
// extra LDC.  ABSOLUTE error ~2e-16 (the result multiplies sqrt(-2 ln u1) <= 6.7).
// [0] 2 pi 2^-32, [1] 2^52 + 2^23, [2] 1/120, [3] 1/720, [4] pi 2^-32 (the +0.5 of the word, scaled)
__constant__ double kFS[5] = {1.4629180792671596e-09, 4503599635759104.0, 1.0 / 120.0, 1.0 / 720.0, 7.314590396335798e-10};
__device__ __forceinline__ void tsincos2pi_word(uint32_t w, unsigned st, double& s, double& c) {
    const double2 cs = lds_f64x2(st + ((w >> 20) & 0xff0u));  // (cos a_j, sin a_j), 16 B per entry
    // r = 2 pi 2^-32 ((w & 0xffffff) - 2^23 + 0.5): the integer part through a magic-number subtract (exact)
    const double r = fma(__hiloint2double(0x43300000, (int)(w & 0x00ffffffu)) - kFS[1], kFS[0], kFS[4]);
    const double r2 = r * r;
    const double sr = fma(r, r2 * fma(r2, FM_C120, -kFM[9]), r);           // sin r
    const double cm = r2 * fma(fma(r2, -FM_C720, kFM[8]), r2, -0.5);       // cos r - 1
    s = cs.y + fma(cs.y, cm, cs.x * sr);
    c = cs.x + fma(cs.x, cm, -(cs.y * sr));
}

// libdevice log() behind a call: it serves one draw in a thousand, and inlined into the unrolled noise code its ~60
// instructions and live registers were paid for (spills at the 64-register cap) by every other draw
static __device__ __noinline__ double log_rare(double x) { return log(x); }

// the Box-Muller pair (z_even, z_odd) from two Philox words; `lt` = shared log table, `st` = shared (cos, sin) table
__device__ __forceinline__ void box_muller_fast(uint32_t wa, uint32_t wb, unsigned lt, unsigned st, double& z0, double& z1) {
    const double n1 = u32_plus_half(wa);
    // u1 = n1 * 2^-32 through the exponent field (integer subtract, exact)
    const double u1 = __hiloint2double(__double2hiint(n1) - (32 << 20), __double2loint(n1));
    double L;
    if (wa >= 0xFFC00000u) {
        // u1 within 2^-10 of 1 (probability 1e-3): the table-driven log has ABSOLUTE error ~1e-16, not enough
        // relative accuracy when log(u1) -> 0
        L = log_rare(u1);
    } else {
        L = tlog_pos(u1, lt);  // |L| >= 9.7e-4
    }
    const double r = tsqrt_pos(-2.0 * L);
    double s, c;
    tsincos2pi_word(wb, st, s, c);
    z0 = r * c;
    z1 = r * s;
}

// the four normals z(4q), z(4q+1), z(4q+2), z(4q+3) of (seed, rep, quad q = t>>2, y)
__device__ __forceinline__ void normal_quad_fast(uint64_t seed, uint64_t rep, uint32_t quad, uint32_t y, unsigned lt,
                                                 unsigned st, double z[4]) {
    uint32_t c0 = (uint32_t)rep, c1 = (uint32_t)(rep >> 32), c2 = quad, c3 = y;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    box_muller_fast(c0, c1, lt, st, z[0], z[1]);
    box_muller_fast(c2, c3, lt, st, z[2], z[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// The sampler's draws (philox_normal / philox_uniform of bam_device.cuh: 53-bit uniforms, z = sqrt(-2 ln u1) cos(2 pi u2),
// accept iff ln u < delta) with the table-driven functions.  Values agree with the libdevice formulation to ~1e-15.
// ---------------------------------------------------------------------------------------------------------------
// log(u), u in (0,1): the table log has ABSOLUTE error ~1e-16, so libdevice takes over within 2^-10 of 1
__device__ __forceinline__ double tlog_unit(double u, unsigned lt) {
    return ((unsigned)__double2hiint(u) >= 0x3fefff80u) ? log(u) : tlog_pos(u, lt);
}
__device__ __forceinline__ double philox_normal_fast(uint64_t seed, uint64_t a, uint32_t b, uint32_t c, unsigned lt) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = c;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u53(c0, c1), u2 = u53(c2, c3);
    double sn, cs;
    tsincospi(2.0 * u2, sn, cs);
    return tsqrt_pos(-2.0 * tlog_unit(u1, lt)) * cs;
}
__device__ __forceinline__ double philox_log_uniform_fast(uint64_t seed, uint64_t a, uint32_t b, uint32_t c, unsigned lt) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = c;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    return tlog_unit(u53(c0, c1), lt);
}

// one normal of the prediction noise stream (bam_device.cuh:philox_normal_pred) with the table-driven functions: for the
// generic prediction kernel, which visits (replication, input, output) one at a time
__device__ __forceinline__ double philox_normal_pred_fast(uint64_t seed, uint64_t rep, uint32_t t, uint32_t y, unsigned lt,
                                                          unsigned st) {
    uint32_t c0 = (uint32_t)rep, c1 = (uint32_t)(rep >> 32), c2 = t >> 2, c3 = y;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    double z0, z1;
    box_muller_fast((t & 2u) ? c2 : c0, (t & 2u) ? c3 : c1, lt, st, z0, z1);
    return (t & 1u) ? z1 : z0;
}
