// bam_baratin_cb2.cuh -- K1 fast path, second generation: BaRatin log-likelihood, thread = chain, PERSISTENT blocks.
//
// Same mapping as the first generation (a block of 256 chains walks a tile of stage-sorted gaugings staged in shared
// memory, every observation read is a broadcast, the partial of a chain stays in registers), re-designed around what
// the ncu profile of that kernel showed (profiles/r1_logpost_v11_ncu.json: FP64 pipe 63.7 %, 46 of 86.6 instructions
// per chain x gauging were integer / LDS / control, 7.4e7 bank conflicts):
//
//  * JOINT (exponent, mantissa) log table: 16 binades x 256 entries {1/c, log c} in shared memory, circular in the
//    exponent (entry = bits 12..23 of the high word), so  log x = log c + log1p(x/c - 1)  needs no mantissa surgery,
//    no exponent conversion (I2F runs on the quarter-rate XU pipe) and no  e ln2  FMA: 7 FP64 + 3 other instructions
//    instead of 8 + 11.  The window [2^(E0-1023), 2^(E0-1007)) is chosen from the stage data at creation.
//  * exp table with the entry's own index pre-subtracted from its high word: T 2^k is ONE integer add (LEA) on the
//    loaded entry, done BEFORE the multiply, so the result is accumulated by the last FMA:  q += T' (1 + r P(r)).
//    9 FP64 (accumulate included) + 4 others instead of 10 + 8.
//  * range tests only at the two ENDS of a group: the gaugings of a group are sorted, (H - b) and c log(H - b) + log a
//    are monotone in H, so min / max of a group are its first and last gauging.
//  * sum of log-variances: the variances of a group are multiplied as they are (4 DMUL per 5 gaugings) and ONE
//    exponent / mantissa split per group feeds the running mantissa product; a per-chain gate on the remnant parameters
//    (and the tested range of the powers) bounds every variance, so the group product can neither overflow nor lose
//    bits.  The gate also proves sigma > 0 for Constant / Linear, which removes the per-gauging sign test.
//  * gaugings staged as padded groups {H[ILP], Y[ILP], uQ^2[ILP]}: 8 LDS.128 broadcasts per 5 gaugings instead of 15
//    LDS.64.
//  * persistent blocks (2 per SM) loop over (tile, chain block) items: the 66 KB of tables are staged once per block,
//    and small chain counts (512 per GPU in the strong-scaling configuration) still fill every SM with small tiles.
//
// Anything the group path cannot prove safe -- a stage boundary inside the group, H - b outside the table window,
// |c log(H-b) + log a| >= 44, a chain whose remnant parameters fail the gate -- goes gauging by gauging through
// cb2_obs_checked (table functions with every test), which itself hands values outside the table window to
// cb2_obs_slow (libdevice, out of line).  All three tiers evaluate the same formulas; they agree to ~1e-15.
#pragma once
#include "bam_fastmath.cuh"
#include "bam_handle.h"

#ifndef BAM_CB2_ILP
#define BAM_CB2_ILP 5   /* gaugings per group */
#endif
#ifndef BAM_CB2_MINB
#define BAM_CB2_MINB 2
#endif
#ifndef BAM_CB2_EXP1C
#define BAM_CB2_EXP1C 0 /* 1: one-constant argument reduction of exp (8 FP64, relative error |x| 1.1e-16) */
#endif
#define CB2_GS ((3 * BAM_CB2_ILP + 1) & ~1)        /* doubles per staged group (multiple of 16 bytes) */
#define CB2_XX_OOR 0x40460000u                      /* high word of 44.0 */
#define CB2_WIN ((unsigned)BAM_LOG2_NB << 20)       /* width of the log window in high-word units */
#define CB2_G_LO 0x3c300000                         /* high word of 2^-60 */
#define CB2_G_HI 0x41300000                         /* high word of 2^20  */

// log(x) for x inside the joint table's window (the CALLER tests the window): 7 FP64, SHF + LOP3 + LDS.128
__device__ __forceinline__ void tlog2_parts(double x, unsigned lt2, double& ty, double& pr) {
    const double2 t = lds_f64x2(lt2 + (((unsigned)__double2hiint(x) >> 8) & 0xfff0u));
    const double r = fma(x, t.x, -1.0);
    double q = fma(FM_C5, r, -0.25);
    q = fma(q, r, kFM[2]);
    q = fma(q, r, -0.5);
    pr = fma(q, r * r, r);
    ty = t.y;
}

// acc + exp(x), |x| < 708 (the CALLER tests the range); `et2` = exp table with pre-subtracted index
__device__ __forceinline__ double texp2_acc(double x, double acc, unsigned et2) {
    const double t = fma(x, kFM[5], kFM[11]);
    const int ki = __double2loint(t);
    const double kd = t - kFM[11];
#if BAM_CB2_EXP1C
    const double r = fma(kd, kFM[6] + kFM[7], x);
#else
    double r = fma(kd, kFM[6], x);
    r = fma(kd, kFM[7], r);
#endif
    const double T0 = lds_f64(et2 + ((ki << 3) & 0x7f8));
    const double T = __hiloint2double(__double2hiint(T0) + (ki << 12), __double2loint(T0));  // 2^(ki/256)
    double q = fma(FM_C24, r, kFM[9]);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    return fma(T, q, acc);
}

struct Cb2Obs {
    double s2;  // squared standardised residual
    double lv;  // log of the total variance (cb2_obs_slow) or the variance itself (cb2_obs_checked)
    int bad;    // sigma <= 0
};

// one gauging with libdevice functions and every test: values outside the table window, overflowing powers
template <int NC, int SIGF>
__device__ __noinline__ Cb2Obs cb2_obs_slow(const double* __restrict__ cr, unsigned mask4, double g1, double g2, double h, double y,
                                            double v2) {
    int r = 0;
#pragma unroll
    for (int j = 0; j < NC; ++j) r += (h >= cr[3 * j]) ? 1 : 0;
    const unsigned m = (mask4 >> (4 * r)) & 0xfu;
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j)
        if ((m >> j) & 1u) {
            const double x = h - cr[3 * NC + j], c = cr[3 * j + 2], la = cr[4 * NC + j];
            q += (x == 0.0) ? pow(0.0, c) * exp(la) : exp(fma(c, log(x), la));
        }
    double sig;
    if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
    else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
    else sig = g1 * fabs(q);
    Cb2Obs o;
    o.bad = !(sig > 0.0) || __double2hiint(sig) < 0x00100000;
    const double var = fma(sig, sig, v2), res = y - q;
    o.s2 = res * res / var;
    o.lv = log(var);
    return o;
}

template <int NC, int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK, BAM_CB2_MINB) logpost_baratin_cb2(DevProblem p, int K, int nItems, int chainBlocks,
                                                                         const int* __restrict__ tStart, const int* __restrict__ tEnd,
                                                                         const int* __restrict__ tCombo, const int* __restrict__ tileIds,
                                                                         const double* __restrict__ tab, const int* status,
                                                                         double* __restrict__ part, int* statusOut) {
    constexpr int ILP = BAM_CB2_ILP;
    constexpr int CH = (SIGF == BAMGPU_SIG_PROPORTIONAL) ? (ILP < 4 ? ILP : 4) : (ILP < 5 ? ILP : 5);  // variances per product
    extern __shared__ __align__(16) double sm[];
    double2* sLog = reinterpret_cast<double2*>(sm);  // BAM_LOG2_N x 16 B
    double* sExp = sm + 2 * BAM_LOG2_N;
    double* sObs = sExp + BAM_EXPTAB_N;              // groups of CB2_GS doubles

    const int tid = threadIdx.x;
    for (int i = tid; i < BAM_LOG2_N; i += BAM_BLOCK) sLog[i] = p.logTab2[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExp[i] = p.expTab2[i];
    const unsigned lt = smem_addr(sLog), et = smem_addr(sExp), ot = smem_addr(sObs);
    const unsigned baseHi = (unsigned)p.logBaseHi;
    // control matrix as nibbles: nibble r = row r-1 (nibble 0 = no control: H below every activation stage)
    unsigned mask4 = 0;
#pragma unroll
    for (int r = 1; r <= NC; ++r) mask4 |= ((unsigned)(p.ctrlMask >> ((r - 1) * 8)) & 0xfu) << (4 * r);

    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int tsel = item / chainBlocks, cb = item - tsel * chainBlocks;
        const int tile = tileIds ? tileIds[tsel] : tsel;
        const int i0 = tStart[tile];
        const int nt = tEnd[tile] - i0;
        __syncthreads();  // the previous item's gaugings are no longer read
        for (int i = tid; i < nt; i += BAM_BLOCK) {
            const int g = i / ILP, u = i - g * ILP;
            const double yu = p.Yu[i0 + i];
            double* grp = sObs + (size_t)g * CB2_GS;
            grp[u] = p.X[i0 + i];
            grp[ILP + u] = p.Y[i0 + i];
            grp[2 * ILP + u] = yu * yu;
        }
        __syncthreads();
        const int c = cb * BAM_BLOCK + tid;
        if (c >= K) continue;
        const size_t o = ((size_t)tile * K + c) * 2;
        part[o + 1] = 0.0;
        if (status[c] != 0) { part[o] = 0.0; continue; }
        const double* __restrict__ row = tab + (size_t)c * p.tabStride;
        const double g1 = row[p.tabGamma], g2 = (SIGF == BAMGPU_SIG_LINEAR) ? row[p.tabGamma + 1] : 0.0;
        const double* __restrict__ cr = row + p.tabCombo + (size_t)tCombo[tile] * p.comboStride;
        double k[NC], b[NC], cc[NC], la[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) { k[j] = cr[3 * j]; cc[j] = cr[3 * j + 2]; b[j] = cr[3 * NC + j]; la[j] = cr[4 * NC + j]; }
        // gate: remnant parameters inside [2^-60, 2^20] (g2 may also be 0) bound every variance of the group path
        bool gate;
        {
            const int h1 = __double2hiint(g1), h2 = __double2hiint(g2);
            gate = h1 >= CB2_G_LO && h1 < CB2_G_HI;
            if (SIGF == BAMGPU_SIG_LINEAR) gate = gate && ((h2 >= CB2_G_LO && h2 < CB2_G_HI) || g2 == 0.0);
        }

        double s2 = 0.0;  // sum of squared standardised residuals
        double pm = 1.0;  // product of the variance mantissas
        int pe = 0;       // sum of their exponents
        int bad = 0;

        auto range = [&](double h) -> int {  // GetHRange (BaRatin_model.f90:398-436): k increasing, range = #{k_j <= H}
            int r = 0;
#pragma unroll
            for (int j = 0; j < NC; ++j) r += (h >= k[j]) ? 1 : 0;
            return r;
        };
        auto next_stage = [&](int r) -> double {  // k[r], or +inf above the last activation stage
            double kn = CUDART_INF_F;
#pragma unroll
            for (int j = NC - 1; j >= 0; --j) kn = (r == j) ? k[j] : kn;
            return kn;
        };
        auto add_var = [&](double var) {  // var > 0 and normal
            const int vh = __double2hiint(var);
            pe += (vh >> 20) - 1023;
            pm *= __hiloint2double((vh & 0x000fffff) | 0x3ff00000, __double2loint(var));
        };
        // one gauging, table functions with every test (BaRatin_computeQ :236-249 for a set that passed ApplyContinuity:
        // the "H-b<0" exits are dead, see the first-generation kernel)
        auto obs_checked = [&](double h, double y, double v2) {
            const unsigned m = (mask4 >> (4 * range(h))) & 0xfu;
            double q = 0.0;
            bool oor = false;
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if ((m >> j) & 1u) {
                    const double x = h - b[j];
                    double ty, pr;
                    tlog2_parts(x, lt, ty, pr);
                    const double xx = fma(cc[j], pr, fma(cc[j], ty, la[j]));
                    oor |= ((unsigned)__double2hiint(x) - baseHi >= CB2_WIN) | (((unsigned)__double2hiint(xx) & 0x7fffffffu) >= CB2_XX_OOR);
                    q = texp2_acc(xx, q, et);
                }
            double sig;
            if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
            else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
            else sig = g1 * fabs(q);
            const double var = fma(sig, sig, v2);
            const unsigned vh = (unsigned)__double2hiint(var);
            if (oor || !gate || __double2hiint(sig) < 0x00100000 || vh - 0x00100000u >= 0x7fe00000u) {
                const Cb2Obs so = cb2_obs_slow<NC, SIGF>(cr, mask4, g1, g2, h, y, v2);
                bad |= so.bad;
                s2 += so.s2;
                // log var = e ln2 + log m: keep the (e, m) bookkeeping, with m = exp(lv - e ln2) in [1, 2)
                const double e = floor(so.lv * 1.4426950408889634);
                pe += (int)e;
                pm *= exp(fma(-e, kFM[4], so.lv));
            } else {
                const double res = y - q;
                s2 = fma(res * res, trcp(var), s2);
                add_var(var);
            }
        };

        const int ngFull = gate ? nt / ILP : 0;
        if (ngFull > 0) {
            int r = range(lds_f64(ot));
            unsigned m = (mask4 >> (4 * r)) & 0xfu;
            double knext = next_stage(r);
            for (int g = 0; g < ngFull; ++g) {
                double w[CB2_GS];
                const unsigned ga = ot + (unsigned)g * (CB2_GS * 8);
#pragma unroll
                for (int q2 = 0; q2 < CB2_GS / 2; ++q2) {
                    const double2 v = lds_f64x2(ga + 16 * q2);
                    w[2 * q2] = v.x;
                    w[2 * q2 + 1] = v.y;
                }
                bool regular = !(w[ILP - 1] >= knext);
                if (SIGF == BAMGPU_SIG_PROPORTIONAL) regular = regular && m != 0;  // Q = 0 makes sigma = 0: flagged gauging by gauging
                double qq[ILP];
                if (regular) {
                    unsigned oorA = 0, oorB = 0;
#pragma unroll
                    for (int u = 0; u < ILP; ++u) qq[u] = 0.0;
#pragma unroll
                    for (int j = 0; j < NC; ++j)
                        if ((m >> j) & 1u) {
                            double xx[ILP];
#pragma unroll
                            for (int u = 0; u < ILP; ++u) {
                                const double x = w[u] - b[j];
                                double ty, pr;
                                tlog2_parts(x, lt, ty, pr);
                                xx[u] = fma(cc[j], pr, fma(cc[j], ty, la[j]));
                                if (u == 0 || u == ILP - 1) {  // sorted gaugings, monotone functions: the ends bound the group
                                    oorA = max(oorA, (unsigned)__double2hiint(x) - baseHi);
                                    oorB = max(oorB, (unsigned)__double2hiint(xx[u]) & 0x7fffffffu);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < ILP; ++u) qq[u] = texp2_acc(xx[u], qq[u], et);
                        }
                    regular = (oorA < CB2_WIN) & (oorB < CB2_XX_OOR);
                }
                if (!regular) {  // rare: stage boundary inside the group, value outside a table
#pragma unroll 1
                    for (int u = 0; u < ILP; ++u) obs_checked(lds_f64(ga + 8 * u), lds_f64(ga + 8 * (ILP + u)), lds_f64(ga + 8 * (2 * ILP + u)));
                    if (w[ILP - 1] >= knext) {
                        r = range(w[ILP - 1]);
                        m = (mask4 >> (4 * r)) & 0xfu;
                        knext = next_stage(r);
                    }
                    continue;
                }
                // Gaussian log-density given Q (GetLogLikelihood :3208-3225) in variance form; sigma > 0 by the gate
                double var[ILP];
#pragma unroll
                for (int u = 0; u < ILP; ++u) {
                    double sig;
                    if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
                    else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(qq[u]), g1);
                    else sig = g1 * fabs(qq[u]);
                    var[u] = fma(sig, sig, w[2 * ILP + u]);
                    const double res = w[ILP + u] - qq[u];
                    s2 = fma(res * res, trcp(var[u]), s2);
                }
#pragma unroll
                for (int u0 = 0; u0 < ILP; u0 += CH) {
                    double pc = var[u0];
#pragma unroll
                    for (int u = u0 + 1; u < u0 + CH && u < ILP; ++u) pc *= var[u];
                    add_var(pc);
                }
            }
        }
#pragma unroll 1
        for (int i = ngFull * ILP; i < nt; ++i) {
            const int g = i / ILP, u = i - g * ILP;
            const unsigned ga = ot + (unsigned)g * (CB2_GS * 8);
            obs_checked(lds_f64(ga + 8 * u), lds_f64(ga + 8 * (ILP + u)), lds_f64(ga + 8 * (2 * ILP + u)));
        }
        // sum_i log N(y_i; q_i, sqrt(var_i)) = -n log sqrt(2 pi) - 0.5 [ sum log var_i + sum res_i^2 / var_i ]
        const double lsum = fma((double)pe, kFM[4], log(pm));
        part[o] = fma(-0.5, lsum + s2, -BAM_LOG_SQRT_2PI * (double)nt);
        if (bad) atomicOr(&statusOut[c], ST_LKH_INFEAS);
    }
}

using Fast2Kernel = void (*)(DevProblem, int, int, int, const int*, const int*, const int*, const int*, const double*, const int*,
                             double*, int*);

template <int NC>
Fast2Kernel pick_fast2_sig(int sigf) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return logpost_baratin_cb2<NC, BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return logpost_baratin_cb2<NC, BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return logpost_baratin_cb2<NC, BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}
