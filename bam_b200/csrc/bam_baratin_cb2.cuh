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
//  * generation 3, TAYLOR SUB-TILES.  The gaugings of a tile are sorted by stage, so a sub-tile of 80 consecutive gaugings
//    spans a short stage interval [Hf, Hl] around its centre H0.  When no activation stage of the chain falls inside it,
//    the rating curve  Q(H) = sum_{j active} a_j (H - b_j)^c_j  is analytic there and
//        Q(H0 + d) = sum_m T_m d^m,   T_m = sum_j a_j (H0-b_j)^c_j binom(c_j, m) (H0-b_j)^-m :
//    ONE degree-8 Horner evaluation (8 DFMA, whatever the number of active controls) replaces one table-driven power per
//    active control (18 FP64 each).  The coefficients cost one power + a 4-instruction recurrence step per degree and
//    control, once per (sub-tile, chain).  Lagrange remainder of control j, relative to its value at H0, for 0 < c <= D+1:
//        |R_j| <= |binom(c_j, D+1)| rho_j^(D+1),   rho_j = max|d| / (Hf - b_j)
//    (the smallest argument of the sub-tile in the denominator); a thread takes the expansion only if that bound is
//    <= 2^-51 for every active control, rho <= 1/8, and the power at H0 is inside both tables.  The decision is per THREAD
//    (a chain's value must not depend on its warp mates); neighbouring chains agree except next to a threshold, where
//    the warp runs both paths.  A thread that declines runs the group path on the same sub-tile.  d = H - H0 is chain
//    independent and staged next to the gaugings.  Sub-tiles start every TSUB groups from the start of a VAR combination,
//    whatever the tile size, so the expansion points do not depend on the tile plan.
//
// Anything the group path cannot prove safe -- a stage boundary inside the group, H - b outside the table window,
// |c log(H-b) + log a| >= 44, a chain whose remnant parameters fail the gate -- goes gauging by gauging through
// cb2_obs_checked (table functions with every test), which itself hands values outside the table window to
// cb2_obs_slow (libdevice, out of line).  All three tiers evaluate the same formulas; they agree to ~1e-15.
#pragma once
#include <type_traits>

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
#ifndef BAM_CB2_COFACTOR
#define BAM_CB2_COFACTOR 1 /* 1: one reciprocal per group of gaugings (common denominator) */
#endif
#ifndef BAM_CB2_TAYLOR
#define BAM_CB2_TAYLOR 1 /* 1: Taylor sub-tile path (generation 3), see below */
#endif
#ifndef BAM_CB2_TDEG
#define BAM_CB2_TDEG 16  /* highest degree of the per-(sub-tile, chain) expansion of the rating curve */
#endif
#define CB2_TDMIN 4      /* lowest degree (the remainder bound needs c <= degree + 1) */
#ifndef BAM_CB2_TSUB
#define BAM_CB2_TSUB 16  /* groups per sub-tile (x BAM_CB2_ILP gaugings) */
#endif
#define CB2_TEPS 0x1p-52                            /* bound on the relative truncation error of one control's power */
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

// ---- staging: packed tiles in global memory, moved by the TMA engine (cp.async.bulk + mbarrier), double buffered ----
// A tile of the plan is stored once per plan (cb2_pack_tiles) exactly as the kernel wants it in shared memory:
//   [ {int nt, int combo, 0, 0} | G groups {H[ILP], Y[ILP], uQ^2[ILP], pad} | G groups {H - H0, Y, uQ^2, pad} | nSub x {H0, max|H-H0|, Hf, Hl} ]
// with G = ceil(nt / ILP) and nSub = ceil(floor(nt / ILP) / TSUB), so staging is ONE bulk copy issued by one thread while
// the block still works on the previous item, and the block needs one barrier per item instead of two.
#define CB2_HDR 2 /* doubles in front of the groups */
__host__ __device__ inline int cb2_tile_doubles(int nt) {
    const int G = (nt + BAM_CB2_ILP - 1) / BAM_CB2_ILP, nSub = (nt / BAM_CB2_ILP + BAM_CB2_TSUB - 1) / BAM_CB2_TSUB;
    return CB2_HDR + 2 * G * CB2_GS + 4 * nSub;
}

static __global__ void cb2_pack_tiles(DevProblem p, int nTiles, const int* __restrict__ tStart, const int* __restrict__ tEnd,
                                      const int* __restrict__ tCombo, const long long* __restrict__ tOff, double* __restrict__ out) {
    constexpr int ILP = BAM_CB2_ILP, TSUB = BAM_CB2_TSUB;
    const int tile = blockIdx.x;
    if (tile >= nTiles) return;
    const int i0 = tStart[tile], nt = tEnd[tile] - i0;
    const int G = (nt + ILP - 1) / ILP, ngAll = nt / ILP;
    double* base = out + tOff[tile];
    double* sObs = base + CB2_HDR;
    double* sTay = sObs + (size_t)G * CB2_GS;
    double* sMeta = sTay + (size_t)G * CB2_GS;
    if (threadIdx.x == 0) {
        int* hd = reinterpret_cast<int*>(base);
        hd[0] = nt; hd[1] = tCombo[tile]; hd[2] = 0; hd[3] = 0;
    }
    for (int i = threadIdx.x; i < G * CB2_GS; i += blockDim.x) { sObs[i] = 0.0; sTay[i] = 0.0; }
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        const int g = i / ILP, u = i - g * ILP;
        const double yu = p.Yu[i0 + i], hh = p.X[i0 + i], yy = p.Y[i0 + i];
        double* grp = sObs + (size_t)g * CB2_GS;
        grp[u] = hh;
        grp[ILP + u] = yy;
        grp[2 * ILP + u] = yu * yu;
        if (g < ngAll) {
            const int sb = g / TSUB;
            const int s0 = sb * TSUB * ILP, s1 = min((sb + 1) * TSUB, ngAll) * ILP;
            const double hf = p.X[i0 + s0], hl = p.X[i0 + s1 - 1];
            const double h0 = 0.5 * (hf + hl);
            double* tg = sTay + (size_t)g * CB2_GS;
            tg[u] = hh - h0;
            tg[ILP + u] = yy;
            tg[2 * ILP + u] = yu * yu;
            if (i == s0) {
                double* me = sMeta + 4 * sb;
                me[0] = h0; me[1] = fmax(h0 - hf, hl - h0); me[2] = hf; me[3] = hl;
            }
        }
    }
}

__device__ __forceinline__ void cb2_mbar_init(unsigned bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void cb2_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cb2_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CB2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CB2_DONE;\n"
        "bra CB2_WAIT;\n"
        "CB2_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

#ifdef BAM_CB2_TRACE
__device__ long long g_cb2Trace[4 * 1024 + 2 * 16384];
__device__ unsigned long long g_cb2Cnt[8];
#define CB2_COUNT(i) atomicAdd(&g_cb2Cnt[i], 1ull)
__device__ __forceinline__ long long cb2_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#else
#define CB2_COUNT(i)
#endif
// sched[0] = next item to hand out (beyond the first one of every block), sched[1] = blocks that left the loop; the last
// block to leave re-arms both, so the counters are zero again when the kernel ends.
template <int NC, int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK, BAM_CB2_MINB) logpost_baratin_cb2(DevProblem p, int K, int nItems, int chainBlocks, int maxDoubles,
                                                                         const double* __restrict__ packed, const long long* __restrict__ tOff,
                                                                         const int* __restrict__ tBytes, const int* __restrict__ tileIds,
                                                                         const double* __restrict__ tab, const int* status,
                                                                         double* __restrict__ part, int* statusOut, int* sched,
                                                                         const int* __restrict__ order, float* __restrict__ cost) {
    constexpr int ILP = BAM_CB2_ILP;
    constexpr int CH = (SIGF == BAMGPU_SIG_PROPORTIONAL) ? (ILP < 4 ? ILP : 4) : (ILP < 5 ? ILP : 5);  // variances per product
    constexpr int TD = BAM_CB2_TDEG, TSUB = BAM_CB2_TSUB;
    extern __shared__ __align__(16) double sm[];
    double2* sLog = reinterpret_cast<double2*>(sm);  // BAM_LOG2_N x 16 B
    double* sExp = sm + 2 * BAM_LOG2_N;
    double* sLog1 = sExp + BAM_EXPTAB_N;             // BAM_LOGTAB_N x 16 B: exponent-free log table for arguments outside the window
    double* sCtl = sLog1 + 2 * BAM_LOGTAB_N;         // 2 mbarriers, 2 item ids
    double* sBuf = sCtl + 4;                         // 2 x maxDoubles: packed tiles

    const int tid = threadIdx.x;
#ifdef BAM_CB2_TRACE
    int trItems = 0;
    if (tid == 0 && blockIdx.x < 1024) { g_cb2Trace[4 * blockIdx.x] = cb2_gtime(); unsigned sid; asm("mov.u32 %0, %smid;" : "=r"(sid)); g_cb2Trace[4 * blockIdx.x + 3] = sid; }
#endif
    for (int i = tid; i < BAM_LOG2_N; i += BAM_BLOCK) sLog[i] = p.logTab2[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExp[i] = p.expTab2[i];
    for (int i = tid; i < BAM_LOGTAB_N; i += BAM_BLOCK) reinterpret_cast<double2*>(sLog1)[i] = p.logTab[i];
    const unsigned lt = smem_addr(sLog), et = smem_addr(sExp), lt1 = smem_addr(sLog1);
    const unsigned bar0 = smem_addr(sCtl), buf0 = smem_addr(sBuf);
    volatile int* sItem = reinterpret_cast<volatile int*>(sCtl + 2);  // item id of each buffer (-1: end of the queue)
    int* sCnt = reinterpret_cast<int*>(sCtl + 3);                     // warps that have left each buffer
    const unsigned baseHi = (unsigned)p.logBaseHi;
    // control matrix as nibbles: nibble r = row r-1 (nibble 0 = no control: H below every activation stage)
    unsigned mask4 = 0;
#pragma unroll
    for (int r = 1; r <= NC; ++r) mask4 |= ((unsigned)(p.ctrlMask >> ((r - 1) * 8)) & 0xfu) << (4 * r);

    // queue position -> item (profile-guided order: expensive items first, see cb2_reorder) -> (chain block, tile): chain
    // block major, so that a block that takes consecutive items mostly keeps its 256 chains and their VAR combination
    const int nTilesLaunch = nItems / chainBlocks;
    auto item_of = [&](int q) -> int { return order ? order[q] : q; };
    auto issue = [&](int item, int slot) {  // one thread: the bulk copy of `item`'s tile into buffer `slot`
        const int tsel = item % nTilesLaunch;
        const int tile = tileIds ? tileIds[tsel] : tsel;
        cb2_bulk_load(buf0 + (unsigned)slot * (unsigned)maxDoubles * 8u, packed + tOff[tile], (unsigned)tBytes[tile], bar0 + 8u * slot);
    };
    // hand out the next queue position into `slot`: item id for the consumers, then the copy (or, at the end of the queue,
    // a plain arrival that completes the phase with the sentinel -1 in place)
    auto produce = [&](int slot) {
        const int q = atomicAdd(&sched[0], 1) + (int)gridDim.x;
        if (q < nItems) {
            const int it = item_of(q);
            sItem[slot] = it;
            issue(it, slot);
        } else {
            sItem[slot] = -1;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * slot) : "memory");
        }
    };
    if (tid == 0) {
        cb2_mbar_init(bar0);
        cb2_mbar_init(bar0 + 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sCnt[0] = 0;
        sCnt[1] = 0;
        const int it0 = item_of(blockIdx.x);  // grid <= nItems
        sItem[0] = it0;
        issue(it0, 0);
        produce(1);
    }
    __syncthreads();
    unsigned parity = 0;  // bit s = phase of buffer s
    // chain parameters of the current (chain block, VAR combination): reloaded only when either changes
    int curCb = -1, curCombo = -1;
    bool live = false, gate = false;
    double g1 = 1.0, g2 = 0.0;
    double k[NC], b[NC], cc[NC], la[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) { k[j] = 0.0; cc[j] = 1.0; b[j] = 0.0; la[j] = 0.0; }
    const double* cr = tab;
    for (int cur = 0;; cur ^= 1) {
        // the warps of a block are coupled only through the two tile buffers: a warp that is done with an item goes on
        // to the next one (already staged); the LAST warp to leave a buffer refills it (arrival counter in shared memory)
        cb2_mbar_wait(bar0 + 8u * cur, (parity >> cur) & 1u);
        parity ^= 1u << cur;
        const int item = sItem[cur];
        if (item < 0) break;
        const long long tItem0 = clock64();
        const int cb = item / nTilesLaunch, tsel = item - cb * nTilesLaunch;
        const int tile = tileIds ? tileIds[tsel] : tsel;
#ifdef BAM_CB2_TRACE
        if (tid == 0 && item < 16384) { g_cb2Trace[4096 + 2 * item] = cb2_gtime(); }
#endif
        const unsigned hb = buf0 + (unsigned)cur * (unsigned)maxDoubles * 8u;
        int nt, combo;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(nt), "=r"(combo) : "r"(hb));
        const int ngAll = nt / ILP;  // full groups
        const unsigned ot = hb + 8u * CB2_HDR;
        const unsigned tt = ot + (unsigned)((nt + ILP - 1) / ILP) * (CB2_GS * 8);
        const unsigned mt = tt + (unsigned)((nt + ILP - 1) / ILP) * (CB2_GS * 8);
        const int c = cb * BAM_BLOCK + tid;
        const size_t o = ((size_t)tile * K + (c < K ? c : 0)) * (p.partPacked ? 1 : 2);
        if (cb != curCb || combo != curCombo) {
            curCb = cb; curCombo = combo;
            // dead lanes (beyond K, or a vector whose prior already failed) stay in the loop for the arrival counter
            live = c < K && status[c] == 0;
            const double* __restrict__ row = tab + (size_t)(live ? c : 0) * p.tabStride;
            cr = row + p.tabCombo + (size_t)combo * p.comboStride;
            g1 = 1.0; g2 = 0.0;
            if (live) {
                g1 = row[p.tabGamma];
                if (SIGF == BAMGPU_SIG_LINEAR) g2 = row[p.tabGamma + 1];
#pragma unroll
                for (int j = 0; j < NC; ++j) { k[j] = cr[3 * j]; cc[j] = cr[3 * j + 2]; b[j] = cr[3 * NC + j]; la[j] = cr[4 * NC + j]; }
            }
            // gate: remnant parameters inside [2^-60, 2^20] (g2 may also be 0) bound every variance of the group path
            const int h1 = __double2hiint(g1), h2 = __double2hiint(g2);
            gate = h1 >= CB2_G_LO && h1 < CB2_G_HI;
            if (SIGF == BAMGPU_SIG_LINEAR) gate = gate && ((h2 >= CB2_G_LO && h2 < CB2_G_HI) || g2 == 0.0);
        }
        if (c < K) {
            if (!p.partPacked) part[o + 1] = 0.0;
            if (!live) part[o] = 0.0;
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
            CB2_COUNT(0);
            const unsigned m = (mask4 >> (4 * range(h))) & 0xfu;
            double q = 0.0;
            bool oor = false;
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if ((m >> j) & 1u) {
                    const double x = h - b[j];
                    const unsigned xh = (unsigned)__double2hiint(x);
                    double ty = 0.0, pr = 0.0;
                    if (xh - baseHi < CB2_WIN) tlog2_parts(x, lt, ty, pr);
                    else if (xh - 0x00100000u < 0x7fe00000u) ty = tlog_pos(x, lt1);  // positive and normal, outside the window (just above an activation stage)
                    else oor = true;                                                  // zero, denormal, negative, inf, NaN
                    const double xx = fma(cc[j], pr, fma(cc[j], ty, la[j]));
                    oor |= (((unsigned)__double2hiint(xx) & 0x7fffffffu) >= BAM_EXP_OOR);
                    q = texp2_acc(xx, q, et);
                }
            double sig;
            if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
            else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
            else sig = g1 * fabs(q);
            const double var = fma(sig, sig, v2);
            const unsigned vh = (unsigned)__double2hiint(var);
            if (oor || !gate || __double2hiint(sig) < 0x00100000 || vh - 0x00100000u >= 0x7fe00000u) {
                CB2_COUNT(1);
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
        // Gaussian log-density of a group given its Q (GetLogLikelihood :3208-3225) in variance form; sigma > 0 by the gate
        auto group_tail = [&](const double (&qq)[ILP], const double* yv, const double* vv) {
            double var[ILP], r2[ILP];
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                double sig;
                if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
                else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(qq[u]), g1);
                else sig = g1 * fabs(qq[u]);
                var[u] = fma(sig, sig, vv[u]);
                const double res = yv[u] - qq[u];
                r2[u] = res * res;
            }
            if (CH == ILP && ILP >= 3 && BAM_CB2_COFACTOR) {
                // ONE reciprocal per group: sum_u r2_u / v_u = [sum_u r2_u prod_{w != u} v_w] / prod_w v_w.  The prefix products
                // are the ones the log-variance sum needs anyway; suffix products and cofactors cost 2 (ILP - 2) + 1 more
                // multiplications against 3 (ILP - 1) saved reciprocal steps.  The gate bounds every variance to
                // [2^-120, 2^172], the data check |Y| <= 2^100 bounds r2, so no product leaves the normal range.
                double pre[ILP], suf[ILP];
                pre[0] = var[0];
#pragma unroll
                for (int u = 1; u < ILP; ++u) pre[u] = pre[u - 1] * var[u];
                suf[ILP - 1] = var[ILP - 1];
#pragma unroll
                for (int u = ILP - 2; u >= 1; --u) suf[u] = suf[u + 1] * var[u];
                double num = r2[0] * suf[1];
#pragma unroll
                for (int u = 1; u < ILP - 1; ++u) num = fma(r2[u], pre[u - 1] * suf[u + 1], num);
                num = fma(r2[ILP - 1], pre[ILP - 2], num);
                s2 = fma(num, trcp(pre[ILP - 1]), s2);
                add_var(pre[ILP - 1]);
            } else {
#pragma unroll
                for (int u = 0; u < ILP; ++u) s2 = fma(r2[u], trcp(var[u]), s2);
#pragma unroll
                for (int u0 = 0; u0 < ILP; u0 += CH) {
                    double pc = var[u0];
#pragma unroll
                    for (int u = u0 + 1; u < u0 + CH && u < ILP; ++u) pc *= var[u];
                    add_var(pc);
                }
            }
        };
        // the group path on groups [gA, gB): one table-driven power per (gauging, active control)
        auto direct_groups = [&](int gA, int gB) {
            int r = range(lds_f64(ot + (unsigned)gA * (CB2_GS * 8)));
            unsigned m = (mask4 >> (4 * r)) & 0xfu;
            double knext = next_stage(r);
            for (int g = gA; g < gB; ++g) {
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
                    CB2_COUNT(2);
                    // The five gaugings together, each with its own control mask and the exponent-free log table (any
                    // positive normal argument): one such group costs ~0.3 us of dependent latency instead of ~2 us for
                    // five serial obs_checked calls -- and every lane meets its stage boundaries in groups of its own,
                    // so a warp pays for each of them separately.
                    bool hard = false;
#pragma unroll
                    for (int u = 0; u < ILP; ++u) qq[u] = 0.0;
                    unsigned mu[ILP], mall = 0;
#pragma unroll
                    for (int u = 0; u < ILP; ++u) {
                        mu[u] = (mask4 >> (4 * range(w[u]))) & 0xfu;
                        mall |= mu[u];
                        if (SIGF == BAMGPU_SIG_PROPORTIONAL) hard |= mu[u] == 0;
                    }
#pragma unroll
                    for (int j = 0; j < NC; ++j)
                      if ((mall >> j) & 1u) {  // controls no gauging of the group uses are skipped
                        double xx[ILP];
#pragma unroll
                        for (int u = 0; u < ILP; ++u) {
                            const bool act = (mu[u] >> j) & 1u;
                            const double x = w[u] - b[j];
                            const bool okx = (unsigned)__double2hiint(x) - 0x00100000u < 0x7fe00000u;  // positive and normal
                            xx[u] = fma(cc[j], tlog_pos((act && okx) ? x : 1.0, lt1), la[j]);
                            hard |= act && (!okx || ((unsigned)__double2hiint(xx[u]) & 0x7fffffffu) >= CB2_XX_OOR);
                        }
#pragma unroll
                        for (int u = 0; u < ILP; ++u) {
                            const double e = texp2_acc(xx[u], 0.0, et);
                            qq[u] += ((mu[u] >> j) & 1u) ? e : 0.0;
                        }
                    }
                    if (!hard) group_tail(qq, &w[ILP], &w[2 * ILP]);
                    else
#pragma unroll 1
                        for (int u = 0; u < ILP; ++u) obs_checked(lds_f64(ga + 8 * u), lds_f64(ga + 8 * (ILP + u)), lds_f64(ga + 8 * (2 * ILP + u)));
                    if (w[ILP - 1] >= knext) {
                        r = range(w[ILP - 1]);
                        m = (mask4 >> (4 * r)) & 0xfu;
                        knext = next_stage(r);
                    }
                    continue;
                }
                group_tail(qq, &w[ILP], &w[2 * ILP]);
            }
        };
        auto checked_groups = [&](int gA, int gB) {  // a chain whose remnant parameters fail the gate
#pragma unroll 1
            for (int i = gA * ILP; i < gB * ILP; ++i) {
                const int g = i / ILP, u = i - g * ILP;
                const unsigned ga = ot + (unsigned)g * (CB2_GS * 8);
                obs_checked(lds_f64(ga + 8 * u), lds_f64(ga + 8 * (ILP + u)), lds_f64(ga + 8 * (2 * ILP + u)));
            }
        };

#if BAM_CB2_TAYLOR
        for (int sb = 0; sb * TSUB < ngAll; ++sb) {
            const int gA = sb * TSUB, gB = min(gA + TSUB, ngAll);
            // ---- expansion of the rating curve about the centre of the sub-tile, and the proof that it may be used
            double T[TD + 1];
#pragma unroll
            for (int q2 = 0; q2 <= TD; ++q2) T[q2] = 0.0;
            bool el = live && gate;
            int deg = CB2_TDMIN;
            if (el) {
                const double2 m01 = lds_f64x2(mt + 32 * sb), m23 = lds_f64x2(mt + 32 * sb + 16);
                const double H0 = m01.x, dmax = m01.y, Hf = m23.x, Hl = m23.y;
                const int r = range(Hf);
                el = !(Hl >= next_stage(r));  // the same stage range from the first to the last gauging
                const unsigned m = (mask4 >> (4 * r)) & 0xfu;
                if (SIGF == BAMGPU_SIG_PROPORTIONAL) el = el && m != 0;
#pragma unroll
                for (int j = 0; j < NC; ++j)
                    if ((m >> j) & 1u) {
                        const double x0 = H0 - b[j], xmin = Hf - b[j];
                        double ty, pr;
                        tlog2_parts(x0, lt, ty, pr);
                        const double xx = fma(cc[j], pr, fma(cc[j], ty, la[j]));
                        const double w0 = texp2_acc(xx, 0.0, et);  // a (H0 - b)^c
                        const double inv = trcp(x0);
                        const double rho = dmax * trcp(xmin);
                        // t_q = w0 binom(c,q) / x0^q (coefficient), s_q = w0 binom(c,q) rho^q (remainder bound if the series
                        // stops before q); the expansion of this control ends at the first q > TDMIN with |s_q| <= eps w0
                        double t = w0, sq = w0;
                        const double thr = CB2_TEPS * w0;
                        int dj = TD + 1;
                        T[0] += t;
#pragma unroll
                        for (int q2 = 1; q2 <= TD + 1; ++q2) {
                            const double gq = fma(cc[j], 1.0 / q2, -(double)(q2 - 1) / q2);  // (c - q + 1) / q
                            sq *= gq * rho;
                            if (q2 > CB2_TDMIN && fabs(sq) <= thr) { dj = q2 - 1; break; }
                            if (q2 <= TD) {
                                t *= gq * inv;
                                T[q2] += t;
                            }
                        }
                        const bool okj = ((unsigned)__double2hiint(x0) - baseHi < CB2_WIN) & (((unsigned)__double2hiint(xx) & 0x7fffffffu) < CB2_XX_OOR) &
                                         (xmin > 0.0) & (rho <= 0.25) & (cc[j] > 0.0) & (cc[j] <= (double)(CB2_TDMIN + 1)) & (dj <= TD);
                        el = el && okj;
                        deg = max(deg, dj);
                    }
            }
            // degree class of the WARP: coefficients above a lane's own degree are exact zeros, so evaluating at a
            // higher degree returns the same bits -- the lanes of a warp can share one loop without depending on each other
            const int degw = __reduce_max_sync(0xffffffffu, el ? deg : 0);
            if (!live) continue;
            CB2_COUNT(el ? 3 : 4);
            if (!el) {  // per THREAD: a chain's value never depends on the chains that share its warp
                if (gate) direct_groups(gA, gB);
                else checked_groups(gA, gB);
                continue;
            }
            // Horner evaluation at the smallest compiled degree that covers `deg` (higher coefficients are exact zeros)
            auto taylor_groups = [&](auto DEG) {
                constexpr int D = decltype(DEG)::value;
                for (int g = gA; g < gB; ++g) {
                    double w[CB2_GS];
                    const unsigned ga = tt + (unsigned)g * (CB2_GS * 8);
#pragma unroll
                    for (int q2 = 0; q2 < CB2_GS / 2; ++q2) {
                        const double2 v = lds_f64x2(ga + 16 * q2);
                        w[2 * q2] = v.x;
                        w[2 * q2 + 1] = v.y;
                    }
                    double qq[ILP];
#pragma unroll
                    for (int u = 0; u < ILP; ++u) qq[u] = fma(T[D], w[u], T[D - 1]);
#pragma unroll
                    for (int q2 = D - 2; q2 >= 0; --q2)
#pragma unroll
                        for (int u = 0; u < ILP; ++u) qq[u] = fma(qq[u], w[u], T[q2]);
                    group_tail(qq, &w[ILP], &w[2 * ILP]);
                }
            };
            if (degw <= 5) taylor_groups(std::integral_constant<int, 5>{});
            else if (degw <= 6) taylor_groups(std::integral_constant<int, 6>{});
            else if (degw <= 8) taylor_groups(std::integral_constant<int, 8>{});
            else if (TD >= 12 && degw <= 12) taylor_groups(std::integral_constant<int, (TD >= 12 ? 12 : TD)>{});
            else taylor_groups(std::integral_constant<int, TD>{});
        }
#else
        if (live) {
            if (gate) direct_groups(0, ngAll);
            else checked_groups(0, ngAll);
        }
#endif
        if (live) {
#pragma unroll 1
            for (int i = ngAll * ILP; i < nt; ++i) {
                const int g = i / ILP, u = i - g * ILP;
                const unsigned ga = ot + (unsigned)g * (CB2_GS * 8);
                obs_checked(lds_f64(ga + 8 * u), lds_f64(ga + 8 * (ILP + u)), lds_f64(ga + 8 * (2 * ILP + u)));
            }
            // sum_i log N(y_i; q_i, sqrt(var_i)) = -n log sqrt(2 pi) - 0.5 [ sum log var_i + sum res_i^2 / var_i ]
            const double lsum = fma((double)pe, kFM[4], log(pm));
            part[o] = fma(-0.5, lsum + s2, -BAM_LOG_SQRT_2PI * (double)nt);
            if (bad) atomicOr(&statusOut[c], ST_LKH_INFEAS);
        }
        __syncwarp();
        if ((tid & 31) == 0) {
            if (atomicAdd(&sCnt[cur], 1) == BAM_BLOCK / 32 - 1) {  // the last warp out: nobody reads this buffer any more
                sCnt[cur] = 0;
                if (cost) cost[item] = (float)(clock64() - tItem0);
                produce(cur);
            }
        }
#ifdef BAM_CB2_TRACE
        ++trItems;
        if (tid == 0 && item < 16384) { g_cb2Trace[4096 + 2 * item + 1] = cb2_gtime(); }
#endif
    }
    __syncthreads();
#ifdef BAM_CB2_TRACE
    if (tid == 0 && blockIdx.x < 1024) { g_cb2Trace[4 * blockIdx.x + 1] = cb2_gtime(); g_cb2Trace[4 * blockIdx.x + 2] = trItems; }
#endif
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {  // every block has fetched its last item
            sched[0] = 0;
            sched[1] = 0;
            __threadfence();
        }
    }
}

// Profile-guided queue order: every launch records what each item cost (clock cycles of the last warp to leave it); from
// time to time this kernel sorts the items by that cost, most expensive first, so that the end of the queue -- where a
// late 200 us item would leave 295 blocks waiting -- is made of cheap, uniform items.  The order never changes a value
// (one partial per (tile, chain), summed in fixed order afterwards).  One block, bitonic sort in shared memory.
static __global__ void __launch_bounds__(1024) cb2_reorder(const float* __restrict__ cost, int n, int* __restrict__ order) {
    extern __shared__ unsigned long long sk[];
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        // key = power-of-two cost class, then the natural (chain block, tile) order: a block that takes consecutive items of
        // a class mostly keeps its chains and their VAR combination (no parameter reload)
        sk[i] = i < n ? (((unsigned long long)(__float_as_uint(fmaxf(cost[i], 1.f)) >> 23) << 32) | (unsigned)(0x7fffffff - i)) : 0ull;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = sk[i], bq = sk[j];
                    if ((a < bq) == desc) { sk[i] = bq; sk[j] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = 0x7fffffff - (int)(unsigned)(sk[i] & 0xffffffffu);
}
#define CB2_REORDER_MAX 16384

using Fast2Kernel = void (*)(DevProblem, int, int, int, int, const double*, const long long*, const int*, const int*, const double*,
                             const int*, double*, int*, int*, const int*, float*);

template <int NC>
Fast2Kernel pick_fast2_sig(int sigf) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return logpost_baratin_cb2<NC, BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return logpost_baratin_cb2<NC, BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return logpost_baratin_cb2<NC, BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}
