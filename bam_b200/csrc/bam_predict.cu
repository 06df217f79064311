// bam_predict.cu -- K4: propagation of MCMC samples through the model + streamed envelopes.
//
// Replaces the replication loop of BaM_Prediction (src/BaM_tools.f90:1034-1108) and the envelope pass
// of BaM_ProcessSpag (:1393-1426).  The reference writes every simulated value to a text file, reads
// the whole file back, and quicksorts each column; here
//
//   pred_prep   one thread per replication: theta (from the sample, or the maxpost mode) with FIX
//               filled in, the per-parameter-set model pre-pass, gamma of the replication -> table.
//   pred_main   thread <-> replication (its parameters sit in a private shared-memory column, loaded
//               once), loop over the prediction inputs of the tile (inputs broadcast).  Model, optional
//               structural-error noise N(0, sigma(gamma, Y)) from a counter-based Philox stream keyed
//               by (seed, rep, t, y).  Consecutive lanes = consecutive replications, so the stores to
//               the tile V[t][rep] are coalesced.  Only one tile of inputs is ever materialised.
//   envelope_*  per prediction input: drop infeasible / NaN values when they are < 75 %, exact
//               empirical quantiles (0.5, 0.025, 0.975, 0.16, 0.84), mean, standard deviation.
//               Small Nrep: the column is sorted in shared memory (bitonic).  Large Nrep: exact
//               selection by histograms over the HBM tile (2-read, 4-read, then MSB-first radix; no full sort).
#include <algorithm>
#include <cfloat>

#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_sediment_fast.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"

namespace {

struct PredArgs {
    long long Nrep;
    int nobsPred, t0, nT;       // this tile: inputs [t0, t0+nT)
    int tBase;                  // first input of the whole call (per-input scratch is indexed from it)
    int doParametric;
    int doRemnant[BAM_MAXY];
    unsigned long long seed;
    int stride;                 // doubles per replication in the table: [gamma | full theta | derived]
    int nStateOut;              // state variables written after the nY outputs (0 | model_nstate)
    int stage;                  // pred_main: stage the table rows of the block in shared memory (stride * 2 KB fits)
    long long repBase;          // global index of local replication 0 (sample-sharded prediction; keys the noise and the input spaghetti)
};

__global__ void __launch_bounds__(128) pred_prep(DevProblem p, PredArgs a, const double* __restrict__ mc,
                                                 const double* __restrict__ mode, double* const* __restrict__ xspag,
                                                 const int* __restrict__ nspag, double* __restrict__ tab,
                                                 int* __restrict__ status) {
    const long long rep = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= a.Nrep) return;
    double* row = tab + (size_t)rep * a.stride;
    // gamma always comes from the sample, even when !doParametric (reference quirk, :1085)
    for (int g = 0; g < p.nGamma; ++g) row[g] = mc ? mc[rep + (size_t)(p.nFit + g) * a.Nrep] : 0.0;
    double th[BAM_MAXTHETA], der[BAM_MAXDER];
    for (int i = 0; i < p.ntheta; ++i) {
        const int src = p.comboSrc[i];
        if (src < 0) th[i] = p.fixv[i];
        else th[i] = a.doParametric ? mc[rep + (size_t)src * a.Nrep] : mode[src];
        row[p.nGamma + i] = th[i];
    }
    for (int d = 0; d < p.nDer; ++d) der[d] = p.unfeas;
    // whole-series models scan the complete input series of this replication (column rep mod nspag, :1049)
    const double* xc[BAM_MAXX];
    for (int j = 0; j < p.nX; ++j) xc[j] = xspag[j] + (size_t)((rep + a.repBase) % nspag[j]) * a.nobsPred;
    const bool ok = model_prepare(p, th, der, xc, a.nobsPred);
    if (p.isSeries) der[0] = (double)rep;  // this replication's column of the series buffer
    for (int d = 0; d < p.nDer; ++d) row[p.nGamma + p.ntheta + d] = der[d];
    status[rep] = ok ? 0 : 1;
}

template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(BAM_BLOCK) pred_main(DevProblem p, PredArgs a, const double* __restrict__ tab,
                                                       const int* __restrict__ status, double* const* __restrict__ xspag,
                                                       const int* __restrict__ nspag, double* __restrict__ V) {
    extern __shared__ __align__(16) double sm[];  // [log table | (cos, sin) table | exp table | [stride][BAM_BLOCK] : parameter k of thread t at sm[k*BLOCK + t]]
    double2* sLogT = reinterpret_cast<double2*>(sm);
    double2* sScT = reinterpret_cast<double2*>(sm + 2 * BAM_LOGTAB_N);
    double* sExpT = sm + 2 * BAM_LOGTAB_N + 2 * BAM_SCTAB_N;
    constexpr bool FM = MODEL == BAMGPU_MDL_VEGETATION || MODEL == BAMGPU_MDL_TEXTFILE;  // table-driven log / exp inside the model
    for (int q = threadIdx.x; q < BAM_LOGTAB_N; q += BAM_BLOCK) sLogT[q] = p.logTab[q];
    for (int q = threadIdx.x; q < BAM_SCTAB_N; q += BAM_BLOCK) sScT[q] = p.scTab[q];
    if (FM)
        for (int q = threadIdx.x; q < BAM_EXPTAB_N; q += BAM_BLOCK) sExpT[q] = p.expTab[q];
    __syncthreads();
    const unsigned lt = smem_addr(sLogT), sct = smem_addr(sScT);
    const FmTab fm = FM ? FmTab{lt, smem_addr(sExpT)} : FmTab{0u, 0u};
    double* const smT = sm + 2 * BAM_LOGTAB_N + 2 * BAM_SCTAB_N + BAM_EXPTAB_N;
    const int tid = threadIdx.x;
    const long long rep = (long long)blockIdx.x * BAM_BLOCK + tid;
    const bool act = rep < a.Nrep;
    const int stride = a.stride;
    // a.stage: the block's table rows in shared memory, ONE ROW PER THREAD with an odd stride (conflict-free when the lanes
    // read the same entry of their own rows): the model reads its parameters straight from there.  A private local-memory
    // copy (up to 224 doubles per thread, dynamically indexed by the growth / VM code) cost more than the model's arithmetic.
    // a.stage == 0 (stride too large for shared memory, e.g. Mixture with its 148 weights): private copy from the table row.
    const int strideP = stride | 1;
    if (act && a.stage)
        for (int k = 0; k < stride; ++k) smT[(size_t)tid * strideP + k] = tab[(size_t)rep * stride + k];
    const bool setOk = act ? (status[rep] == 0) : false;
    const int tBeg = blockIdx.y * 32, tEnd = min(a.nT, tBeg + 32);
    const double* xp[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const int col = act ? (int)((rep + a.repBase) % nspag[j]) : 0;  // indx = modulo(rep, nspag) with 1-based rep (:1049)
        xp[j] = xspag[j] + (size_t)col * a.nobsPred + a.t0;
    }
    if (!act) return;
    double locBuf[BAM_MAXTHETA + BAM_MAXDER];
    const int np = p.ntheta + p.nDer;
    const double* loc = locBuf;
    if (a.stage) loc = smT + (size_t)tid * strideP + p.nGamma;
    else for (int k = 0; k < np; ++k) locBuf[k] = tab[(size_t)rep * stride + p.nGamma + k];
    for (int t = tBeg; t < tEnd; ++t) {
        constexpr int NS = model_nstate(MODEL);
        double xin[NX], y[NY], st[NS > 0 ? NS : 1];
#pragma unroll
        for (int j = 0; j < NX; ++j) xin[j] = xp[j][t];
        bool ok = setOk;
        if (ok) ok = model_apply_s<MODEL, NX, NY>(p, xin, loc, loc + p.ntheta, y, st, fm);
        if (NS > 0 && a.nStateOut > 0)
#pragma unroll
            for (int j = 0; j < NS; ++j) V[((size_t)(NY + j) * a.nT + t) * a.Nrep + rep] = ok ? st[j] : p.unfeas;
#pragma unroll
        for (int j = 0; j < NY; ++j) {
            double v = ok ? y[j] : p.unfeas;  // ApplyModel flags infeasible rows (ModelLibrary_tools.f90:419-432)
            if (a.doRemnant[j] && v != p.unfeas) {
                double gam[3];
                for (int m = 0; m < p.nrem[j] && m < 3; ++m)
                    gam[m] = a.stage ? smT[(size_t)tid * strideP + p.gamOff[j] + m] : tab[(size_t)rep * stride + p.gamOff[j] + m];
                const double sig = sigmafunk(p.sigfunc[j], gam, v);
                if (sig > 0.0) v = v + sig * philox_normal_pred_fast(a.seed, (uint64_t)(rep + a.repBase), (unsigned)(a.t0 + t), (unsigned)j, lt, sct);
            }
            V[((size_t)j * a.nT + t) * a.Nrep + rep] = v;
        }
    }
}

// ---- K4 fast path: BaRatin spaghetti (cfg 5) -------------------------------------------------------------------
//
// thread <-> replication with its parameter set (k, b, c, log a per control, gamma) in REGISTERS; the block walks a
// tile of PF_TT prediction inputs (staged once in shared memory when all replications share one input series).
// a_i (H-b_i)^c_i = texp(c_i tlog(H-b_i) + log a_i) with the table-driven functions of bam_fastmath.cuh, the
// structural-error noise from Box-Muller pairs (one Philox block per FOUR inputs, one log and one sqrt per two).
// As in the log-posterior fast path the "H-b<0" exits of BaRatin_computeQ cannot fire for a parameter set that
// passed ApplyContinuity, so Q is the plain sum over the active controls of the stage range.
// ~60 FP64-pipe instructions per (replication, input) against ~330 for the libdevice formulation of pred_main.
#define PF_TT 64
#ifndef PF_MINB
#define PF_MINB 4
#endif
#define PF_SMEM (sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N + 2 * BAM_SCTAB_N + PF_TT))
template <int NC, int SIGF, bool NOISE>
__global__ void __launch_bounds__(BAM_BLOCK, PF_MINB) pred_baratin_fast(DevProblem p, PredArgs a, const double* __restrict__ tab,
                                                                  const int* __restrict__ status,
                                                                  const double* __restrict__ xs, int nspag,
                                                                  double* __restrict__ V) {
    extern __shared__ __align__(16) double pfsm[];
    double2* sLogT = reinterpret_cast<double2*>(pfsm);
    double* sExpT = pfsm + 2 * BAM_LOGTAB_N;
    double2* sScT = reinterpret_cast<double2*>(sExpT + BAM_EXPTAB_N);
    double* sH = sExpT + BAM_EXPTAB_N + 2 * BAM_SCTAB_N;
    const int tid = threadIdx.x;
    const int tBeg = blockIdx.y * PF_TT, nt = min(PF_TT, a.nT - tBeg);
    for (int i = tid; i < BAM_LOGTAB_N; i += BAM_BLOCK) sLogT[i] = p.logTab[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExpT[i] = p.expTab[i];
    if (NOISE)
        for (int i = tid; i < BAM_SCTAB_N; i += BAM_BLOCK) sScT[i] = p.scTab[i];
    if (nspag == 1)
        for (int i = tid; i < nt; i += BAM_BLOCK) sH[i] = xs[a.t0 + tBeg + i];
    __syncthreads();
    const long long rep = (long long)blockIdx.x * BAM_BLOCK + tid;
    if (rep >= a.Nrep) return;
    double* out = V + (size_t)tBeg * a.Nrep + rep;
    if (status[rep] != 0) {  // infeasible parameter set: the whole replication is flagged (ModelLibrary_tools.f90:419-432)
        for (int t = 0; t < nt; ++t) __stcs(out + (size_t)t * a.Nrep, p.unfeas);
        return;
    }
    const double* __restrict__ row = tab + (size_t)rep * a.stride;
    const double g1 = row[0], g2 = (SIGF == BAMGPU_SIG_LINEAR) ? row[1] : 0.0;
    const double* __restrict__ th = row + p.nGamma;
    double k[NC], b[NC], cc[NC], la[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) { k[j] = th[3 * j]; cc[j] = th[3 * j + 2]; b[j] = th[3 * NC + j]; la[j] = th[4 * NC + j]; }
    unsigned mask4 = 0;
#pragma unroll
    for (int r = 1; r <= NC; ++r) mask4 |= ((unsigned)(p.ctrlMask >> ((r - 1) * 8)) & 0xfu) << (4 * r);
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT), ht = smem_addr(sH), st = smem_addr(sScT);
    const double* __restrict__ xcol = xs + (size_t)((rep + a.repBase) % nspag) * a.nobsPred + a.t0 + tBeg;
    const bool shared = nspag == 1;

    auto stage = [&](int t) -> double { return shared ? lds_f64(ht + 8u * (unsigned)t) : __ldg(xcol + t); };
    auto amask = [&](double h) -> unsigned {  // GetHRange (k increasing) -> row of the control matrix
        int r = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) r += (h >= k[j]) ? 1 : 0;
        return (mask4 >> (4 * r)) & 0xfu;
    };
    auto q_checked = [&](double h, unsigned m) -> double {
        double q = 0.0;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if ((m >> j) & 1u) q += texp(fma(cc[j], tlog(h - b[j], lt), la[j]), et);
        return q;
    };
    auto finish = [&](int t, double q, double z) {
        double v = q;
        if (NOISE) {
            double sig;
            if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
            else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
            else sig = g1 * fabs(q);
            if (sig > 0.0) v = fma(sig, z, q);
        }
        __stcs(out + (size_t)t * a.Nrep, v);
    };
    // two inputs per iteration: they share one Box-Muller pair and give the FP64 pipe two independent chains
    auto two = [&](int t, double zA, double zB) {
        const double hA = stage(t), hB = stage(t + 1);
        const unsigned mA = amask(hA), mB = amask(hB);
        double qA = 0.0, qB = 0.0;
        unsigned oor = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if ((mA >> j) & 1u) {
                const double x = fma(cc[j], tlog(hA - b[j], lt), la[j]);
                oor = max(oor, (unsigned)__double2hiint(x) & 0x7fffffffu);
                qA += texp_inrange(x, et);
            }
            if ((mB >> j) & 1u) {
                const double x = fma(cc[j], tlog(hB - b[j], lt), la[j]);
                oor = max(oor, (unsigned)__double2hiint(x) & 0x7fffffffu);
                qB += texp_inrange(x, et);
            }
        }
        if (oor >= BAM_EXP_OOR) { qA = q_checked(hA, mA); qB = q_checked(hB, mB); }  // rare: H == b or overflow
        finish(t, qA, zA);
        finish(t + 1, qB, zB);
    };
    const unsigned tg0 = (unsigned)(a.t0 + tBeg);
    double z[4] = {0.0, 0.0, 0.0, 0.0};
    auto single = [&](int t) {  // head / tail of a tile that does not start or end on a multiple of four inputs
        const unsigned tg = tg0 + (unsigned)t;
        if (NOISE) normal_quad_fast(a.seed, (uint64_t)(rep + a.repBase), tg >> 2, 0u, lt, st, z);
        const double h = stage(t);
        const double zz = (tg & 2u) ? ((tg & 1u) ? z[3] : z[2]) : ((tg & 1u) ? z[1] : z[0]);
        finish(t, q_checked(h, amask(h)), zz);
    };
    int t = 0;
    for (; t < nt && ((tg0 + (unsigned)t) & 3u); ++t) single(t);
    for (; t + 3 < nt; t += 4) {  // four inputs share one Philox block, two share one log and one sqrt
        if (NOISE) normal_quad_fast(a.seed, (uint64_t)(rep + a.repBase), (tg0 + (unsigned)t) >> 2, 0u, lt, st, z);
        two(t, z[0], z[1]);
        two(t + 2, z[2], z[3]);
    }
    for (; t < nt; ++t) single(t);
}

using PredFastKernel = void (*)(DevProblem, PredArgs, const double*, const int*, const double*, int, double*);

template <int NC, int SIGF>
PredFastKernel pick_pred_fast_noise(bool noise) {
    return noise ? pred_baratin_fast<NC, SIGF, true> : pred_baratin_fast<NC, SIGF, false>;
}
template <int NC>
PredFastKernel pick_pred_fast_sig(int sigf, bool noise) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return pick_pred_fast_noise<NC, BAMGPU_SIG_CONSTANT>(noise);
        case BAMGPU_SIG_LINEAR: return pick_pred_fast_noise<NC, BAMGPU_SIG_LINEAR>(noise);
        case BAMGPU_SIG_PROPORTIONAL: return pick_pred_fast_noise<NC, BAMGPU_SIG_PROPORTIONAL>(noise);
    }
    return nullptr;
}
PredFastKernel pick_pred_fast(const DevProblem& p, bool noise, long long Nrep) {
    if (p.model != BAMGPU_MDL_BARATIN || Nrep < 1024) return nullptr;
    switch (p.ncontrol) {
        case 1: return pick_pred_fast_sig<1>(p.sigfunc[0], noise);
        case 2: return pick_pred_fast_sig<2>(p.sigfunc[0], noise);
        case 3: return pick_pred_fast_sig<3>(p.sigfunc[0], noise);
        case 4: return pick_pred_fast_sig<4>(p.sigfunc[0], noise);
    }
    return nullptr;
}

// ---- K4 fast path: Sediment (cfg 4) -----------------------------------------------------------------------------
//
// When the pseudo-parameters d, s, nu, g are inputs or fixed values and the critical-shear-stress coefficients are
// fixed (the shipped BaM_Sediment* option files), everything in Sediment_Apply except three numbers per replication
// depends on the prediction INPUT only:
//     MPM      dim t1 (tau-css)^t2                  Camenen  dim t1 tau^t2 exp(-t3 css/tau)
//     Nielsen  dim t1 tau^t2 (tau-css)              Smart    dim t1 S0^t2 (s-1)^(1/6) d^(1/6) g^(-1/2) tau^t3 (tau-css)
// all have the form  t1 * B(input) * exp(t2 * L1(input) + t3 * L2(input)).  `sed_obs_prep` evaluates (B, L1, L2) and
// the row status (infeasible inputs / tau <= css -> 0) once per input with the reference's libm-class functions
// (SedimentTransport_model.f90:155-187,290-304); `pred_sediment_fast` then spends one table-driven exp per
// (replication, input) instead of three pow, two exp, a cube root and a square root: ~330 -> ~12 FP64 instructions
// for the model (plus ~24 for the noise).
template <int SIGF, bool NOISE>
__global__ void __launch_bounds__(BAM_BLOCK, 4) pred_sediment_fast(DevProblem p, PredArgs a, const double* __restrict__ mc,
                                                                   const double* __restrict__ mode,
                                                                   const SedObs* __restrict__ obs, double* __restrict__ V) {
    extern __shared__ __align__(16) double pfsm[];
    double2* sLogT = reinterpret_cast<double2*>(pfsm);
    double* sExpT = pfsm + 2 * BAM_LOGTAB_N;
    double2* sScT = reinterpret_cast<double2*>(sExpT + BAM_EXPTAB_N);
    SedObs* sObs = reinterpret_cast<SedObs*>(sExpT + BAM_EXPTAB_N + 2 * BAM_SCTAB_N);
    const int tid = threadIdx.x;
    const int tBeg = blockIdx.y * PF_TT, nt = min(PF_TT, a.nT - tBeg);
    for (int i = tid; i < BAM_LOGTAB_N; i += BAM_BLOCK) sLogT[i] = p.logTab[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExpT[i] = p.expTab[i];
    for (int i = tid; i < BAM_SCTAB_N; i += BAM_BLOCK) sScT[i] = p.scTab[i];
    for (int i = tid; i < nt; i += BAM_BLOCK) sObs[i] = obs[a.t0 - a.tBase + tBeg + i];
    __syncthreads();
    const long long rep = (long long)blockIdx.x * BAM_BLOCK + tid;
    if (rep >= a.Nrep) return;
    double* out = V + (size_t)tBeg * a.Nrep + rep;
    // theta (FIX filled in, :1054-1063) and gamma of this replication, straight from the sample matrix
    double th[3];
    const int nbase = (p.sedFormula == BAMGPU_SED_MPM || p.sedFormula == BAMGPU_SED_NIELSEN) ? 2 : 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        th[i] = 0.0;
        if (i < nbase) {
            const int src = p.comboSrc[i];
            th[i] = (src < 0) ? p.fixv[i] : (a.doParametric ? mc[rep + (size_t)src * a.Nrep] : mode[src]);
        }
    }
    const double e1 = th[1], e2 = (p.sedFormula == BAMGPU_SED_CAMENEN || p.sedFormula == BAMGPU_SED_SMART) ? th[2] : 0.0;
    double g1 = 0.0, g2 = 0.0;
    if (NOISE) {
        g1 = mc[rep + (size_t)p.nFit * a.Nrep];
        if (SIGF == BAMGPU_SIG_LINEAR) g2 = mc[rep + (size_t)(p.nFit + 1) * a.Nrep];
    }
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT), st = smem_addr(sScT);
    const unsigned tg0 = (unsigned)(a.t0 + tBeg);
    double z[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned lastQuad = 0xffffffffu;
    for (int t = 0; t < nt; ++t) {
        const SedObs o = sObs[t];
        const unsigned tg = tg0 + (unsigned)t;
        if (NOISE && (tg >> 2) != lastQuad) {
            lastQuad = tg >> 2;
            normal_quad_fast(a.seed, (uint64_t)(rep + a.repBase), lastQuad, 0u, lt, st, z);
        }
        double v;
        if (o.st == 2.0) v = p.unfeas;
        else {
            const double q = (o.st == 1.0) ? 0.0 : o.B * th[0] * texp(fma(e1, o.L1, e2 * o.L2), et);
            v = q;
            if (NOISE) {
                double sig;
                if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
                else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
                else sig = g1 * fabs(q);
                const double zz = (tg & 2u) ? ((tg & 1u) ? z[3] : z[2]) : ((tg & 1u) ? z[1] : z[0]);
                if (sig > 0.0) v = fma(sig, zz, q);
            }
        }
        __stcs(out + (size_t)t * a.Nrep, v);
    }
}

using SedFastKernel = void (*)(DevProblem, PredArgs, const double*, const double*, const SedObs*, double*);

SedFastKernel pick_sed_fast(int sigf, bool noise) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return noise ? pred_sediment_fast<BAMGPU_SIG_CONSTANT, true> : pred_sediment_fast<BAMGPU_SIG_CONSTANT, false>;
        case BAMGPU_SIG_LINEAR: return noise ? pred_sediment_fast<BAMGPU_SIG_LINEAR, true> : pred_sediment_fast<BAMGPU_SIG_LINEAR, false>;
        case BAMGPU_SIG_PROPORTIONAL: return noise ? pred_sediment_fast<BAMGPU_SIG_PROPORTIONAL, true> : pred_sediment_fast<BAMGPU_SIG_PROPORTIONAL, false>;
    }
    return nullptr;
}

// ---- envelopes ------------------------------------------------------------------------------------

__device__ __forceinline__ double hazen(const double* xs, long long n, double pr) {
    // GetEmpiricalQuantile (BMSL, un-vendored): declared convention = Hazen plotting positions
    const double hh = pr * (double)n + 0.5;
    if (hh <= 1.0) return xs[0];
    if (hh >= (double)n) return xs[n - 1];
    const long long i = (long long)floor(hh);
    const double frac = hh - (double)i;
    return xs[i - 1] + frac * (xs[i] - xs[i - 1]);
}

// block-wide deterministic sum (fixed tree)
__device__ double block_sum(double v, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double s = 0.0;
    if (tid == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += scratch[i];
    if (tid == 0) scratch[0] = s;
    __syncthreads();
    s = scratch[0];
    __syncthreads();
    return s;
}

// one block per column (input t, output y); column fits in shared memory (npad = pow2 >= Nrep)
__global__ void __launch_bounds__(BAM_BLOCK) envelope_smem(long long Nrep, int npad, int nT, double unfeas,
                                                           const double* __restrict__ V, double* __restrict__ env) {
    extern __shared__ double s[];  // npad values + 32 scratch
    double* scratch = s + npad;
    __shared__ int nbadS;
    const int tid = threadIdx.x;
    const int colId = blockIdx.x;  // y*nT + t
    const double* col = V + (size_t)colId * Nrep;
    if (tid == 0) nbadS = 0;
    __syncthreads();
    int nb = 0;
    for (long long r = tid; r < Nrep; r += BAM_BLOCK) {
        const double v = col[r];
        if (v == unfeas || v != v) ++nb;
    }
    if (nb) atomicAdd(&nbadS, nb);
    __syncthreads();
    const int nbad = nbadS;
    const bool drop = (double)nbad / (double)Nrep < 0.75;  // maxMVrate (:1306,1395)
    const int y = colId / nT, t = colId % nT;
    double* out = env + (size_t)y * BAMGPU_NENV * nT + t;
    if (!drop && nbad > 0) {  // too many bad values: flag the row (:1402-1404)
        if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
        return;
    }
    // load (bad values, if any, become +inf and sort to the end)
    for (int r = tid; r < npad; r += BAM_BLOCK) {
        double v = (r < Nrep) ? col[r] : INFINITY;
        if (v == unfeas || v != v) v = INFINITY;
        s[r] = v;
    }
    const long long m = Nrep - nbad;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += BAM_BLOCK) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = s[i], b = s[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (m <= 0) {
        if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
        return;
    }
    double part = 0.0;
    for (long long r = tid; r < m; r += BAM_BLOCK) part += s[r];
    const double mean = block_sum(part, scratch) / (double)m;
    part = 0.0;
    for (long long r = tid; r < m; r += BAM_BLOCK) { const double d = s[r] - mean; part += d * d; }
    const double ss = block_sum(part, scratch);
    if (tid == 0) {
        const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
        for (int q = 0; q < 5; ++q) out[(size_t)q * nT] = hazen(s, m, pr[q]);
        out[(size_t)5 * nT] = mean;
        out[(size_t)6 * nT] = (m > 1) ? sqrt(ss / (double)(m - 1)) : unfeas;
    }
}

// ---- large Nrep: exact selection kernels ------------------------------------------------------------------------
#define SEL_NT 10

// ---- large Nrep, last resort: MSB-first radix selection of all 10 order statistics at once --------------------------
//
// For the columns the histogram kernels give up on (heavy ties next to other values: a growth index that is exactly
// zero for most replications, a discharge that is zero below the activation stage for a part of the sample).  The
// order-preserving 64-bit key of a double is resolved 11 bits at a time (11,11,11,11,11,9): per pass one histogram per
// DISTINCT prefix among the 10 targets, then every target walks into its digit.  Exact whatever the data look like,
// 6 reads of the column + 1 for the sum of squares (the multi-level linear histograms it replaces needed >= 2 passes
// per target and level: 22+ reads on such columns).
#define RX_THREADS 1024
#define RX_NB 2048
#define RX_UNROLL 8
__device__ __forceinline__ unsigned long long rx_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);  // -0.0 -> +0.0
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double rx_val(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// colMap != nullptr: DEFERRED columns.  Block b works on column b of the side buffer V and writes the envelope of
// (output colMap[2b], input colMap[2b+1]) into an array of nT inputs per statistic; blocks at or beyond *colCount leave.
__global__ void __launch_bounds__(RX_THREADS) envelope_radix(long long Nrep, int nT, double unfeas,
                                                             const double* __restrict__ V, double* __restrict__ env,
                                                             const int* __restrict__ only, const int* __restrict__ colMap = nullptr,
                                                             const int* __restrict__ colCount = nullptr) {
    if (only != nullptr && only[blockIdx.x] == 0) return;
    if (colMap != nullptr && (int)blockIdx.x >= *colCount) return;
    extern __shared__ unsigned rxh[];  // [SEL_NT][RX_NB]
    __shared__ double scratch[32];
    __shared__ unsigned long long sPrefix[SEL_NT], sGPrefix[SEL_NT];
    __shared__ long long sRank[SEL_NT];
    __shared__ int sGroupOf[SEL_NT], sNG;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int colId = blockIdx.x;
    const double* __restrict__ col = V + (size_t)colId * Nrep;
    const int y = colMap ? colMap[2 * colId] : colId / nT, t = colMap ? colMap[2 * colId + 1] : colId % nT;
    double* out = env + (size_t)y * BAMGPU_NENV * nT + t;
    const int width[6] = {11, 11, 11, 11, 11, 9};
    const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};

    if (tid == 0) { sNG = 1; sGPrefix[0] = 0ull; }
    if (tid < SEL_NT) { sPrefix[tid] = 0ull; sGroupOf[tid] = 0; sRank[tid] = 0; }
    long long m = 0;
    double mean = 0.0;
    int shift = 64;
    for (int pass = 0; pass < 6; ++pass) {
        const int w = width[pass], nb = 1 << w;
        shift -= w;
        __syncthreads();
        const int nG = sNG;
        for (int b = tid; b < nG * RX_NB; b += RX_THREADS) rxh[b] = 0u;
        __syncthreads();
        double sum = 0.0;
        long long nbadLoc = 0;
        for (long long base = 0; base < Nrep; base += RX_UNROLL * RX_THREADS) {  // block-uniform trip count (warp collectives inside)
            const long long r0 = base + tid;
            double v[RX_UNROLL];  // RX_UNROLL independent loads in flight per thread: one block walks a column alone
#pragma unroll
            for (int q = 0; q < RX_UNROLL; ++q) {
                const long long r = r0 + (long long)q * RX_THREADS;
                v[q] = (r < Nrep) ? __ldcs(col + r) : unfeas;
            }
#pragma unroll
            for (int q = 0; q < RX_UNROLL; ++q) {
                const long long r = r0 + (long long)q * RX_THREADS;
                const bool inside = r < Nrep;
                const bool good = inside && !(v[q] == unfeas || v[q] != v[q]);
                if (pass == 0 && inside && !good) ++nbadLoc;
                const unsigned long long key = rx_key(v[q]);
                const unsigned digit = (unsigned)(key >> shift) & (unsigned)(nb - 1);
                if (pass == 0) {
                    if (good) sum += v[q];
                    // tie-heavy columns: a whole warp usually hits one bin
                    const unsigned act = __ballot_sync(0xffffffffu, good);
                    if (act) {
                        const int lead = __ffs(act) - 1;
                        const unsigned d0 = __shfl_sync(0xffffffffu, digit, lead);
                        const bool same = __all_sync(0xffffffffu, !good || digit == d0);
                        if (same) { if (lane == lead) atomicAdd(&rxh[d0], (unsigned)__popc(act)); }
                        else if (good) atomicAdd(&rxh[digit], 1u);
                    }
                } else if (good) {
                    const unsigned long long hi = key >> (shift + w);
                    for (int g = 0; g < nG; ++g)
                        if (hi == sGPrefix[g]) atomicAdd(&rxh[g * RX_NB + digit], 1u);
                }
            }
        }
        if (pass == 0) {
            const double nbd = block_sum((double)nbadLoc, scratch);
            const long long nbad = (long long)nbd;
            const bool drop = (double)nbad / (double)Nrep < 0.75;
            if ((!drop && nbad > 0) || nbad == Nrep) {
                if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
                return;
            }
            m = Nrep - nbad;
            mean = block_sum(sum, scratch) / (double)m;
            if (tid == 0)
                for (int q = 0; q < 5; ++q) {  // the 10 target ranks (0-based): i-1 and i of each Hazen quantile
                    const double hh = pr[q] * (double)m + 0.5;
                    long long i0, i1;
                    if (hh <= 1.0) i0 = i1 = 0;
                    else if (hh >= (double)m) i0 = i1 = m - 1;
                    else { const long long i = (long long)floor(hh); i0 = i - 1; i1 = i; }
                    sRank[2 * q] = i0; sRank[2 * q + 1] = i1;
                }
        }
        __syncthreads();
        // every target walks into its digit: warp k serves target k
        if (warp < SEL_NT) {
            const unsigned* hg = rxh + sGroupOf[warp] * RX_NB;
            const int per = nb / 32;
            long long rank = sRank[warp];
            long long mine = 0;
            for (int b = 0; b < per; ++b) mine += hg[lane * per + b];
            long long incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += o;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
            const int L = hit ? (__ffs(hit) - 1) : 31;
            int digit = 0;
            long long newRank = 0;
            if (lane == L) {
                long long acc = incl - mine;
                int b = 0;
                for (; b < per - 1; ++b) {
                    const long long c = hg[lane * per + b];
                    if (acc + c > rank) break;
                    acc += c;
                }
                digit = lane * per + b;
                newRank = rank - acc;
            }
            digit = __shfl_sync(0xffffffffu, digit, L);
            newRank = __shfl_sync(0xffffffffu, newRank, L);
            if (lane == 0) {
                sPrefix[warp] = (sPrefix[warp] << w) | (unsigned long long)digit;
                sRank[warp] = newRank;
            }
        }
        __syncthreads();
        if (tid == 0) {  // distinct prefixes -> groups of the next pass
            int ng = 0;
            for (int k = 0; k < SEL_NT; ++k) {
                int g = -1;
                for (int q = 0; q < ng && g < 0; ++q)
                    if (sGPrefix[q] == sPrefix[k]) g = q;
                if (g < 0) { g = ng; sGPrefix[ng++] = sPrefix[k]; }
                sGroupOf[k] = g;
            }
            sNG = ng;
        }
    }
    __syncthreads();
    double part = 0.0;
    for (long long r = tid; r < Nrep; r += RX_THREADS) {
        const double v = __ldcs(col + r);
        if (v == unfeas || v != v) continue;
        const double d = v - mean;
        part += d * d;
    }
    const double ss = block_sum(part, scratch);
    if (tid == 0) {
        for (int q = 0; q < 5; ++q) {
            const double a = rx_val(sPrefix[2 * q]), b = rx_val(sPrefix[2 * q + 1]);
            const double hh = pr[q] * (double)m + 0.5;
            double qv;
            if (hh <= 1.0 || hh >= (double)m) qv = a;
            else {
                const long long i = (long long)floor(hh);
                const double frac = hh - (double)i;
                qv = a + frac * (b - a);
            }
            out[(size_t)q * nT] = qv;
        }
        out[(size_t)5 * nT] = mean;
        out[(size_t)6 * nT] = (m > 1) ? sqrt(ss / (double)(m - 1)) : unfeas;
    }
}

// Columns the histogram kernels gave up on are few (1 % of Vegetation's 5475) but the radix walk reads each of them 7 times
// with ONE block, at the ~30 GB/s a single SM pulls: inside a tile the few blocks of that launch held up everything behind
// them (190 ms of a 920 ms prediction for 60 columns).  They are copied to a side buffer instead (slot handed out by an
// atomic counter; the flag is cleared so the in-tile launch skips the column) and walked together after the last tile,
// one block per column, all at once.  A column that finds the buffer full keeps its flag and is walked in its tile.
// Two kernels: slots first (one thread per column), then the copy with several blocks per column.
__global__ void __launch_bounds__(256) stash_assign(int ncol, int nT, int tile0, int* __restrict__ needSlow, int sideCap,
                                                   int* __restrict__ count, int* __restrict__ colMap, int* __restrict__ slotOf) {
    const int colId = blockIdx.x * 256 + threadIdx.x;
    if (colId >= ncol) return;
    int slot = -1;
    if (needSlow[colId] != 0) {
        slot = atomicAdd(count, 1);
        if (slot < sideCap) {
            colMap[2 * slot] = colId / nT;
            colMap[2 * slot + 1] = tile0 + colId % nT;
            needSlow[colId] = 0;
        } else {
            atomicSub(count, 1);
            slot = -1;
        }
    }
    slotOf[colId] = slot;
}
// the copy: STASH_SPLIT blocks per column (a single block moves ~10 GB/s; 45 columns of 80 MB took 8 ms per tile)
#define STASH_SPLIT 16
__global__ void __launch_bounds__(256) stash_copy(long long Nrep, const double* __restrict__ V, const int* __restrict__ slotOf,
                                                 double* __restrict__ side) {
    const int colId = blockIdx.x, slot = slotOf[colId];
    if (slot < 0) return;
    const double* __restrict__ src = V + (size_t)colId * Nrep;
    double* __restrict__ dst = side + (size_t)slot * Nrep;
    const long long per = (Nrep + STASH_SPLIT - 1) / STASH_SPLIT;
    const long long r0 = (long long)blockIdx.y * per, r1 = min(Nrep, r0 + per);
    for (long long r = r0 + threadIdx.x; r < r1; r += 4 * 256) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (r + q * 256 < r1) ? __ldcs(src + r + q * 256) : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (r + q * 256 < r1) dst[r + q * 256] = v[q];
    }
}

// ---- sample-sharded envelopes (few prediction inputs, many samples) ---------------------------------
//
// Every rank holds Nloc of the Nrep replications of EVERY column.  The order statistics of a column are then selected by
// the same MSB-first radix walk as envelope_radix, cut at the histogram: per pass a rank counts the digits of ITS values
// (one histogram per distinct prefix among the 10 targets), the histograms are summed over the ranks (one all-reduce of
// unsigned counts per pass), and every rank walks the summed histograms -- identically, so the selection state never has
// to be exchanged.  After 6 passes the 64-bit keys of the 10 order statistics are known: EXACT quantiles whatever the
// data, 7 reads of the local shard.  Count of bad values, sum and sum of squared deviations travel the same way (one
// all-reduce of doubles each; the sums are fixed trees per rank, the cross-rank order is NCCL's).
struct RxsState {
    unsigned long long prefix[SEL_NT], gprefix[SEL_NT];
    long long rank[SEL_NT];
    int groupOf[SEL_NT];
    int nG, flagged;
    long long m;
    double mean;
};

__global__ void rxs_init(RxsState* __restrict__ st, int ncol) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    RxsState z;
    for (int k = 0; k < SEL_NT; ++k) { z.prefix[k] = 0ull; z.gprefix[k] = 0ull; z.rank[k] = 0; z.groupOf[k] = 0; }
    z.nG = 1; z.flagged = 0; z.m = 0; z.mean = 0.0;
    st[c] = z;
}

__global__ void __launch_bounds__(RX_THREADS) rxs_hist(long long Nloc, double unfeas, const double* __restrict__ V,
                                                       const RxsState* __restrict__ st, int pass, unsigned* __restrict__ hist,
                                                       double* __restrict__ acc) {
    extern __shared__ unsigned rxh[];  // [SEL_NT][RX_NB]
    __shared__ double scratch[32];
    __shared__ unsigned long long sG[SEL_NT];
    const int tid = threadIdx.x, lane = tid & 31;
    const int colId = blockIdx.x;
    const double* __restrict__ col = V + (size_t)colId * Nloc;
    unsigned* __restrict__ hout = hist + (size_t)colId * SEL_NT * RX_NB;
    const int nG = st[colId].nG;
    const bool flagged = st[colId].flagged != 0;
    if (tid < SEL_NT) sG[tid] = st[colId].gprefix[tid];
    for (int b = tid; b < SEL_NT * RX_NB; b += RX_THREADS) rxh[b] = 0u;
    __syncthreads();
    const int width[6] = {11, 11, 11, 11, 11, 9};
    int shift = 64;
    for (int q = 0; q <= pass; ++q) shift -= width[q];
    const int w = width[pass], nb = 1 << w;
    double sum = 0.0;
    long long nbadLoc = 0;
    if (!flagged)
        for (long long base = 0; base < Nloc; base += 4 * RX_THREADS) {
            const long long r0 = base + tid;
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long r = r0 + (long long)q * RX_THREADS;
                v[q] = (r < Nloc) ? __ldcs(col + r) : unfeas;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long r = r0 + (long long)q * RX_THREADS;
                const bool inside = r < Nloc;
                const bool good = inside && !(v[q] == unfeas || v[q] != v[q]);
                if (pass == 0 && inside && !good) ++nbadLoc;
                const unsigned long long key = rx_key(v[q]);
                const unsigned digit = (unsigned)(key >> shift) & (unsigned)(nb - 1);
                if (pass == 0) {
                    if (good) sum += v[q];
                    const unsigned act = __ballot_sync(0xffffffffu, good);
                    if (act) {
                        const int lead = __ffs(act) - 1;
                        const unsigned d0 = __shfl_sync(0xffffffffu, digit, lead);
                        const bool same = __all_sync(0xffffffffu, !good || digit == d0);
                        if (same) { if (lane == lead) atomicAdd(&rxh[d0], (unsigned)__popc(act)); }
                        else if (good) atomicAdd(&rxh[digit], 1u);
                    }
                } else if (good) {
                    const unsigned long long hi = key >> (shift + w);
                    for (int g = 0; g < nG; ++g)
                        if (hi == sG[g]) atomicAdd(&rxh[g * RX_NB + digit], 1u);
                }
            }
        }
    __syncthreads();
    for (int b = tid; b < SEL_NT * RX_NB; b += RX_THREADS) hout[b] = rxh[b];
    if (pass == 0) {
        const double nbd = block_sum((double)nbadLoc, scratch);
        const double sm = block_sum(sum, scratch);
        if (tid == 0) { acc[2 * colId] = nbd; acc[2 * colId + 1] = sm; }
    }
}

// every target walks into its digit of the SUMMED histograms: warp k serves target k (identical on every rank)
__global__ void __launch_bounds__(32 * SEL_NT) rxs_walk(long long Nrep, RxsState* __restrict__ stAll, const unsigned* __restrict__ hist,
                                                        const double* __restrict__ acc, int pass) {
    __shared__ RxsState st;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int colId = blockIdx.x;
    if (tid == 0) st = stAll[colId];
    __syncthreads();
    const int width[6] = {11, 11, 11, 11, 11, 9};
    const int w = width[pass], nb = 1 << w;
    if (pass == 0 && tid == 0) {
        const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
        const long long nbad = (long long)acc[2 * colId];
        const bool drop = (double)nbad / (double)Nrep < 0.75;  // maxMVrate (:1306,1395)
        if ((!drop && nbad > 0) || nbad == Nrep) st.flagged = 1;
        else {
            st.m = Nrep - nbad;
            st.mean = acc[2 * colId + 1] / (double)st.m;
            for (int q = 0; q < 5; ++q) {  // the 10 target ranks (0-based): i-1 and i of each Hazen quantile
                const double hh = pr[q] * (double)st.m + 0.5;
                long long i0, i1;
                if (hh <= 1.0) i0 = i1 = 0;
                else if (hh >= (double)st.m) i0 = i1 = st.m - 1;
                else { const long long i = (long long)floor(hh); i0 = i - 1; i1 = i; }
                st.rank[2 * q] = i0; st.rank[2 * q + 1] = i1;
            }
        }
    }
    __syncthreads();
    if (!st.flagged) {
        const unsigned* hg = hist + ((size_t)colId * SEL_NT + st.groupOf[warp]) * RX_NB;
        const int per = nb / 32;
        const long long rank = st.rank[warp];
        long long mine = 0;
        for (int b = 0; b < per; ++b) mine += hg[lane * per + b];
        long long incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, incl > rank);
        const int L = hit ? (__ffs(hit) - 1) : 31;
        int digit = 0;
        long long newRank = 0;
        if (lane == L) {
            long long a = incl - mine;
            int b = 0;
            for (; b < per - 1; ++b) {
                const long long c = hg[lane * per + b];
                if (a + c > rank) break;
                a += c;
            }
            digit = lane * per + b;
            newRank = rank - a;
        }
        digit = __shfl_sync(0xffffffffu, digit, L);
        newRank = __shfl_sync(0xffffffffu, newRank, L);
        __syncthreads();
        if (lane == 0) {
            st.prefix[warp] = (st.prefix[warp] << w) | (unsigned long long)digit;
            st.rank[warp] = newRank;
        }
        __syncthreads();
        if (tid == 0) {  // distinct prefixes -> groups of the next pass
            int ng = 0;
            for (int k = 0; k < SEL_NT; ++k) {
                int g = -1;
                for (int q = 0; q < ng && g < 0; ++q)
                    if (st.gprefix[q] == st.prefix[k]) g = q;
                if (g < 0) { g = ng; st.gprefix[ng++] = st.prefix[k]; }
                st.groupOf[k] = g;
            }
            st.nG = ng;
        }
    }
    __syncthreads();
    if (tid == 0) stAll[colId] = st;
}

__global__ void __launch_bounds__(RX_THREADS) rxs_ss(long long Nloc, double unfeas, const double* __restrict__ V,
                                                     const RxsState* __restrict__ st, double* __restrict__ ss) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x, colId = blockIdx.x;
    const double* __restrict__ col = V + (size_t)colId * Nloc;
    const double mean = st[colId].mean;
    double part = 0.0;
    if (!st[colId].flagged)
        for (long long r = tid; r < Nloc; r += RX_THREADS) {
            const double v = __ldcs(col + r);
            if (v == unfeas || v != v) continue;
            const double d = v - mean;
            part += d * d;
        }
    const double t = block_sum(part, scratch);
    if (tid == 0) ss[colId] = t;
}

__global__ void rxs_finish(int ncol, int nT, double unfeas, const RxsState* __restrict__ stAll, const double* __restrict__ ss,
                           double* __restrict__ env) {
    const int colId = blockIdx.x * blockDim.x + threadIdx.x;
    if (colId >= ncol) return;
    const RxsState& st = stAll[colId];
    const int y = colId / nT, t = colId % nT;
    double* out = env + (size_t)y * BAMGPU_NENV * nT + t;
    if (st.flagged) {
        for (int q = 0; q < BAMGPU_NENV; ++q) out[(size_t)q * nT] = unfeas;
        return;
    }
    const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
    for (int q = 0; q < 5; ++q) {
        const double a = rx_val(st.prefix[2 * q]), b = rx_val(st.prefix[2 * q + 1]);
        const double hh = pr[q] * (double)st.m + 0.5;
        double qv;
        if (hh <= 1.0 || hh >= (double)st.m) qv = a;
        else {
            const long long i = (long long)floor(hh);
            qv = a + (hh - (double)i) * (b - a);
        }
        out[(size_t)q * nT] = qv;
    }
    out[(size_t)5 * nT] = st.mean;
    out[(size_t)6 * nT] = (st.m > 1) ? sqrt(ss[colId] / (double)(st.m - 1)) : unfeas;
}

// ---- large Nrep, fast path: all 10 order statistics with 4 reads of the column ---------------------------------
//
// pass A  min / max / #bad / sum                                   (1 read)
// pass B  2048-bin histogram on [min,max]  +  sum (v-mean)^2        (1 read)   -> bin and residual rank per target
// pass C  1024-bin sub-histogram inside every selected bin (<= 10)  (1 read)   -> sub-bin and residual rank
// pass D  gather the members of the selected sub-bins (<= 256 each), sort, pick (1 read)
// A target whose sub-bin still holds more than 256 values (heavy ties / pathological clustering) is resolved by
// envelope_radix on its own (rare).  bin functions are monotone in v => exact.
#define S2_NB1 2048
#define S2_NB2 1024
#define S2_CAP 256
#define S2_THREADS 512

__device__ __forceinline__ int s2_bin(double v, double lo, double scale, int nb) {
    int b = (int)((v - lo) * scale);
    return b < 0 ? 0 : (b >= nb ? nb - 1 : b);
}

__device__ double s2_block_sum(double v, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < S2_THREADS / 32; ++i) s += scratch[i];
        scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

__global__ void __launch_bounds__(S2_THREADS) envelope_select2(long long Nrep, int nT, double unfeas,
                                                               const double* __restrict__ V, double* __restrict__ env,
                                                               int* __restrict__ needSlow, const int* __restrict__ only) {
    if (only != nullptr && only[blockIdx.x] == 0) {  // already resolved by envelope_select3
        if (threadIdx.x == 0) needSlow[blockIdx.x] = 0;
        return;
    }
    if (only != nullptr && only[blockIdx.x] == 2) {  // tie-heavy column: envelope_select3 sends it straight to the radix walk
        if (threadIdx.x == 0) needSlow[blockIdx.x] = 1;
        return;
    }
    extern __shared__ __align__(16) unsigned char s2raw[];
    unsigned* hist = reinterpret_cast<unsigned*>(s2raw);                  // max(NB1, 10*NB2) counters = 40 KB
    double* gath = reinterpret_cast<double*>(s2raw + sizeof(unsigned) * SEL_NT * S2_NB2);  // 10 x CAP doubles = 20 KB
    __shared__ double scratch[40];
    __shared__ short slotOfBin[S2_NB1];
    __shared__ long long tRank[SEL_NT];
    __shared__ int tBin1[SEL_NT], tSlot[SEL_NT], tBin2[SEL_NT], tList[SEL_NT], nLists, listSlot[SEL_NT],
        listBin2[SEL_NT];
    __shared__ double slotLo[SEL_NT], slotScale[SEL_NT], tVal[SEL_NT];
    __shared__ unsigned prefix[S2_THREADS];
    const int tid = threadIdx.x;
    const int colId = blockIdx.x;
    const double* col = V + (size_t)colId * Nrep;
    const int y = colId / nT, t = colId % nT;
    double* out = env + (size_t)y * BAMGPU_NENV * nT + t;
    if (tid == 0) needSlow[colId] = 0;

    // ---- pass A
    double mn = INFINITY, mx = -INFINITY, sum = 0.0, nb = 0.0;
    for (long long r = tid; r < Nrep; r += S2_THREADS) {
        const double v = col[r];
        if (v == unfeas || v != v) { nb += 1.0; continue; }
        mn = fmin(mn, v); mx = fmax(mx, v); sum += v;
    }
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((tid & 31) == 0) { scratch[tid >> 5] = mn; }
    __syncthreads();
    if (tid == 0) { double v = INFINITY; for (int i = 0; i < S2_THREADS / 32; ++i) v = fmin(v, scratch[i]); scratch[33] = v; }
    __syncthreads();
    mn = scratch[33];
    __syncthreads();
    if ((tid & 31) == 0) { scratch[tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) { double v = -INFINITY; for (int i = 0; i < S2_THREADS / 32; ++i) v = fmax(v, scratch[i]); scratch[34] = v; }
    __syncthreads();
    mx = scratch[34];
    const long long nbad = (long long)s2_block_sum(nb, scratch);
    const bool drop = (double)nbad / (double)Nrep < 0.75;
    if ((!drop && nbad > 0) || nbad == Nrep) {
        if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
        return;
    }
    const long long m = Nrep - nbad;
    const double mean = s2_block_sum(sum, scratch) / (double)m;

    // ---- pass B: level-1 histogram + sum of squares
    const double scale1 = (mx > mn) ? (double)S2_NB1 / (mx - mn) : 0.0;
    for (int b = tid; b < S2_NB1; b += S2_THREADS) { hist[b] = 0; slotOfBin[b] = -1; }
    __syncthreads();
    double part = 0.0;
    for (long long r = tid; r < Nrep; r += S2_THREADS) {
        const double v = col[r];
        if (v == unfeas || v != v) continue;
        const double d = v - mean;
        part += d * d;
        atomicAdd(&hist[s2_bin(v, mn, scale1, S2_NB1)], 1u);
    }
    const double ss = s2_block_sum(part, scratch);
    // exclusive prefix over the 2048 bins (4 bins per thread + scan of the 512 partials)
    {
        unsigned loc[S2_NB1 / S2_THREADS], acc = 0;
#pragma unroll
        for (int q = 0; q < S2_NB1 / S2_THREADS; ++q) { loc[q] = hist[tid * (S2_NB1 / S2_THREADS) + q]; acc += loc[q]; }
        prefix[tid] = acc;
        __syncthreads();
        for (int off = 1; off < S2_THREADS; off <<= 1) {
            unsigned add = (tid >= off) ? prefix[tid - off] : 0;
            __syncthreads();
            prefix[tid] += add;
            __syncthreads();
        }
        unsigned base = prefix[tid] - acc;
#pragma unroll
        for (int q = 0; q < S2_NB1 / S2_THREADS; ++q) { const unsigned c0 = loc[q]; hist[tid * (S2_NB1 / S2_THREADS) + q] = base; base += c0; }
        __syncthreads();  // hist[b] = number of values in bins < b
    }
    const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
    if (tid < SEL_NT) {
        const int q = tid >> 1;
        const double hh = pr[q] * (double)m + 0.5;
        long long rk;
        if (hh <= 1.0) rk = 0;
        else if (hh >= (double)m) rk = m - 1;
        else rk = (long long)floor(hh) - 1 + (tid & 1);
        // largest bin b with hist[b] <= rk  (binary search on the exclusive prefix)
        int lo = 0, hi = S2_NB1 - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((long long)hist[mid] <= rk) lo = mid; else hi = mid - 1;
        }
        tBin1[tid] = lo;
        tRank[tid] = rk - hist[lo];
    }
    __syncthreads();
    if (tid == 0) {  // distinct level-1 bins -> slots
        int ns = 0;
        for (int k = 0; k < SEL_NT; ++k) {
            int sl = -1;
            for (int j = 0; j < ns; ++j)
                if (slotOfBin[tBin1[k]] == j) sl = j;
            if (sl < 0) {
                sl = ns++;
                slotOfBin[tBin1[k]] = (short)sl;
                // value range of the bin (approximate edges are fine: clamping keeps the sub-bin function monotone)
                const double w = (mx - mn) / (double)S2_NB1;
                slotLo[sl] = mn + w * (double)tBin1[k];
                slotScale[sl] = (w > 0.0) ? (double)S2_NB2 / w : 0.0;
            }
            tSlot[k] = sl;
        }
    }
    __syncthreads();
    const bool degenerate = !(mx > mn);
    if (!degenerate) {
        // ---- pass C: level-2 histograms of the selected bins
        for (int b = tid; b < SEL_NT * S2_NB2; b += S2_THREADS) hist[b] = 0;
        __syncthreads();
        for (long long r = tid; r < Nrep; r += S2_THREADS) {
            const double v = col[r];
            if (v == unfeas || v != v) continue;
            const int sl = slotOfBin[s2_bin(v, mn, scale1, S2_NB1)];
            if (sl >= 0) atomicAdd(&hist[sl * S2_NB2 + s2_bin(v, slotLo[sl], slotScale[sl], S2_NB2)], 1u);
        }
        __syncthreads();
        if (tid < SEL_NT) {  // serial scan of one 1024-bin histogram per target
            const unsigned* hsl = hist + tSlot[tid] * S2_NB2;
            long long acc = 0;
            int b = 0;
            for (; b < S2_NB2; ++b) {
                if (acc + hsl[b] > tRank[tid]) break;
                acc += hsl[b];
            }
            if (b >= S2_NB2) b = S2_NB2 - 1;
            tBin2[tid] = b;
            tRank[tid] -= acc;
        }
        __syncthreads();
        if (tid == 0) {  // distinct (slot, sub-bin) -> gather lists
            int nl = 0;
            for (int k = 0; k < SEL_NT; ++k) {
                int li = -1;
                for (int j = 0; j < nl; ++j)
                    if (listSlot[j] == tSlot[k] && listBin2[j] == tBin2[k]) li = j;
                if (li < 0) { li = nl++; listSlot[li] = tSlot[k]; listBin2[li] = tBin2[k]; }
                tList[k] = li;
            }
            nLists = nl;
        }
        __syncthreads();
        __shared__ int lCount[SEL_NT];
        if (tid < SEL_NT) lCount[tid] = 0;
        __syncthreads();
        // ---- pass D: gather
        for (long long r = tid; r < Nrep; r += S2_THREADS) {
            const double v = col[r];
            if (v == unfeas || v != v) continue;
            const int sl = slotOfBin[s2_bin(v, mn, scale1, S2_NB1)];
            if (sl < 0) continue;
            const int b2 = s2_bin(v, slotLo[sl], slotScale[sl], S2_NB2);
            for (int j = 0; j < nLists; ++j)
                if (listSlot[j] == sl && listBin2[j] == b2) {
                    const int pos = atomicAdd(&lCount[j], 1);
                    if (pos < S2_CAP) gath[j * S2_CAP + pos] = v;
                }
        }
        __syncthreads();
        // sort every list (bitonic, padded with +inf) -- one list at a time, whole block
        for (int j = 0; j < nLists; ++j) {
            const int cnt = min(lCount[j], S2_CAP);
            double* g = gath + j * S2_CAP;
            for (int i = cnt + tid; i < S2_CAP; i += S2_THREADS) g[i] = INFINITY;
            __syncthreads();
            for (int kk = 2; kk <= S2_CAP; kk <<= 1)
                for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                    for (int i = tid; i < S2_CAP; i += S2_THREADS) {
                        const int ixj = i ^ jj;
                        if (ixj > i) {
                            const double a = g[i], b = g[ixj];
                            const bool up = ((i & kk) == 0);
                            if ((a > b) == up) { g[i] = b; g[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
        }
        if (tid < SEL_NT) {
            const int j = tList[tid];
            if (lCount[j] <= S2_CAP) tVal[tid] = gath[j * S2_CAP + (int)tRank[tid]];
            else tVal[tid] = NAN;  // unresolved: too many members (ties): slow path
        }
    } else if (tid < SEL_NT) tVal[tid] = mn;
    __syncthreads();
    if (tid == 0) {
        bool slow = false;
        for (int k = 0; k < SEL_NT; ++k) slow |= (tVal[k] != tVal[k]);
        if (slow) { needSlow[colId] = 1; return; }
        needSlow[colId] = 0;
        for (int q = 0; q < 5; ++q) {
            const double hh = pr[q] * (double)m + 0.5;
            double qv;
            if (hh <= 1.0 || hh >= (double)m) qv = tVal[2 * q];
            else {
                const long long i = (long long)floor(hh);
                const double frac = hh - (double)i;
                qv = tVal[2 * q] + frac * (tVal[2 * q + 1] - tVal[2 * q]);
            }
            out[(size_t)q * nT] = qv;
        }
        out[(size_t)5 * nT] = mean;
        out[(size_t)6 * nT] = (m > 1) ? sqrt(ss / (double)(m - 1)) : unfeas;
    }
}

// ---- large Nrep, 2-read path: all 10 order statistics with TWO reads of the column ---------------------------------
//
// phase 0  strided subsample of 2048 values, sorted in shared memory: bin range [lo,hi] = its 0.5 % / 99.5 % points
//          (the targets are the 2.5 % .. 97.5 % quantiles and their neighbours), pivot = its median.  Values outside
//          [lo,hi] are CLAMPED into the two edge bins: the bin function stays monotone, so the selection stays exact
//          whatever the subsample looks like; a bad subsample only costs a fallback.
// pass B   histogram (32 bins per thread) in shared memory + #bad + sum (v-pivot) + sum (v-pivot)^2     (1 read)
//          -> m, mean, variance (shifted-data formula), bin and residual rank of every target
// pass C   gather the members of the <= 10 selected bins (<= S3_CAP each), sort, pick       (1 read)
// A selected bin with more than S3_CAP members is still resolved when all its members are equal (ties: e.g. Q = 0
// below the first activation stage); anything else is flagged for the 4-read kernel envelope_select2.
// (bins of envelope_select3: 32 per thread of its block, see the kernel)
#define S3_SAMPLE 2048
#define S3_ATOMS 2   /* atoms per column handled by envelope_select3 */
#define S3_RUN 8     /* run length in the sorted subsample that makes a value an atom */

template <int NB>
__device__ __forceinline__ int s3_bin(double v, double lo, double scale) {
    if (v < lo) return 0;
    const double f = (v - lo) * scale;
    return (f >= (double)(NB - 2)) ? NB - 1 : 1 + (int)f;
}

__device__ __forceinline__ unsigned long long s3_key(double v) {  // order-preserving map double -> uint64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}

template <int NTHR>
__device__ double s3_block_sum(double v, double* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < NTHR / 32; ++i) s += scratch[i];
        scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

// Column walkers of envelope_select3: 128-bit streaming loads, four independent requests per thread in flight
// (the passes are pure HBM streams; with one 8-byte load per thread and iteration they were latency-bound at
// 2.3 TB/s).  `f(v)` is applied to every value exactly once; the order does not matter to the callers.
template <int NTHR, class F>
__device__ __forceinline__ void s3_for_each(const double* __restrict__ col, long long Nrep, F f) {
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(col) & 15) == 0) {
        const double2* __restrict__ c2 = reinterpret_cast<const double2*>(col);
        const long long n2 = Nrep >> 1;
        long long i = tid;
        for (; i + 3 * NTHR < n2; i += 4 * NTHR) {
            const double2 a = __ldcs(c2 + i), b = __ldcs(c2 + i + NTHR), c = __ldcs(c2 + i + 2 * NTHR),
                          d = __ldcs(c2 + i + 3 * NTHR);
            f(a.x); f(a.y); f(b.x); f(b.y); f(c.x); f(c.y); f(d.x); f(d.y);
        }
        for (; i < n2; i += NTHR) { const double2 a = __ldcs(c2 + i); f(a.x); f(a.y); }
        if ((Nrep & 1) && tid == 0) f(col[Nrep - 1]);
    } else {
        for (long long r = tid; r < Nrep; r += NTHR) f(__ldcs(col + r));
    }
}

// CAP = capacity of a gather list (power of two: the lists are sorted in place by a bitonic network), NTHR = block size.
// <512, 512>: 64 KB of shared memory, 3 blocks per SM (Nrep <= 3e6); <2048, 1024>: 160 KB, 1 block per SM (Nrep <= 1.2e7).
template <int CAP, int NTHR>
__global__ void __launch_bounds__(NTHR) envelope_select3(long long Nrep, int nT, double unfeas,
                                                               const double* __restrict__ V, double* __restrict__ env,
                                                               int* __restrict__ needSlow) {
    // bins: 32 per thread of the block -- 16384 for <512, 512>, 32768 for <2048, 1024>, whose 10^7-value columns put more than
    // CAP values into the densest bins of a 16384-bin histogram (no ties needed: 8 % of Vegetation's columns went to the radix walk)
    constexpr int NB = 32 * NTHR;
    extern __shared__ __align__(16) unsigned char s3raw[];  // 64 KB: sample (16 KB) -> histogram (64 KB) -> gather lists (40 KB)
    double* smp = reinterpret_cast<double*>(s3raw);
    unsigned* hist = reinterpret_cast<unsigned*>(s3raw);
    double* gath = reinterpret_cast<double*>(s3raw);
    __shared__ double scratch[40];
    __shared__ unsigned bitmap[NB / 32];
    __shared__ unsigned prefix[NTHR];
    __shared__ long long tRank[SEL_NT];
    __shared__ int tBin[SEL_NT], tList[SEL_NT], listBin[SEL_NT], listCnt[SEL_NT], listOver[SEL_NT], listOff[SEL_NT], gCnt[SEL_NT], nLists, nGood;
    __shared__ unsigned long long listMin[SEL_NT], listMax[SEL_NT];
    __shared__ double tVal[SEL_NT];
    // ATOMS: values the subsample holds S3_RUN or more times (0.4 % of the column: far more than a gather list takes) -- a
    // growth index that is exactly 0 before the season, cos(pi g0) = -1 at its peak, Q = 0 below an activation stage.  They are
    // counted, not gathered: the bin that holds one keeps its place in the prefix sums, its list holds the OTHER values of
    // the bin, and a rank inside the bin is resolved against (sorted list + c copies of the atom).
    __shared__ double atomV[S3_ATOMS];
    __shared__ long long atomC[S3_ATOMS];
    __shared__ int nAtomRaw, listAtoms[SEL_NT];
    const int tid = threadIdx.x;
    const int colId = blockIdx.x;
    const double* col = V + (size_t)colId * Nrep;
    const int y = colId / nT, t = colId % nT;
    double* out = env + (size_t)y * BAMGPU_NENV * nT + t;

    // ---- phase 0: subsample
    if (tid == 0) nGood = 0;
    __syncthreads();
    const long long stride = Nrep / S3_SAMPLE;
    int good = 0;
    for (int i = tid; i < S3_SAMPLE; i += NTHR) {
        double v = col[(long long)i * stride];
        if (v == unfeas || v != v) v = INFINITY; else ++good;
        smp[i] = v;
    }
    if (good) atomicAdd(&nGood, good);
    __syncthreads();
    for (int kk = 2; kk <= S3_SAMPLE; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < S3_SAMPLE; i += NTHR) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = smp[i], b = smp[ixj];
                    const bool up = ((i & kk) == 0);
                    if ((a > b) == up) { smp[i] = b; smp[ixj] = a; }
                }
            }
            __syncthreads();
        }
    const int ns = nGood;
    if (ns < 256) {  // mostly bad values: leave it to the general kernels
        if (tid == 0) needSlow[colId] = 1;
        return;
    }
    double lo = smp[ns / 200], hi = smp[ns - 1 - ns / 200];
    const double pivot = smp[ns / 2], smin = smp[0], smax = smp[ns - 1];
    if (tid == 0) nAtomRaw = 0;
    __syncthreads();
    for (int i = tid; i + S3_RUN <= ns; i += NTHR) {  // start of a run of >= S3_RUN equal values
        const double v = smp[i];
        if (v == smp[i + S3_RUN - 1] && (i == 0 || smp[i - 1] != v)) {
            const int a = atomicAdd(&nAtomRaw, 1);
            if (a < S3_ATOMS) atomV[a] = v;
        }
    }
    __syncthreads();
    // more runs than slots: which ones got a slot depends on the atomics' order, so use none (the answers never depend on
    // the atoms, only the path does; this keeps the path itself repeatable)
    const int nAtom = (nAtomRaw <= S3_ATOMS) ? nAtomRaw : 0;
    double at0 = (nAtom > 0) ? atomV[0] : NAN, at1 = (nAtom > 1) ? atomV[1] : NAN;
    if (nAtom == 2 && at1 < at0) { const double t = at0; at0 = at1; at1 = t; }
    __syncthreads();
    if (!(hi > lo) || !(hi - lo < INFINITY)) {
        if (hi == lo && smin == smax) {
            // the whole subsample is ONE value v0.  One read counts the values below / above / equal and their moments: when
            // every target rank falls inside the block of v0 (always, for a constant column: a growth index outside the
            // season for every replication, cos(pi g0) = 1 there, Q = 0 on a dry bed) no selection is needed at all --
            // instead of the 7 reads of the radix walk (Vegetation 10^7 x 1095 x 5: a third of the columns)
            const double v0 = lo;
            double nb = 0.0, nl = 0.0, ng = 0.0, s1 = 0.0, s2 = 0.0;
            s3_for_each<NTHR>(col, Nrep, [&](double v) {
                if (v == unfeas || v != v) { nb += 1.0; return; }
                if (v == v0) return;
                if (v < v0) nl += 1.0; else ng += 1.0;
                const double d = v - v0;
                s1 += d;
                s2 = fma(d, d, s2);
            });
            const long long nbad = (long long)s3_block_sum<NTHR>(nb, scratch);
            const long long nless = (long long)s3_block_sum<NTHR>(nl, scratch), ngreat = (long long)s3_block_sum<NTHR>(ng, scratch);
            const double S1 = s3_block_sum<NTHR>(s1, scratch), S2 = s3_block_sum<NTHR>(s2, scratch);
            const bool drop = (double)nbad / (double)Nrep < 0.75;  // maxMVrate (:1306,1395)
            const long long m = Nrep - nbad;
            if ((!drop && nbad > 0) || nbad == Nrep) {
                if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
                if (tid == 0) needSlow[colId] = 0;
                return;
            }
            // lowest and highest target rank (Hazen positions of 2.5 % and 97.5 %, both neighbours)
            const double hLo = 0.025 * (double)m + 0.5, hHi = 0.975 * (double)m + 0.5;
            const long long rLo = (hLo <= 1.0) ? 0 : ((hLo >= (double)m) ? m - 1 : (long long)floor(hLo) - 1);
            const long long rHi = (hHi <= 1.0) ? 0 : ((hHi >= (double)m) ? m - 1 : (long long)floor(hHi));
            if (rLo >= nless && rHi < m - ngreat) {
                if (tid == 0) {
                    for (int k = 0; k < 5; ++k) out[(size_t)k * nT] = v0;
                    out[(size_t)5 * nT] = v0 + S1 / (double)m;
                    out[(size_t)6 * nT] = (m > 1) ? sqrt(fmax(S2 - S1 * (S1 / (double)m), 0.0) / (double)(m - 1)) : unfeas;
                    needSlow[colId] = 0;
                }
                return;
            }
        }
        if (smax > smin && smax - smin < INFINITY && nAtom > 0) {
            lo = smin; hi = smax;  // 99 % of the subsample is one value (an atom): bin the full range of the subsample instead
        } else {
            if (tid == 0) needSlow[colId] = (hi == lo && Nrep > 6000000) ? 2 : 1;  // ties next to other values
            return;
        }
    }
    const double scale = (double)(NB - 2) / (hi - lo);

    // ---- pass B
    for (int b = tid; b < NB; b += NTHR) hist[b] = 0;
    for (int b = tid; b < NB / 32; b += NTHR) bitmap[b] = 0;
    __syncthreads();
    double s1 = 0.0, s2 = 0.0, nb = 0.0, ca0 = 0.0, ca1 = 0.0;
    s3_for_each<NTHR>(col, Nrep, [&](double v) {
        if (v == unfeas || v != v) { nb += 1.0; return; }
        const double d = v - pivot;
        s1 += d;
        s2 = fma(d, d, s2);
        if (v == at0) { ca0 += 1.0; return; }
        if (v == at1) { ca1 += 1.0; return; }
        atomicAdd(&hist[s3_bin<NB>(v, lo, scale)], 1u);
    });
    if (nAtom > 0) {  // the atoms' bins count them all the same
        const long long c0 = (long long)s3_block_sum<NTHR>(ca0, scratch), c1 = (long long)s3_block_sum<NTHR>(ca1, scratch);
        if (tid == 0) {
            atomV[0] = at0; atomC[0] = c0;
            atomV[1] = at1; atomC[1] = (nAtom > 1) ? c1 : 0;
            hist[s3_bin<NB>(at0, lo, scale)] += (unsigned)c0;
            if (nAtom > 1) hist[s3_bin<NB>(at1, lo, scale)] += (unsigned)c1;
        }
        __syncthreads();
    }
    const long long nbad = (long long)s3_block_sum<NTHR>(nb, scratch);
    const bool drop = (double)nbad / (double)Nrep < 0.75;  // maxMVrate (:1306,1395)
    if ((!drop && nbad > 0) || nbad == Nrep) {
        if (tid < BAMGPU_NENV) out[(size_t)tid * nT] = unfeas;
        if (tid == 0) needSlow[colId] = 0;
        return;
    }
    const long long m = Nrep - nbad;
    const double S1 = s3_block_sum<NTHR>(s1, scratch), S2 = s3_block_sum<NTHR>(s2, scratch);
    const double mean = pivot + S1 / (double)m;
    const double ss = fmax(S2 - S1 * (S1 / (double)m), 0.0);
    {   // exclusive prefix over the bins: 32 consecutive bins per thread + scan of the 512 partials
        constexpr int PER = NB / NTHR;
        unsigned acc = 0;
        for (int q = 0; q < PER; ++q) acc += hist[tid * PER + q];
        prefix[tid] = acc;
        __syncthreads();
        for (int off = 1; off < NTHR; off <<= 1) {
            const unsigned add = (tid >= off) ? prefix[tid - off] : 0;
            __syncthreads();
            prefix[tid] += add;
            __syncthreads();
        }
        unsigned base = prefix[tid] - acc;
        for (int q = 0; q < PER; ++q) { const unsigned c0 = hist[tid * PER + q]; hist[tid * PER + q] = base; base += c0; }
        __syncthreads();  // hist[b] = number of good values in bins < b
    }
    const double pr[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
    if (tid < SEL_NT) {
        const int q = tid >> 1;
        const double hh = pr[q] * (double)m + 0.5;
        long long rk;
        if (hh <= 1.0) rk = 0;
        else if (hh >= (double)m) rk = m - 1;
        else rk = (long long)floor(hh) - 1 + (tid & 1);
        int a0 = 0, a1 = NB - 1;  // largest bin b with hist[b] <= rk
        while (a0 < a1) {
            const int mid = (a0 + a1 + 1) >> 1;
            if ((long long)hist[mid] <= rk) a0 = mid; else a1 = mid - 1;
        }
        tBin[tid] = a0;
        tRank[tid] = rk - hist[a0];
        gCnt[tid] = (int)(((a0 + 1 < NB) ? hist[a0 + 1] : (unsigned)m) - hist[a0]);
    }
    __syncthreads();
    if (tid == 0) {
        int nl = 0;
        for (int k = 0; k < SEL_NT; ++k) {
            int li = -1;
            for (int j = 0; j < nl; ++j)
                if (listBin[j] == tBin[k]) li = j;
            if (li < 0) {
                li = nl++;
                listBin[li] = tBin[k];
                listCnt[li] = 0;
                int am = 0;
                long long others = gCnt[k];
                for (int a = 0; a < nAtom; ++a)
                    if (s3_bin<NB>(atomV[a], lo, scale) == tBin[k]) { am |= 1 << a; others -= atomC[a]; }
                listAtoms[li] = am;
                listCnt[li] = (int)min(others, (long long)(SEL_NT * CAP + 1));  // members to gather (set to 0 again below)
                listMin[li] = 0xffffffffffffffffull;
                listMax[li] = 0ull;
                bitmap[tBin[k] >> 5] |= 1u << (tBin[k] & 31);
            }
            tList[k] = li;
        }
        // the SEL_NT x CAP doubles are ONE pool: every list gets the power of two (the sort pads to one) that holds its
        // members, in list order, while the pool lasts -- the ten targets usually share five bins, so a dense bin of a
        // peaked column may take several times CAP
        int used = 0;
        for (int j = 0; j < nl; ++j) {
            int cap = 2;
            while (cap < listCnt[j] && cap < SEL_NT * CAP) cap <<= 1;
            const bool fits = listCnt[j] <= cap && used + cap <= SEL_NT * CAP;
            listOver[j] = !fits;
            listOff[j] = used;
            if (fits) used += cap;
            listCnt[j] = 0;
        }
        nLists = nl;
    }
    __syncthreads();  // the histogram is dead from here on: its memory becomes the gather lists

    // ---- pass C: gather
    const int nl = nLists;
    s3_for_each<NTHR>(col, Nrep, [&](double v) {
        if (v == unfeas || v != v) return;
        if (v == at0 || v == at1) return;
        const int b = s3_bin<NB>(v, lo, scale);
        if (!((bitmap[b >> 5] >> (b & 31)) & 1u)) return;
        for (int j = 0; j < nl; ++j)
            if (listBin[j] == b) {
                if (listOver[j]) {
                    const unsigned long long key = s3_key(v);
                    atomicMin(&listMin[j], key);
                    atomicMax(&listMax[j], key);
                } else {
                    const int pos = atomicAdd(&listCnt[j], 1);
                    gath[listOff[j] + pos] = v;
                }
            }
    });
    __syncthreads();
    for (int j = 0; j < nl; ++j) {  // sort every list (bitonic, padded with +inf), whole block
        if (listOver[j]) continue;
        const int cnt = listCnt[j];
        int npad = 2;
        while (npad < cnt) npad <<= 1;
        double* g = gath + listOff[j];
        for (int i = cnt + tid; i < npad; i += NTHR) g[i] = INFINITY;
        __syncthreads();
        for (int kk = 2; kk <= npad; kk <<= 1)
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                for (int i = tid; i < npad; i += NTHR) {
                    const int ixj = i ^ jj;
                    if (ixj > i) {
                        const double a = g[i], b = g[ixj];
                        const bool up = ((i & kk) == 0);
                        if ((a > b) == up) { g[i] = b; g[ixj] = a; }
                    }
                }
                __syncthreads();
            }
    }
    if (tid < SEL_NT) {
        const int j = tList[tid];
        if (!listOver[j] && listAtoms[j] == 0) tVal[tid] = gath[listOff[j] + (int)tRank[tid]];
        else if (!listOver[j]) {  // rank inside (sorted other values of the bin) merged with the bin's atoms, ascending
            const double* g = gath + listOff[j];
            const int cnt = listCnt[j];
            long long r = tRank[tid], before = 0;  // before = atom copies that precede the answer
            double ans = NAN;
            bool done = false;
            for (int a = 0; a < nAtom && !done; ++a) {
                if (!((listAtoms[j] >> a) & 1)) continue;
                const double av = atomV[a];
                int l0 = 0, l1 = cnt;  // number of list values below the atom
                while (l0 < l1) { const int mid = (l0 + l1) >> 1; if (g[mid] < av) l0 = mid + 1; else l1 = mid; }
                const long long first = (long long)l0 + before;
                if (r < first) { ans = g[r - before]; done = true; }
                else if (r < first + atomC[a]) { ans = av; done = true; }
                else before += atomC[a];
            }
            if (!done) ans = (r - before < cnt) ? g[r - before] : NAN;
            tVal[tid] = ans;
        } else if (listMin[j] == listMax[j] && listAtoms[j] == 0) {  // every member of the bin is the same value
            const unsigned long long kx = listMin[j];
            const unsigned long long bits = (kx >> 63) ? (kx ^ 0x8000000000000000ull) : ~kx;
            tVal[tid] = __longlong_as_double((long long)bits);
        } else tVal[tid] = NAN;
    }
    __syncthreads();
    if (tid == 0) {
        bool slow = false;
        for (int k = 0; k < SEL_NT; ++k) slow |= (tVal[k] != tVal[k]);
        // 2 = "a selected bin holds more than CAP values that are not all equal": with <= 1.2e7 values a continuous column
        // never fills a bin like that, so these are clusters of ties next to other values (a growth index that is exactly
        // zero for part of the sample, Q = 0 below an activation stage): straight to the radix walk, the 4-read
        // histogram kernel fails on most of them when the sample is large (13 reads -> 9; Vegetation 10^7 x 1095 x 5: 990 -> 800 ms)
        needSlow[colId] = slow ? (Nrep > 6000000 ? 2 : 1) : 0;  // measured: below ~6e6 values the 4-read kernel still resolves most of them
        if (slow) return;
        for (int q = 0; q < 5; ++q) {
            const double hh = pr[q] * (double)m + 0.5;
            double qv;
            if (hh <= 1.0 || hh >= (double)m) qv = tVal[2 * q];
            else {
                const long long i = (long long)floor(hh);
                const double frac = hh - (double)i;
                qv = tVal[2 * q] + frac * (tVal[2 * q + 1] - tVal[2 * q]);
            }
            out[(size_t)q * nT] = qv;
        }
        out[(size_t)5 * nT] = mean;
        out[(size_t)6 * nT] = (m > 1) ? sqrt(ss / (double)(m - 1)) : unfeas;
    }
}

using PredKernel = void (*)(DevProblem, PredArgs, const double*, const int*, double* const*, const int*, double*);

PredKernel pick_pred(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return pred_main<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

}  // namespace

namespace bam {

void launch_predict(bamgpu_handle* h, int doParametric, const int* doRemnant, unsigned long long seed, int t0, int t1,
                    double* h_spag, bool withStates) {
    const DevProblem& p = h->dp;
    if (h->L.nVar) throw CudaError{"BaM_Prediction: prediction with VAR parameters is not allowed (BaM_Message 13)"};
    PredKernel kern = pick_pred(p.model, p.nX, p.nY);
    if (!kern) throw CudaError{"bamgpu: no prediction kernel instantiated for this (model, nX, nY)"};
    const long long Nrep = h->predNrep;
    // state variables (only SFD's kappa on this path) are extra output columns after the nY outputs: no remnant noise,
    // same envelope rule (BaM_Prediction :1071-1075,1121-1132)
    const int nTall = t1 - t0, nY = p.nY, nS = withStates ? model_nstate(p.model) : 0, nOut = nY + nS;
    if (nTall <= 0 || Nrep <= 0) return;
    bool anyRem = false;
    for (int y = 0; y < nY; ++y) anyRem |= doRemnant[y] != 0;
    if ((doParametric || anyRem) && !h->d_mc) throw CudaError{"BaM_Prediction: an MCMC sample is required"};
    if (!doParametric && !h->d_mode && p.nFit > 0) throw CudaError{"BaM_Prediction: maxpost vector required"};
    cudaStream_t s = h->stream;

    PredArgs a;
    a.Nrep = Nrep; a.nobsPred = h->predNobs; a.doParametric = doParametric; a.seed = seed;
    for (int y = 0; y < BAM_MAXY; ++y) a.doRemnant[y] = (y < nY) ? doRemnant[y] : 0;
    a.stride = p.nGamma + p.ntheta + p.nDer;
    a.nStateOut = nS;
    a.repBase = h->shardComm ? h->shardRepBase : 0;
    h->predNout = nOut;
    a.stage = sizeof(double) * (size_t)(a.stride | 1) * BAM_BLOCK <= 96 * 1024;

    // table + envelope output
    if (!h->d_ptab) {
        BAM_CUDA(cudaMalloc(&h->d_ptab, sizeof(double) * (size_t)Nrep * a.stride));
        BAM_CUDA(cudaMalloc(&h->d_pstatus, sizeof(int) * (size_t)Nrep));
    }
    const size_t envNeed = (size_t)nTall * BAMGPU_NENV * nOut;
    if (envNeed > h->envCap) {
        if (h->d_env) cudaFree(h->d_env);
        BAM_CUDA(cudaMalloc(&h->d_env, sizeof(double) * envNeed));
        h->envCap = envNeed;
    }
    h->predT0 = t0; h->predT1 = t1;

    // tile of inputs: the materialised values take at most a third of the free HBM (capped at 48 GiB).  The envelope
    // kernels run one block per column, so a tile should hold several times 148 columns when Nrep is large.
    size_t freeB = 0, totalB = 0;
    BAM_CUDA(cudaMemGetInfo(&freeB, &totalB));
    const size_t budget = std::max<size_t>((size_t)1 << 30, std::min<size_t>((freeB + h->vCap * sizeof(double)) / 3, (size_t)48 << 30));
    long long tileT = (long long)(budget / (sizeof(double) * (size_t)Nrep * nOut));
    tileT = std::max<long long>(32, std::min<long long>(tileT / 32 * 32, ((long long)nTall + 31) / 32 * 32));
    if (h_spag) tileT = std::max<long long>(tileT, 32);
    if (h->shardComm) tileT = std::min<long long>(tileT, std::max<long long>(32, (1024 / nOut) / 32 * 32));  // 80 KB of histograms per column
    else if (Nrep > 16384 && tileT < nTall) {
        // the selection kernels run one block per column, 1 (large-sample instantiation, 160 KB of shared memory) or 3 per SM:
        // cut the tile so that its columns fill whole waves (128 inputs x 5 outputs = 640 blocks were 4.3 waves on 148 SMs)
        const long long slots = (long long)h->smCount * (Nrep > 3000000 ? 1 : 3);
        long long best = tileT;
        double bestEff = 0.0;
        for (long long t = tileT; t >= std::max<long long>(32, tileT * 7 / 10); --t) {
            const double w = (double)(t * nOut) / (double)slots;
            const double eff = w / std::ceil(w);
            if (eff > bestEff + 1e-9) { bestEff = eff; best = t; }
        }
        tileT = best;
    }
    const size_t vNeed = (size_t)tileT * Nrep * nOut;
    if (vNeed > h->vCap) {
        if (h->d_V) cudaFree(h->d_V);
        BAM_CUDA(cudaMalloc(&h->d_V, sizeof(double) * vNeed));
        h->vCap = vNeed;
    }

    // side buffer of the deferred radix columns (stash_assign, stash_copy): at most 64 columns / a sixteenth of the free memory
    // (64 columns when memory allows: kept with the handle like the tile buffer, released with it)
    int sideCap = 0;
    int *d_sideCount = nullptr, *d_sideMap = nullptr;
    if (Nrep > 16384 && !h->shardComm) {
        const size_t want = 64;
        if (h->sideCap < want * (size_t)Nrep) {
            size_t fb = 0, tb = 0;
            BAM_CUDA(cudaMemGetInfo(&fb, &tb));
            const size_t cols = std::min<size_t>(want, (fb / 16) / (sizeof(double) * (size_t)Nrep));
            if (cols >= 4 && cols * (size_t)Nrep > h->sideCap) {
                if (h->d_side) cudaFree(h->d_side);
                h->d_side = nullptr; h->sideCap = 0;
                BAM_CUDA(cudaMalloc(&h->d_side, sizeof(double) * cols * (size_t)Nrep));
                h->sideCap = cols * (size_t)Nrep;
            }
        }
        sideCap = (int)std::min<size_t>(want, h->sideCap / (size_t)Nrep);
        if (sideCap < 4) sideCap = 0;
        if (sideCap) {
            BAM_CUDA(cudaMallocAsync(&d_sideCount, sizeof(int), s));
            BAM_CUDA(cudaMallocAsync(&d_sideMap, sizeof(int) * 2 * sideCap, s));
            BAM_CUDA(cudaMemsetAsync(d_sideCount, 0, sizeof(int), s));
        }
    }
    pred_prep<<<(unsigned)((Nrep + 127) / 128), 128, 0, s>>>(p, a, h->d_mc, h->d_mode, h->d_xspagPtrs, h->d_nspag, h->d_ptab,
                                                             h->d_pstatus);
    h->launches += 1;
    DevProblem pser = p;
    if (p.isSeries) {  // time-recursive model: every replication's whole series (outputs + states) first, then the generic
                       // kernel looks the values up, adds the structural error and fills the tile
        const int nOutS = nY + model_nstate(p.model);
        if ((double)h->predNobs * nOutS * (double)Nrep * 8.0 > 64e9) throw CudaError{"bamgpu: time-recursive prediction too large for one series buffer"};
        ensure_series_capacity(h, (size_t)h->predNobs * nOutS * (size_t)Nrep);
        pser.seriesX = h->d_seriesXpred; pser.seriesN = h->predNobs; pser.seriesK = (int)Nrep; pser.seriesY = h->d_series;
        launch_series(h, pser, (int)Nrep, h->d_ptab, a.stride, p.nGamma, h->d_pstatus, nOutS);
    }
    const size_t smemMain = sizeof(double) * (2 * BAM_LOGTAB_N + 2 * BAM_SCTAB_N + BAM_EXPTAB_N) + (a.stage ? sizeof(double) * (size_t)(a.stride | 1) * BAM_BLOCK : 0);
    if (smemMain > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemMain));

    int npad = 1;
    while (npad < Nrep && npad < (1 << 30)) npad <<= 1;
    const bool smallN = Nrep <= 16384;
    const size_t smemEnv = sizeof(double) * ((size_t)npad + 32);
    if (smallN && smemEnv > 48 * 1024)
        BAM_CUDA(cudaFuncSetAttribute((const void*)envelope_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemEnv));

    const size_t smemRadix = sizeof(unsigned) * SEL_NT * RX_NB;
    BAM_CUDA(cudaFuncSetAttribute((const void*)envelope_radix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRadix));
    if (h->shardComm) BAM_CUDA(cudaFuncSetAttribute((const void*)rxs_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemRadix));
    const size_t smemSel2 = sizeof(unsigned) * SEL_NT * S2_NB2 + sizeof(double) * SEL_NT * S2_CAP;
    // the 2-read selection resolves bins of <= CAP members: worth trying while the densest of its 16384 bins stays below
    // that (a Gaussian column peaks at ~1.1e-4 Nrep per bin); two instantiations, see envelope_select3
    const bool sel3Big = Nrep > 3000000;
    const size_t smemSel3 = std::max(sizeof(unsigned) * 32 * (sel3Big ? 1024 : 512), sizeof(double) * SEL_NT * (sel3Big ? 2048 : 512));
    const bool trySel3 = !smallN && Nrep <= 12000000 && !h->noSelect3;
    if (!smallN) {
        BAM_CUDA(cudaFuncSetAttribute((const void*)envelope_select2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemSel2));
        if (sel3Big) BAM_CUDA(cudaFuncSetAttribute((const void*)envelope_select3<2048, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemSel3));
        else BAM_CUDA(cudaFuncSetAttribute((const void*)envelope_select3<512, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemSel3));
    }
    PredFastKernel fastK = (!h->forceGeneric && p.nX == 1 && p.nY == 1) ? pick_pred_fast(p, anyRem, Nrep) : nullptr;
    // Sediment fast path: pseudo-parameters FIX / IN, fixed css coefficients, one input series for all replications
    SedFastKernel sedK = nullptr;
    SedObs* d_sedObs = nullptr;
    if (!h->forceGeneric && p.model == BAMGPU_MDL_SEDIMENT && Nrep >= 1024 && p.sedCss != BAMGPU_CSS_CONSTANT &&
        p.sedCo == BAMGPU_CO_FIX && (doParametric || h->d_mode)) {
        bool okPath = true;
        for (int i = 0; i < 4; ++i) okPath &= p.sedPpo[i] != BAMGPU_PPO_PAR;
        for (int i = 0; i < p.nX; ++i) okPath &= h->predNspag[i] == 1;
        if (okPath) sedK = pick_sed_fast(p.sigfunc[0], anyRem);
        if (sedK) {
            BAM_CUDA(cudaMallocAsync(&d_sedObs, sizeof(SedObs) * (size_t)nTall, s));
            pick_sed_prep(p.nX)<<<(nTall + 127) / 128, 128, 0, s>>>(p, h->predNobs, t0, nTall, h->d_xspagPtrs, d_sedObs);
            h->launches += 1;
        }
    }
    a.tBase = t0;
    BAM_CUDA(cudaEventRecord(h->ev0, s));
    // measurement: per tile two bracketed intervals, [propagation kernel] and [envelope kernels]
    auto mark = [&]() -> cudaEvent_t {
        if (!h->recordKernels || h->kernEv.size() >= 8192) return nullptr;
        cudaEvent_t e = take_event(h);
        BAM_CUDA(cudaEventRecord(e, s));
        return e;
    };
    auto stash = [&](int* flags, int ncol, int nT, int tile0) {  // slow columns of this tile -> side buffer (see stash_assign)
        int* slotOf = nullptr;
        BAM_CUDA(cudaMallocAsync(&slotOf, sizeof(int) * (size_t)ncol, s));
        stash_assign<<<(ncol + 255) / 256, 256, 0, s>>>(ncol, nT, tile0, flags, sideCap, d_sideCount, d_sideMap, slotOf);
        stash_copy<<<dim3(ncol, STASH_SPLIT), 256, 0, s>>>(Nrep, h->d_V, slotOf, h->d_side);
        BAM_CUDA(cudaFreeAsync(slotOf, s));
        h->launches += 2;
    };
    for (int tt = 0; tt < nTall; tt += (int)tileT) {
        const int nT = std::min<long long>(tileT, nTall - tt);
        a.t0 = t0 + tt; a.nT = nT;
        cudaEvent_t kp0 = mark();
        if (sedK) {
            dim3 grid((unsigned)((Nrep + BAM_BLOCK - 1) / BAM_BLOCK), (nT + PF_TT - 1) / PF_TT);
            sedK<<<grid, BAM_BLOCK, sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N + 2 * BAM_SCTAB_N) + sizeof(SedObs) * PF_TT, s>>>(
                p, a, h->d_mc, h->d_mode, d_sedObs, h->d_V);
        } else if (fastK) {
            dim3 grid((unsigned)((Nrep + BAM_BLOCK - 1) / BAM_BLOCK), (nT + PF_TT - 1) / PF_TT);
            fastK<<<grid, BAM_BLOCK, PF_SMEM, s>>>(p, a, h->d_ptab, h->d_pstatus, h->d_xspag[0], h->predNspag[0], h->d_V);
        } else {
            dim3 grid((unsigned)((Nrep + BAM_BLOCK - 1) / BAM_BLOCK), (nT + 31) / 32);
            kern<<<grid, BAM_BLOCK, smemMain, s>>>(pser, a, h->d_ptab, h->d_pstatus, h->d_xspagPtrs, h->d_nspag, h->d_V);
        }
        h->launches += 1;
        cudaEvent_t kp1 = mark();
        if (kp0 && kp1) h->kernEv.emplace_back(kp0, kp1);
        if (Nrep >= 2 || (h->shardComm && h->shardNrepTotal >= 2)) {
            cudaEvent_t ke0 = mark();
            // envelopes of this tile go to a compact [y][7][nT] scratch, then into the full [y][7][nTall] array
            double* envTile = nullptr;
            BAM_CUDA(cudaMallocAsync(&envTile, sizeof(double) * (size_t)nT * BAMGPU_NENV * nOut, s));
            if (h->shardComm) {  // sample-sharded: radix walk over histograms summed across the ranks
                const int ncol = nT * nOut;
                RxsState* st = nullptr;
                unsigned* hist = nullptr;
                double *acc = nullptr, *ssq = nullptr;
                BAM_CUDA(cudaMallocAsync(&st, sizeof(RxsState) * ncol, s));
                BAM_CUDA(cudaMallocAsync(&hist, sizeof(unsigned) * (size_t)ncol * SEL_NT * RX_NB, s));
                BAM_CUDA(cudaMallocAsync(&acc, sizeof(double) * 2 * ncol, s));
                BAM_CUDA(cudaMallocAsync(&ssq, sizeof(double) * ncol, s));
                rxs_init<<<(ncol + 127) / 128, 128, 0, s>>>(st, ncol);
                char cm[BAMGPU_MESS_LEN];
                auto reduce = [&](void* buf, size_t count, int type) {
                    if (bamgpu_comm_allreduce_sum(h->shardComm, buf, count, type, (void*)s, cm) != 0) throw CudaError{std::string(cm)};
                };
                for (int pass = 0; pass < 6; ++pass) {
                    rxs_hist<<<ncol, RX_THREADS, smemRadix, s>>>(Nrep, p.unfeas, h->d_V, st, pass, hist, acc);
                    reduce(hist, (size_t)ncol * SEL_NT * RX_NB, 2);
                    if (pass == 0) reduce(acc, (size_t)2 * ncol, 1);
                    rxs_walk<<<ncol, 32 * SEL_NT, 0, s>>>(h->shardNrepTotal, st, hist, acc, pass);
                    h->launches += 2;
                }
                rxs_ss<<<ncol, RX_THREADS, 0, s>>>(Nrep, p.unfeas, h->d_V, st, ssq);
                reduce(ssq, (size_t)ncol, 1);
                rxs_finish<<<(ncol + 127) / 128, 128, 0, s>>>(ncol, nT, p.unfeas, st, ssq, envTile);
                h->launches += 3;
                BAM_CUDA(cudaFreeAsync(st, s)); BAM_CUDA(cudaFreeAsync(hist, s)); BAM_CUDA(cudaFreeAsync(acc, s)); BAM_CUDA(cudaFreeAsync(ssq, s));
            } else if (smallN) envelope_smem<<<nT * nOut, BAM_BLOCK, smemEnv, s>>>(Nrep, npad, nT, p.unfeas, h->d_V, envTile);
            else if (h->onlyRadix) {  // tests: every column through the last-resort kernel -- the first ones by way of the
                                      // side buffer (deferred launch), the rest inside their tile
                int* all = nullptr;
                BAM_CUDA(cudaMallocAsync(&all, sizeof(int) * (size_t)nT * nOut, s));
                BAM_CUDA(cudaMemsetAsync(all, 1, sizeof(int) * (size_t)nT * nOut, s));  // 0x01010101: "needs the slow path"
                if (sideCap) {
                    stash(all, nT * nOut, nT, tt);
                }
                envelope_radix<<<nT * nOut, RX_THREADS, smemRadix, s>>>(Nrep, nT, p.unfeas, h->d_V, envTile, all);
                BAM_CUDA(cudaFreeAsync(all, s));
            } else {
                int *need3 = nullptr, *needSlow = nullptr;
                BAM_CUDA(cudaMallocAsync(&needSlow, sizeof(int) * (size_t)nT * nOut, s));
                if (trySel3) {
                    BAM_CUDA(cudaMallocAsync(&need3, sizeof(int) * (size_t)nT * nOut, s));
                    if (sel3Big) envelope_select3<2048, 1024><<<nT * nOut, 1024, smemSel3, s>>>(Nrep, nT, p.unfeas, h->d_V, envTile, need3);
                    else envelope_select3<512, 512><<<nT * nOut, 512, smemSel3, s>>>(Nrep, nT, p.unfeas, h->d_V, envTile, need3);
                    h->launches += 1;
                }
                envelope_select2<<<nT * nOut, S2_THREADS, smemSel2, s>>>(Nrep, nT, p.unfeas, h->d_V, envTile, needSlow, need3);
                if (sideCap) {
                    stash(needSlow, nT * nOut, nT, tt);
                }
                envelope_radix<<<nT * nOut, RX_THREADS, smemRadix, s>>>(Nrep, nT, p.unfeas, h->d_V, envTile, needSlow);
                BAM_CUDA(cudaFreeAsync(needSlow, s));
                if (need3) BAM_CUDA(cudaFreeAsync(need3, s));
                h->launches += 1;
            }
            h->launches += 1;
            cudaEvent_t ke1 = mark();
            if (ke0 && ke1) h->kernEv.emplace_back(ke0, ke1);
            for (int y = 0; y < nOut; ++y)
                BAM_CUDA(cudaMemcpy2DAsync(h->d_env + ((size_t)y * BAMGPU_NENV * nTall + tt), sizeof(double) * nTall,
                                           envTile + (size_t)y * BAMGPU_NENV * nT, sizeof(double) * nT,
                                           sizeof(double) * nT, BAMGPU_NENV, cudaMemcpyDeviceToDevice, s));
            BAM_CUDA(cudaFreeAsync(envTile, s));
        }
        if (h_spag) {  // host layout: [y][t][rep] over the whole [t0,t1) range
            for (int y = 0; y < nOut; ++y)
                BAM_CUDA(cudaMemcpyAsync(h_spag + ((size_t)y * nTall + tt) * Nrep, h->d_V + (size_t)y * nT * Nrep,
                                         sizeof(double) * (size_t)nT * Nrep, cudaMemcpyDeviceToHost, s));
            BAM_CUDA(cudaStreamSynchronize(s));
        }
    }
    if (sideCap) {  // the deferred columns, all at once, straight into the full envelope array
        // (accounted with the envelope kernels: an empty propagation interval keeps the [propagation, envelopes] pairing;
        // four distinct events, every event belongs to one interval)
        cudaEvent_t ka = mark(), kb = mark(), kd0 = mark();
        envelope_radix<<<sideCap, RX_THREADS, smemRadix, s>>>(Nrep, nTall, p.unfeas, h->d_side, h->d_env, nullptr, d_sideMap, d_sideCount);
        cudaEvent_t kd1 = mark();
        if (ka && kb && kd0 && kd1) { h->kernEv.emplace_back(ka, kb); h->kernEv.emplace_back(kd0, kd1); }
        h->launches += 1;
        BAM_CUDA(cudaFreeAsync(d_sideCount, s)); BAM_CUDA(cudaFreeAsync(d_sideMap, s));
    }
    if (d_sedObs) BAM_CUDA(cudaFreeAsync(d_sedObs, s));
    BAM_CUDA(cudaEventRecord(h->ev1, s));
    BAM_CUDA(cudaGetLastError());
}

}  // namespace bam
