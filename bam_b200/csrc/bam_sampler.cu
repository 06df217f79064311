// bam_sampler.cu -- K3: batched one-at-a-time adaptive Metropolis, K chains in lockstep.
//
// Replaces BMSL's sequential Adaptive_Metro_OAAT as called by BaM_Fit (src/BaM_tools.f90:615-621).
// Declared semantics (BMSL is un-vendored): for every cycle { moves=0; nAdapt
// iterations of { for each component i: candidate = x + std_i*N(0,1) on component i; accept with
// probability min(1, exp(f(cand)-f(x))) if feasible and non-null }; one row [x, f(x), Dpar] per
// iteration; then std_i *= DownMult if rate_i <= MinMoveRate, *= UpMult if rate_i >= MaxMoveRate }.
// RNG = Philox4x32-10 keyed by the seed, counter (global chain index, iteration, 2*i | 2*i+1): a
// chain's trajectory does not depend on K, on the GPU count or on launch geometry.
//
// K5: Welford running moments {count, mean, M2} per (chain, component) over the kept rows; they feed
// the Gelman-Rubin diagnostic (all-gathered across ranks by the caller, NCCL).
#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"

namespace {

__global__ void propose_kernel(int K, int P, int comp, unsigned long long seed, long long chainOffset, unsigned it,
                               const double* __restrict__ sd, double* __restrict__ delta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    delta[c] = sd[(size_t)c * P + comp] * philox_normal(seed, (uint64_t)(chainOffset + c), it, 2u * (unsigned)comp);
}

__global__ void accept_kernel(int K, int P, int nDpar, int comp, unsigned long long seed, long long chainOffset,
                              unsigned it, const double* __restrict__ delta, const double* __restrict__ fxCand,
                              const int* __restrict__ stCand, const double* __restrict__ dparCand, double* __restrict__ x,
                              double* __restrict__ fxCur, double* __restrict__ dparCur, int* __restrict__ moves,
                              int nCombo, const double* __restrict__ lkCcand, double* __restrict__ lkCcur) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (stCand[c] != 0) return;  // infeasible or null candidate: rejected
    const double u = philox_uniform(seed, (uint64_t)(chainOffset + c), it, 2u * (unsigned)comp + 1u);
    if (log(u) < fxCand[c] - fxCur[c]) {
        x[(size_t)c * P + comp] += delta[c];
        fxCur[c] = fxCand[c];
        for (int d = 0; d < nDpar; ++d) dparCur[(size_t)c * nDpar + d] = dparCand[(size_t)c * nDpar + d];
        moves[(size_t)c * P + comp] += 1;
        // per-combination likelihood sums of the accepted state (fast path with partial re-evaluation)
        if (lkCcur != nullptr)
            for (int v = 0; v < nCombo; ++v) lkCcur[(size_t)v * K + c] = lkCcand[(size_t)v * K + c];
    }
}

__global__ void init_state_kernel(int K, int nDpar, const double* __restrict__ fx, const double* __restrict__ dpar,
                                  const int* __restrict__ st, double* __restrict__ fxCur, double* __restrict__ dparCur,
                                  int* __restrict__ nBad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    fxCur[c] = fx[c];
    for (int d = 0; d < nDpar; ++d) dparCur[(size_t)c * nDpar + d] = dpar[(size_t)c * nDpar + d];
    if (st[c] != 0) atomicAdd(nBad, 1);
}

// one kept iteration: store the row and update the running moments; one thread per (chain, column)
__global__ void record_kernel(int K, int P, int nDpar, long long keepIdx, long long nKeep, const double* __restrict__ x,
                              const double* __restrict__ fxCur, const double* __restrict__ dparCur,
                              double* __restrict__ rows, double* __restrict__ mcount, double* __restrict__ mmean,
                              double* __restrict__ mm2) {
    const int rowlen = P + 1 + nDpar;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)K * rowlen) return;
    const int c = (int)(q / rowlen), j = (int)(q % rowlen);
    double v;
    if (j < P) v = x[(size_t)c * P + j];
    else if (j == P) v = fxCur[c];
    else v = dparCur[(size_t)c * nDpar + (j - P - 1)];
    if (rows) rows[((size_t)c * nKeep + keepIdx) * rowlen + j] = v;
    if (j < P) {  // Welford
        const double cnt = (double)(keepIdx + 1);
        const size_t o = (size_t)c * P + j;
        const double m = mmean[o], d = v - m, mn = m + d / cnt;
        mmean[o] = mn;
        mm2[o] += d * (v - mn);
        if (j == 0) mcount[c] = cnt;
    }
}

__global__ void adapt_kernel(int K, int P, int nAdapt, double minMR, double maxMR, double downMult, double upMult,
                             double* __restrict__ sd, int* __restrict__ moves) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)K * P) return;
    const double rate = (double)moves[q] / (double)nAdapt;
    double s = sd[q];
    if (rate <= minMR) s *= downMult;
    if (rate >= maxMR) s *= upMult;
    sd[q] = s;
    moves[q] = 0;
}


// ------------------------------------------------------------------------------------------------
// K3p: persistent sampler for small data sets (the reference's own workspaces have 38..812 observations).
//
// With few observations a launch per proposal is pure launch latency, so the whole Metropolis loop runs inside
// one kernel: ONE WARP PER CHAIN, lanes stride over the observations (and over parameters for the priors),
// the chain's vector / table / jump scales live in shared memory, accept / reject, move counters, scale
// adaptation, row recording and the Welford moments all happen in the kernel.  One launch covers a chunk of
// iterations (a cycle), so the host can keep the reference's .monitor file alive between launches.
// Same RNG streams and the same accept rule as the launch-per-proposal path: both give the same chains up to
// the rounding of the (differently ordered) sums.
// ------------------------------------------------------------------------------------------------
#define WS_WARPS 4       /* chains per block */
#define WS_MAXP 64       /* vector length limit of this path (no latent inputs) */

struct WarpSamplerArgs {
    int K, nIterChunk, nAdapt;
    long long it0, burn, keepEvery, nKeep;
    double minMR, maxMR, downMult, upMult;
    unsigned long long seed;
    long long chainOffset;
    double *x, *sd, *fxcur, *dparcur, *rows, *mcount, *mmean, *mm2;
    int* moves;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// log-posterior of the vector xs (shared memory) for one chain, computed by one warp.  Returns status bits.
template <int MODEL, int NX, int NY>
__device__ int warp_logpost(const DevProblem& p, const double* xs, double* tab, double* vcop, double& fx) {
    const int lane = threadIdx.x & 31;
    int st = 0;
    // ---- prior (GetPriorLogPdf :3006-3138): parameters spread over lanes
    double pr = 0.0;
    const int npar = p.nFit + p.nGamma;
    for (int k = lane; k < npar; k += 32) {
        double lp;
        bool feas, isnull;
        if (k < p.nFit) dist_logpdf(p.ptDist[k], p.ptPar + (size_t)k * BAMGPU_PRIOR_NPAR, xs[k], lp, feas, isnull);
        else dist_logpdf(p.pgDist[k - p.nFit], p.pgPar + (size_t)(k - p.nFit) * BAMGPU_PRIOR_NPAR, xs[k], lp, feas, isnull);
        if (!feas) st |= ST_PRIOR_INFEAS;
        else if (isnull) st |= ST_PRIOR_NULL;
        else pr += lp;
    }
    if (p.priorM != nullptr) {
        for (int k = lane; k < npar; k += 32) {
            double u;
            bool f;
            if (k < p.nFit) dist_cdf(p.ptDist[k], p.ptPar + (size_t)k * BAMGPU_PRIOR_NPAR, xs[k], u, f);
            else dist_cdf(p.pgDist[k - p.nFit], p.pgPar + (size_t)(k - p.nFit) * BAMGPU_PRIOR_NPAR, xs[k], u, f);
            if (!f || !(u > 0.0 && u < 1.0)) { st |= ST_PRIOR_INFEAS; vcop[k] = 0.0; }
            else vcop[k] = normcdfinv(u);
        }
        __syncwarp();
        double q = 0.0;
        for (int cc = lane; cc < npar; cc += 32) {
            double w = 0.0;
            for (int r = 0; r < npar; ++r) w = w + vcop[r] * p.priorM[r + (size_t)cc * npar];
            q += w * vcop[cc];
        }
        pr -= 0.5 * q;
    }
    for (int j = 0; j < p.nX; ++j)
        for (int i = lane; i < p.nXb[j]; i += 32) pr += gauss_logpdf(xs[p.offXb[j] + i], 0.0, 1.0);
    for (int j = 0; j < p.nY; ++j)
        for (int i = lane; i < p.nYb[j]; i += 32) pr += gauss_logpdf(xs[p.offYb[j] + i], 0.0, 1.0);
    pr = warp_sum(pr);
    // ---- per-chain table (as logpost_prep)
    for (int g = lane; g < p.nGamma; g += 32) tab[p.tabGamma + g] = xs[p.nFit + g];
    for (int j = 0; j < p.nX; ++j)
        for (int i = lane; i < p.nXb[j]; i += 32) tab[p.tabOffXb[j] + i] = xs[p.offXb[j] + i];
    for (int j = 0; j < p.nY; ++j)
        for (int i = lane; i < p.nYb[j]; i += 32) tab[p.tabOffYb[j] + i] = xs[p.offYb[j] + i];
    for (int k = lane; k < p.nCombo; k += 32) {
        double th[BAM_MAXTHETA], der[BAM_MAXDER];
        const double* xc[BAM_MAXX];
        for (int j = 0; j < p.nX; ++j) xc[j] = p.X + (size_t)j * p.n;
        double* cr = tab + p.tabCombo + (size_t)k * p.comboStride;
        for (int i = 0; i < p.ntheta; ++i) {
            const int src = p.comboSrc[(size_t)k * p.ntheta + i];
            th[i] = (src < 0) ? p.fixv[i] : xs[src];
            cr[i] = th[i];
        }
        for (int d = 0; d < p.nDer; ++d) der[d] = p.unfeas;
        if (!model_prepare(p, th, der, xc, p.n)) st |= ST_LKH_INFEAS;
        for (int d = 0; d < p.nDer; ++d) cr[p.ntheta + d] = der[d];
    }
    st = __reduce_or_sync(0xffffffffu, st);
    __syncwarp();
    double lk = 0.0;
    if (!st) {
        // ---- likelihood (GetLogLikelihood :3142-3229): lanes stride over the observations
        const int n = p.n;
        bool bad = false;
        for (int i = lane; i < n; i += 32) {
            double xt[NX], yh[NY];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const size_t q = i + (size_t)j * n;
                double v = p.X[q];
                if (p.hasXbias) {
                    const int bi = p.Xbindx[q];
                    if (bi > 0) v = v - p.Xb[q] * tab[p.tabOffXb[j] + bi - 1];
                }
                xt[j] = v;
            }
            const int combo = (p.nCombo > 1) ? p.comboOf[i] : 0;
            const double* th = tab + p.tabCombo + (size_t)combo * p.comboStride;
            if (!model_apply<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh)) { bad = true; continue; }
#pragma unroll
            for (int j = 0; j < NY; ++j) {
                const size_t q = i + (size_t)j * n;
                const double yo = p.Y[q];
                if (p.hasMissing && yo == p.mv) continue;
                const double sig = sigmafunk(p.sigfunc[j], tab + p.tabGamma + p.gamOff[j], yh[j]);
                if (!(sig > 0.0)) { bad = true; continue; }
                const double yu = p.Yu[q];
                double mu = yh[j];
                if (p.hasYbias) {
                    const int bi = p.Ybindx[q];
                    if (bi != 0) mu = yh[j] + p.Yb[q] * tab[p.tabOffYb[j] + bi - 1];
                }
                lk += gauss_logpdf_var(yo, mu, fma(yu, yu, sig * sig));
            }
        }
        if (__any_sync(0xffffffffu, bad)) st |= ST_LKH_INFEAS;
        lk = warp_sum(lk);
        if (!(lk == lk)) st |= ST_LKH_INFEAS;
    }
    fx = st ? BAMGPU_UNDEF_RN : pr + lk;
    return st;
}

template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(32 * WS_WARPS) sampler_warp(DevProblem p, WarpSamplerArgs a) {
    extern __shared__ double wsm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * WS_WARPS + w;
    if (c >= a.K) return;
    const int P = p.P, nD = p.nDpar;
    const int perWarp = 2 * P + p.tabStride + p.kM + 2;  // xs | sd | tab | copula v
    double* xs = wsm + (size_t)w * perWarp;
    double* sd = xs + P;
    double* tab = sd + P;
    double* vcop = tab + p.tabStride;
    for (int j = lane; j < P; j += 32) { xs[j] = a.x[(size_t)c * P + j]; sd[j] = a.sd[(size_t)c * P + j]; }
    int mv0 = (lane < P) ? a.moves[(size_t)c * P + lane] : 0;           // move counter of component lane
    int mv1 = (lane + 32 < P) ? a.moves[(size_t)c * P + lane + 32] : 0; // and of component lane+32
    double fxc = a.fxcur[c];
    __syncwarp();
    const unsigned long long gc = (unsigned long long)(a.chainOffset + c);
    for (int itl = 0; itl < a.nIterChunk; ++itl) {
        const long long it = a.it0 + itl;
        for (int i = 0; i < P; ++i) {
            const double old = xs[i];
            const double delta = sd[i] * philox_normal(a.seed, gc, (unsigned)it, 2u * (unsigned)i);
            __syncwarp();
            if (lane == 0) xs[i] = old + delta;
            __syncwarp();
            double fc;
            const int st = warp_logpost<MODEL, NX, NY>(p, xs, tab, vcop, fc);
            bool acc = false;
            if (st == 0) {
                const double u = philox_uniform(a.seed, gc, (unsigned)it, 2u * (unsigned)i + 1u);
                acc = log(u) < fc - fxc;
            }
            if (acc) {
                fxc = fc;
                if (lane == (i & 31)) { if (i < 32) ++mv0; else ++mv1; }
                // derived parameters of the accepted state: the b's sit in the table
                for (int d = lane; d < nD; d += 32) {
                    const int k = d / p.nDparModel, q = d % p.nDparModel;
                    a.dparcur[(size_t)c * nD + d] = tab[p.tabCombo + (size_t)k * p.comboStride + p.ntheta + q];
                }
            } else {
                __syncwarp();
                if (lane == 0) xs[i] = old;
            }
            __syncwarp();
        }
        const long long it1 = it + 1;
        if (it1 > a.burn && (it1 - a.burn) % a.keepEvery == 0) {
            const long long kept = (it1 - a.burn) / a.keepEvery - 1;
            if (kept < a.nKeep) {
                const int rowlen = P + 1 + nD;
                if (a.rows) {
                    double* row = a.rows + ((size_t)c * a.nKeep + kept) * rowlen;
                    for (int j = lane; j < P; j += 32) row[j] = xs[j];
                    if (lane == 0) row[P] = fxc;
                    for (int d = lane; d < nD; d += 32) row[P + 1 + d] = a.dparcur[(size_t)c * nD + d];
                }
                const double cnt = (double)(kept + 1);
                for (int j = lane; j < P; j += 32) {  // Welford
                    const size_t o = (size_t)c * P + j;
                    const double v = xs[j], m = a.mmean[o], d = v - m, mn = m + d / cnt;
                    a.mmean[o] = mn;
                    a.mm2[o] += d * (v - mn);
                }
                if (lane == 0) a.mcount[c] = cnt;
            }
        }
        if (it1 % a.nAdapt == 0) {  // end of a cycle: scale adaptation (App. E)
            for (int j = lane; j < P; j += 32) {
                const int mvj = (j < 32) ? mv0 : mv1;
                const double rate = (double)mvj / (double)a.nAdapt;
                double s = sd[j];
                if (rate <= a.minMR) s *= a.downMult;
                if (rate >= a.maxMR) s *= a.upMult;
                sd[j] = s;
            }
            mv0 = 0;
            mv1 = 0;
            __syncwarp();
        }
    }
    for (int j = lane; j < P; j += 32) { a.x[(size_t)c * P + j] = xs[j]; a.sd[(size_t)c * P + j] = sd[j]; }
    if (lane < P) a.moves[(size_t)c * P + lane] = mv0;
    if (lane + 32 < P) a.moves[(size_t)c * P + lane + 32] = mv1;
    if (lane == 0) a.fxcur[c] = fxc;
}

// ------------------------------------------------------------------------------------------------
// K3l: one-launch sweep over ALL latent "true input" components (cfg 3: 2 x 10^6 per chain).
//
// Every model on this path is pointwise, so given (theta, gamma, biases) the posterior factorises over
// observations: the latent inputs of observation i only enter its own likelihood terms and its own hyper terms
// (GetLogLikelihood :3204-3227, GetHyperLogPdf :3285-3303).  The one-at-a-time updates of latent components that
// belong to DIFFERENT observations therefore commute, and
//     f(candidate) - f(current) = [lkh_i + hyper_i](candidate) - [lkh_i + hyper_i](current)
// needs O(1) work instead of a pass over all observations.  thread <-> (chain, observation); the latent inputs of
// one observation are visited in column order, which is exactly the order in which the reference's vector layout
// (Packer :4202-4242: column-major over (i,j)) visits them as far as observation i can tell.  Same Philox streams
// as every other path (normal: (chain, it, 2 comp), uniform: (chain, it, 2 comp + 1)), same accept rule.
// Per latent component: 8 B read + 8 B written of state, 8 B jump scale, 8 B move counter: HBM-leaning FP64 work.
// ------------------------------------------------------------------------------------------------
template <int MODEL, int NX, int NY>
__device__ __forceinline__ bool local_logpost(const DevProblem& p, const double* xt, const double* xo, const double* xu,
                                              const double* xb, const int* xbi, const double* yo, const double* yu,
                                              const double* yb, const int* ybi, const double* hyC, const double* hyR,
                                              const double* __restrict__ row, const double* __restrict__ th, unsigned lt,
                                              double& val) {
    double yh[NY];
    if (!model_apply<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh)) return false;
    double lk = 0.0;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        if (p.hasMissing && yo[j] == p.mv) continue;
        const double sig = sigmafunk(p.sigfunc[j], row + p.tabGamma + p.gamOff[j], yh[j]);
        if (!(sig > 0.0)) return false;
        double mu = yh[j];
        if (p.hasYbias && ybi[j] != 0) mu = yh[j] + yb[j] * row[p.tabOffYb[j] + ybi[j] - 1];
        lk += gauss_logpdf_var_t(yo[j], mu, fma(yu[j], yu[j], sig * sig), lt);
    }
    double hy = 0.0;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        if (!(xu[j] > 0.0)) continue;
        double mu = xt[j];
        if (p.hasXbias && xbi[j] != 0) mu = xt[j] + xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
        const double z = (xo[j] - mu) * hyR[j];
        hy += fma(-0.5 * z, z, hyC[j]);  // hyC = -log sqrt(2 pi) - log xu, hyR = 1 / xu: per observation, not per proposal
    }
    val = lk + hy;
    return val == val;
}

template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(BAM_BLOCK) latent_sweep(DevProblem p, int K, unsigned long long seed,
                                                          long long chainOffset, unsigned it, double* __restrict__ x,
                                                          const double* __restrict__ sd, int* __restrict__ moves,
                                                          const double* __restrict__ tab) {
    extern __shared__ __align__(16) double lsm[];  // [log table (4 KB) | table row of this chain]
    double2* sLogT = reinterpret_cast<double2*>(lsm);
    double* srow = lsm + 2 * BAM_LOGTAB_N;
    const int c = blockIdx.y;
    for (int q = threadIdx.x; q < BAM_LOGTAB_N; q += BAM_BLOCK) sLogT[q] = p.logTab[q];
    for (int q = threadIdx.x; q < p.tabStride; q += BAM_BLOCK) srow[q] = tab[(size_t)c * p.tabStride + q];
    __syncthreads();
    const unsigned lt = smem_addr(sLogT);
    const int n = p.n;
    const int i = blockIdx.x * BAM_BLOCK + threadIdx.x;
    if (i >= n) return;
    double xo[NX], xu[NX], xb[NX], yo[NY], yu[NY], yb[NY], xt[NX];
    int xbi[NX], ybi[NY], lat[NX];
    bool any = false;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const size_t q = i + (size_t)j * n;
        xo[j] = p.X[q]; xu[j] = p.Xu[q]; lat[j] = p.latIdx[q];
        xb[j] = 0.0; xbi[j] = 0;
        if (p.hasXbias) { xb[j] = p.Xb[q]; xbi[j] = p.Xbindx[q]; }
        any |= lat[j] >= 0;
    }
    if (!any) return;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        const size_t q = i + (size_t)j * n;
        yo[j] = p.Y[q]; yu[j] = p.Yu[q];
        yb[j] = 0.0; ybi[j] = 0;
        if (p.hasYbias) { yb[j] = p.Yb[q]; ybi[j] = p.Ybindx[q]; }
    }
    const double* __restrict__ row = srow;
    const int combo = (p.nCombo > 1) ? p.comboOf[i] : 0;
    const double* __restrict__ th = row + p.tabCombo + (size_t)combo * p.comboStride;
    const size_t base = (size_t)c * p.P;
#pragma unroll
    for (int j = 0; j < NX; ++j) {  // UnfoldParVector (:3429-3452)
        double v = xo[j];
        if (lat[j] >= 0) v = x[base + lat[j]];
        else if (p.hasXbias && xbi[j] > 0) v = xo[j] - xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
        xt[j] = v;
    }
    double hyC[NX], hyR[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        hyC[j] = 0.0; hyR[j] = 0.0;
        if (xu[j] > 0.0) { hyC[j] = -BAM_LOG_SQRT_2PI - log(xu[j]); hyR[j] = 1.0 / xu[j]; }
    }
    double cur;
    if (!local_logpost<MODEL, NX, NY>(p, xt, xo, xu, xb, xbi, yo, yu, yb, ybi, hyC, hyR, row, th, lt, cur)) return;  // cannot happen for a feasible state
    const unsigned long long gc = (unsigned long long)(chainOffset + c);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        if (lat[j] < 0) continue;
        const unsigned comp = (unsigned)lat[j];
        const double old = xt[j];
        xt[j] = old + sd[base + comp] * philox_normal(seed, gc, it, 2u * comp);
        double cand;
        bool acc = false;
        if (local_logpost<MODEL, NX, NY>(p, xt, xo, xu, xb, xbi, yo, yu, yb, ybi, hyC, hyR, row, th, lt, cand)) {
            const double u = philox_uniform(seed, gc, it, 2u * comp + 1u);
            acc = log(u) < cand - cur;
        }
        if (acc) {
            cur = cand;
            x[base + comp] = xt[j];
            moves[base + comp] += 1;
        } else xt[j] = old;
    }
}

using LatentKernel = void (*)(DevProblem, int, unsigned long long, long long, unsigned, double*, const double*, int*,
                              const double*);
LatentKernel pick_latent_sweep(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return latent_sweep<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

using WarpSamplerKernel = void (*)(DevProblem, WarpSamplerArgs);
WarpSamplerKernel pick_warp_sampler(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return sampler_warp<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

}  // namespace

namespace bam {

void launch_fit(bamgpu_handle* h, int K, int nAdapt, int nCycles, double minMR, double maxMR, double downMult,
                double upMult, unsigned long long seed, int burn, int keepEvery, bool storeRows, long long nKeep) {
    const DevProblem& p = h->dp;
    const int P = p.P, nD = p.nDpar;
    cudaStream_t s = h->stream;
    const int TB = 128, gK = (K + TB - 1) / TB;
    const long long KP = (long long)K * P, Krow = (long long)K * (P + 1 + nD);

    BAM_CUDA(cudaMemsetAsync(h->d_moves, 0, sizeof(int) * KP, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mcount, 0, sizeof(double) * K, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mmean, 0, sizeof(double) * KP, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mm2, 0, sizeof(double) * KP, s));

    // f(start): must be feasible (the reference aborts otherwise)
    launch_logpost(h, K, h->d_x, -1, nullptr, false);
    int* d_nbad = nullptr;
    BAM_CUDA(cudaMalloc(&d_nbad, sizeof(int)));
    BAM_CUDA(cudaMemsetAsync(d_nbad, 0, sizeof(int), s));
    init_state_kernel<<<gK, TB, 0, s>>>(K, nD, h->d_fx, h->d_dpar, h->d_status, h->d_fxcur, h->d_dparcur, d_nbad);
    int nbad = 0;
    BAM_CUDA(cudaMemcpyAsync(&nbad, d_nbad, sizeof(int), cudaMemcpyDeviceToHost, s));
    BAM_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_nbad);
    h->launches += 1;
    if (nbad > 0) throw CudaError{"Adaptive_Metro_OAAT: unfeasible starting point for " + std::to_string(nbad) + " chain(s)"};

    // small data sets: the persistent warp-per-chain sampler (one launch per cycle)
    WarpSamplerKernel wk = pick_warp_sampler(p.model, p.nX, p.nY);
    const size_t wsSmem = sizeof(double) * WS_WARPS * (size_t)(2 * P + p.tabStride + p.kM + 2);
    if (wk && !h->forceGeneric && !p.hasLatent && P <= WS_MAXP && p.n <= 4096 && wsSmem <= 200 * 1024) {
        if (wsSmem > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)wk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsSmem));
        WarpSamplerArgs a;
        a.K = K; a.nAdapt = nAdapt; a.burn = burn; a.keepEvery = keepEvery; a.nKeep = nKeep;
        a.minMR = minMR; a.maxMR = maxMR; a.downMult = downMult; a.upMult = upMult; a.seed = seed;
        a.chainOffset = h->chainOffset;
        a.x = h->d_x; a.sd = h->d_std; a.fxcur = h->d_fxcur; a.dparcur = h->d_dparcur;
        a.rows = storeRows ? h->d_rows : nullptr; a.mcount = h->d_mcount; a.mmean = h->d_mmean; a.mm2 = h->d_mm2;
        a.moves = h->d_moves;
        cudaEvent_t f0 = take_event(h), f1 = take_event(h);  // the iterations as one timed "step" (bamgpu_step_times)
        BAM_CUDA(cudaEventRecord(f0, s));
        for (int cyc = 0; cyc < nCycles; ++cyc) {
            a.it0 = (long long)cyc * nAdapt;
            a.nIterChunk = nAdapt;
            wk<<<(K + WS_WARPS - 1) / WS_WARPS, 32 * WS_WARPS, wsSmem, s>>>(p, a);
            h->launches += 1;
        }
        BAM_CUDA(cudaEventRecord(f1, s));
        if (h->stepEv.size() < 8192) h->stepEv.emplace_back(f0, f1);
        BAM_CUDA(cudaGetLastError());
        BAM_CUDA(cudaMemcpyAsync(h->d_fx, h->d_fxcur, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
        BAM_CUDA(cudaMemcpyAsync(h->d_dpar, h->d_dparcur, sizeof(double) * (size_t)K * std::max(nD, 1), cudaMemcpyDeviceToDevice, s));
        BAM_CUDA(cudaMemsetAsync(h->d_status, 0, sizeof(int) * K, s));
        h->haveMoments = true;
        return;
    }

    // components visited with one launch set per proposal: everything except the latent inputs when the one-launch
    // sweep applies (vector order: theta | gamma | latent inputs | X biases | Y biases, Packer :4202-4242)
    LatentKernel lk = (p.hasLatent && !h->forceGeneric) ? pick_latent_sweep(p.model, p.nX, p.nY) : nullptr;
    const int latBeg = p.nFit + p.nGamma, latEnd = latBeg + p.nLatent;
    // fast path (BaRatin, many chains): per-combination likelihood sums of the current state are kept, so that a
    // proposal on a VAR parameter only re-evaluates the observations of the combinations it touches
    const bool comboSums = fast_path_keeps_combo_sums(h, K) && !lk && !h->noPartial;
    if (comboSums) {  // the full evaluation of the start point above left its sums in d_lkCcand
        BAM_CUDA(cudaMemcpyAsync(h->d_lkCcur, h->d_lkCcand, sizeof(double) * (size_t)K * p.nCombo, cudaMemcpyDeviceToDevice, s));
        h->lkCvalid = true;
    }
    auto propose_one = [&](int i, long long it) {
        propose_kernel<<<gK, TB, 0, s>>>(K, P, i, seed, h->chainOffset, (unsigned)it, h->d_std, h->d_delta);
        launch_logpost(h, K, h->d_x, i, h->d_delta, false, comboSums);
        accept_kernel<<<gK, TB, 0, s>>>(K, P, nD, i, seed, h->chainOffset, (unsigned)it, h->d_delta, h->d_fx,
                                        h->d_status, h->d_dpar, h->d_x, h->d_fxcur, h->d_dparcur, h->d_moves, p.nCombo,
                                        comboSums ? h->d_lkCcand : nullptr, comboSums ? h->d_lkCcur : nullptr);
        h->launches += 2;
    };
    long long it = 0, kept = 0;
    cudaEvent_t f0 = take_event(h), f1 = take_event(h);  // the iterations as one timed "step" (bamgpu_step_times)
    BAM_CUDA(cudaEventRecord(f0, s));
    for (int cyc = 0; cyc < nCycles; ++cyc) {
        for (int a = 0; a < nAdapt; ++a, ++it) {
            for (int i = 0; i < (lk ? latBeg : P); ++i) propose_one(i, it);
            if (lk) {
                // the table of the CURRENT state (theta, gamma, biases of every chain), then the sweep, then the
                // full log-posterior of the swept state becomes the current one
                launch_table(h, K, h->d_x);
                dim3 grid((p.n + BAM_BLOCK - 1) / BAM_BLOCK, K);
                lk<<<grid, BAM_BLOCK, sizeof(double) * (p.tabStride + 2 * BAM_LOGTAB_N), s>>>(p, K, seed, h->chainOffset, (unsigned)it, h->d_x,
                                                                        h->d_std, h->d_moves, h->d_tab);
                launch_logpost(h, K, h->d_x, -1, nullptr, false);
                BAM_CUDA(cudaMemcpyAsync(h->d_fxcur, h->d_fx, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
                h->launches += 1;
                for (int i = latEnd; i < P; ++i) propose_one(i, it);
            }
            const long long it1 = it + 1;
            if (it1 > burn && (it1 - burn) % keepEvery == 0 && kept < nKeep) {
                record_kernel<<<(unsigned)((Krow + 255) / 256), 256, 0, s>>>(K, P, nD, kept, nKeep, h->d_x, h->d_fxcur,
                                                                             h->d_dparcur, storeRows ? h->d_rows : nullptr,
                                                                             h->d_mcount, h->d_mmean, h->d_mm2);
                ++kept;
                h->launches += 1;
            }
        }
        adapt_kernel<<<(unsigned)((KP + 255) / 256), 256, 0, s>>>(K, P, nAdapt, minMR, maxMR, downMult, upMult, h->d_std,
                                                                  h->d_moves);
        h->launches += 1;
        BAM_CUDA(cudaGetLastError());
    }
    BAM_CUDA(cudaEventRecord(f1, s));
    if (h->stepEv.size() < 8192) h->stepEv.emplace_back(f0, f1);
    h->lkCvalid = false;
    // leave the final state consistent for bamgpu_chains_download
    BAM_CUDA(cudaMemcpyAsync(h->d_fx, h->d_fxcur, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    BAM_CUDA(cudaMemcpyAsync(h->d_dpar, h->d_dparcur, sizeof(double) * (size_t)K * std::max(nD, 1), cudaMemcpyDeviceToDevice, s));
    BAM_CUDA(cudaMemsetAsync(h->d_status, 0, sizeof(int) * K, s));
    h->haveMoments = true;
}

}  // namespace bam
