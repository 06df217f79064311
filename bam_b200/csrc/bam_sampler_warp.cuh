// bam_sampler_warp.cuh -- K3p: the persistent warp-per-chain sampler (see bam_sampler.cu for the sampler semantics).
// Included by bam_sampler_warp_{a,b,c}.cu, each of which instantiates it for one group of (model, nX, nY).
#pragma once
#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_baratin_fast.cuh"
#include "bam_sampler_kernels.h"

namespace {

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
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// log-posterior of the vector xs (shared memory) for one chain, computed by one warp.  Returns status bits.
template <int MODEL, int NX, int NY>
__device__ int warp_logpost(const DevProblem& p, const double* xs, double* tab, double* vcop, double& fx, unsigned lt,
                            unsigned et) {
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
        const bool okp = (MODEL == BAMGPU_MDL_BARATIN) ? baratin_prepare_t(p, th, der, lt, et) : model_prepare(p, th, der, xc, p.n);
        if (!okp) st |= ST_LKH_INFEAS;
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
            const bool oka = (MODEL == BAMGPU_MDL_BARATIN) ? baratin_apply_t(p, xt[0], th, th + p.ntheta, yh[0], lt, et)
                                                           : model_apply<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh);
            if (!oka) { bad = true; continue; }
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
                lk += gauss_logpdf_var_t(yo, mu, fma(yu, yu, sig * sig), lt);
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
    extern __shared__ __align__(16) double wsm[];  // [log table | exp table | per warp: xs | sd | tab | copula v]
    double2* sLogT = reinterpret_cast<double2*>(wsm);
    double* sExpT = wsm + 2 * BAM_LOGTAB_N;
    for (int q = threadIdx.x; q < BAM_LOGTAB_N; q += 32 * WS_WARPS) sLogT[q] = p.logTab[q];
    for (int q = threadIdx.x; q < BAM_EXPTAB_N; q += 32 * WS_WARPS) sExpT[q] = p.expTab[q];
    __syncthreads();
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * WS_WARPS + w;
    if (c >= a.K) return;
    const int P = p.P, nD = p.nDpar;
    const int perWarp = 2 * P + p.tabStride + p.kM + 2;  // xs | sd | tab | copula v
    double* xs = wsm + WS_TABLES + (size_t)w * perWarp;
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
            const double delta = sd[i] * philox_normal_fast(a.seed, gc, (unsigned)it, 2u * (unsigned)i, lt);
            __syncwarp();
            if (lane == 0) xs[i] = old + delta;
            __syncwarp();
            double fc;
            const int st = warp_logpost<MODEL, NX, NY>(p, xs, tab, vcop, fc, lt, et);
            bool acc = false;
            if (st == 0) acc = philox_log_uniform_fast(a.seed, gc, (unsigned)it, 2u * (unsigned)i + 1u, lt) < fc - fxc;
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

}  // namespace

