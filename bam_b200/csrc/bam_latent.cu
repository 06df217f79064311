// bam_latent.cu -- K3l: the one-launch sweep over all latent "true input" components (see bam_sampler.cu).
#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_sampler_kernels.h"

namespace {

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

// ------------------------------------------------------------------------------------------------
// K3l fast path (cfg 3): Linear model, every input cell latent, no biases, no missing Y, one sigma function -- the
// eligibility of logpost_linear_latent.  Same sweep, same Philox streams, same accept rule; what changes is the cost:
//   * every load of the thread (state, jump scales, observation tables) is issued before the first use;
//   * the posterior of one observation is  const - T/2,  T = sum_j [log v_j + r_j^2 / v_j] + sum_k z_k^2  (the constants
//     cancel in candidate - current); NY == 2 shares one log and one reciprocal through v_0 v_1;
//   * table-driven log / sqrt / cos for the Box-Muller normal and the log of the acceptance uniform (libdevice where
//     the table log loses relative accuracy, u within 2^-10 of 1).
// ------------------------------------------------------------------------------------------------
template <int NX, int NY, int SIGF>
__device__ __forceinline__ bool linear_T(const double* xt, const double* xo, const double* hr, const double* yo,
                                         const double* yu2, const double* __restrict__ th, const double* __restrict__ gm,
                                         unsigned lt, double& T) {
    double v[NY], r2[NY];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        double yh = 0.0;
#pragma unroll
        for (int k = 0; k < NX; ++k) yh = yh + xt[k] * th[j * NX + k];
        double sig;
        if (SIGF == BAMGPU_SIG_CONSTANT) sig = gm[j];
        else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(gm[2 * j + 1], fabs(yh), gm[2 * j]);
        else sig = gm[j] * fabs(yh);
        ok &= sig > 0.0;
        v[j] = fma(sig, sig, yu2[j]);
        ok &= ((unsigned)__double2hiint(v[j]) - 0x00100000u < 0x7fe00000u);
        const double r = yo[j] - yh;
        r2[j] = r * r;
    }
    double t;
    if (NY == 2) {
        const double vp = v[0] * v[1];
        if ((unsigned)__double2hiint(vp) - 0x00100000u < 0x7fe00000u)
            t = fma(fma(r2[0], v[1], r2[1] * v[0]), trcp(vp), tlog_pos(vp, lt));
        else t = fma(r2[0], trcp(v[0]), tlog_pos(v[0], lt)) + fma(r2[1], trcp(v[1]), tlog_pos(v[1], lt));
    } else {
        t = 0.0;
#pragma unroll
        for (int j = 0; j < NY; ++j) t += fma(r2[j], trcp(v[j]), tlog_pos(v[j], lt));
    }
#pragma unroll
    for (int k = 0; k < NX; ++k) {
        const double z = (xo[k] - xt[k]) * hr[k];
        t = fma(z, z, t);
    }
    T = t;
    return ok && t == t;
}

#ifndef LS_MINB
#define LS_MINB 4
#endif
template <int NX, int NY, int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK, LS_MINB) latent_sweep_linear(DevProblem p, int K, unsigned long long seed,
                                                                 long long chainOffset, unsigned it, double* __restrict__ x,
                                                                 const double* __restrict__ sd, int* __restrict__ moves,
                                                                 const double* __restrict__ tab) {
    constexpr int NG = (SIGF == BAMGPU_SIG_LINEAR) ? 2 : 1;
    extern __shared__ __align__(16) double lsm[];  // [log table (4 KB) | theta | gamma]
    double2* sLogT = reinterpret_cast<double2*>(lsm);
    double* sth = lsm + 2 * BAM_LOGTAB_N;
    double* sgm = sth + NX * NY;
    const int c = blockIdx.y;
    for (int q = threadIdx.x; q < BAM_LOGTAB_N; q += BAM_BLOCK) sLogT[q] = p.logTab[q];
    if (threadIdx.x < NX * NY) sth[threadIdx.x] = tab[(size_t)c * p.tabStride + p.tabCombo + threadIdx.x];
    if (threadIdx.x < NG * NY)
        sgm[threadIdx.x] = tab[(size_t)c * p.tabStride + p.tabGamma + p.gamOff[threadIdx.x / NG] + threadIdx.x % NG];
    __syncthreads();
    const unsigned lt = smem_addr(sLogT);
    const int n = p.n;
    const int i = blockIdx.x * BAM_BLOCK + threadIdx.x;
    if (i >= n) return;
    const size_t base = (size_t)c * p.P + (size_t)p.nFit + p.nGamma;  // cell (i, j) at base + j n + i
    double xt[NX], sdv[NX], xo[NX], hr[NX], yo[NY], yu2[NY], th[NX * NY], gm[NG * NY];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        xt[j] = x[base + (size_t)j * n + i];
        sdv[j] = sd[base + (size_t)j * n + i];
        xo[j] = p.X[i + (size_t)j * n];
        hr[j] = p.XuInv[i + (size_t)j * n];
    }
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        yo[j] = p.Y[i + (size_t)j * n];
        const double yu = p.Yu[i + (size_t)j * n];
        yu2[j] = yu * yu;
    }
#pragma unroll
    for (int q = 0; q < NX * NY; ++q) th[q] = sth[q];
#pragma unroll
    for (int q = 0; q < NG * NY; ++q) gm[q] = sgm[q];
    double cur;
    if (!linear_T<NX, NY, SIGF>(xt, xo, hr, yo, yu2, th, gm, lt, cur)) return;  // cannot happen for a feasible state
    const unsigned long long gc = (unsigned long long)(chainOffset + c);
    const unsigned compBase = (unsigned)(p.nFit + p.nGamma);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const unsigned comp = compBase + (unsigned)j * (unsigned)n + (unsigned)i;
        const double zn = philox_normal_fast(seed, gc, it, 2u * comp, lt);
        const double old = xt[j];
        xt[j] = fma(sdv[j], zn, old);
        double cand;
        bool acc = false;
        if (linear_T<NX, NY, SIGF>(xt, xo, hr, yo, yu2, th, gm, lt, cand)) {
            acc = philox_log_uniform_fast(seed, gc, it, 2u * comp + 1u, lt) < -0.5 * (cand - cur);
        }
        if (acc) {
            cur = cand;
            x[base + (size_t)j * n + i] = xt[j];
            moves[base + (size_t)j * n + i] += 1;
        } else xt[j] = old;
    }
}

}  // namespace

namespace bam {
template <int NX, int NY>
static LatentKernel pick_sweep_linear_sig(int sigf) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return latent_sweep_linear<NX, NY, BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return latent_sweep_linear<NX, NY, BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return latent_sweep_linear<NX, NY, BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}
// same eligibility as the log-posterior fast path of cfg 3 (bam_logpost.cu: pick_latent)
LatentKernel pick_latent_sweep_linear(const bamgpu_handle* h) {
    const DevProblem& p = h->dp;
    const HostLayout& L = h->L;
    if (h->forceGeneric || p.model != BAMGPU_MDL_LINEAR || !p.hasLatent || p.n < 1 || !p.XuInv) return nullptr;
    if ((long long)L.nLatent != (long long)p.n * p.nX || p.n != L.n) return nullptr;
    if (p.hasXbias || p.hasYbias || p.hasMissing || p.hyperInfeasible || p.nCombo != 1) return nullptr;
    for (int j = 1; j < p.nY; ++j)
        if (p.sigfunc[j] != p.sigfunc[0]) return nullptr;
    if (p.nX == 1 && p.nY == 1) return pick_sweep_linear_sig<1, 1>(p.sigfunc[0]);
    if (p.nX == 2 && p.nY == 1) return pick_sweep_linear_sig<2, 1>(p.sigfunc[0]);
    if (p.nX == 1 && p.nY == 2) return pick_sweep_linear_sig<1, 2>(p.sigfunc[0]);
    if (p.nX == 2 && p.nY == 2) return pick_sweep_linear_sig<2, 2>(p.sigfunc[0]);
    return nullptr;
}

LatentKernel pick_latent_sweep(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return latent_sweep<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}
}  // namespace bam
