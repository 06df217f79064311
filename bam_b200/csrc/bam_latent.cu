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

}  // namespace

namespace bam {
LatentKernel pick_latent_sweep(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return latent_sweep<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}
}  // namespace bam
