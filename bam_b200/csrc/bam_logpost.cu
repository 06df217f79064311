// bam_logpost.cu -- batched log-posterior for K parameter vectors x all observations.
//
// Replaces Posterior_wrapper -> UnfoldParVector -> GetPostLogPdf -> {GetPriorLogPdf, GetLogLikelihood
// (ApplyModel_BaM + Sigmafunk_Apply + GetPdf per term), GetHyperLogPdf} (src/BaM_tools.f90:3006-3517,
// 4115-4177) with three kernels:
//
//   K2 logpost_prep     one thread per vector: priors (+ Gaussian copula, bias priors) in the reference's
//                       sequential order with its early exits; expands theta for every VAR combination
//                       (FIX filled in) and runs the per-parameter-set model pre-pass (BaRatin continuity)
//                       ONCE per combination instead of once per observation; writes the per-chain table.
//   K1 logpost_main     grid = (observation tiles) x (chain tiles).  A thread owns one observation: its
//                       X/Y/Yu/... live in registers and are re-used for every chain of the tile, so the
//                       observation tables are read from HBM once per chain tile (24-28 B per observation
//                       shared by CT chains).  The chain tile's tables are staged in shared memory; all
//                       lanes of a warp read the same entry (broadcast).  Model + sigma function +
//                       Gaussian log-pdf (+ hyper term of latent inputs) are fused; a warp-shuffle tree
//                       and a fixed-order cross-warp sum give one partial per (tile, chain):
//                       deterministic, no atomics on doubles.
//   K1b logpost_final   one thread per vector: sums the tile partials in fixed order, applies the
//                       early-exit priority prior -> likelihood -> hyper, writes fx / flags.
#include <algorithm>
#include <cstdio>

#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_baratin_fast.cuh"

#ifndef BAM_CB_GEN
#define BAM_CB_GEN 2  /* generation of the BaRatin fast path: 1 = logpost_baratin_cb, 2 = logpost_baratin_cb2 (persistent) */
#endif

namespace {
#include "bam_baratin_cb2.cuh"
#include "bam_sediment_fast.cuh"

// value q of vector c, with the optional single-component candidate shift
__device__ __forceinline__ double xval(const double* __restrict__ x, size_t base, int q, int comp, double dl) {
    double v = x[base + q];
    return (q == comp) ? v + dl : v;
}

// ------------------------------------------------------------------------------------------------
// K2: prior + per-chain table.  ONE WARP PER VECTOR: lanes stride over the parameters (priors, copula scores), over the
// copula's columns, over the table entries and over the VAR combinations (one model pre-pass per lane).  With one thread
// per vector this kernel was a serial chain of ~4500 instructions (19 log-densities, 19 quantiles + a 19 x 19 product,
// five continuity pre-passes with two pow each for cfg 2): 55 us of a 1.5 ms step at K = 4096 and 80 us of a 0.11 ms
// step for a single chain.
// The reference evaluates the priors in sequence and leaves at the first parameter that is infeasible or has zero
// density (GetPriorLogPdf :3062-3112: remnant parameters, then theta, then the copula); the flag of a vector is
// therefore that of its FIRST failing parameter in that order, found here with a warp minimum over (order, kind).
// ------------------------------------------------------------------------------------------------
// SPEC = a model id compiles the pre-pass of that model only (the generic instance carries the pre-passes of every model:
// 13 000 instructions, and at 512 vectors per GPU the kernel was 25 us of instruction-cache misses), 0 = any model.
#define PREP_WARPS 4
template <int SPEC>
__global__ void __launch_bounds__(32 * PREP_WARPS) logpost_prep(DevProblem p, const double* __restrict__ x, int K, int comp,
                                                                const double* __restrict__ delta, double* __restrict__ tab,
                                                                double* __restrict__ prior_out, int* __restrict__ status,
                                                                double* __restrict__ dpar) {
    extern __shared__ double psm[];  // PREP_WARPS x kM copula scores
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * PREP_WARPS + w;
    if (c >= K) return;
    double* vcop = psm + (size_t)w * max(p.kM, 1);
    const size_t base = (size_t)c * p.P;
    const double dl = (comp >= 0) ? delta[c] : 0.0;
    const int npar = p.nFit + p.nGamma;
    // ---- priors: order index of a parameter = its place in the reference's sequence (gamma first, then theta)
    double pr = 0.0;
    unsigned fail = 0xffffffffu;  // (order << 1) | kind, kind 0 = infeasible, 1 = null; minimum over the warp = first failure
    for (int q = lane; q < npar; q += 32) {
        double lp;
        bool feas, isnull;
        const int order = q;
        // one call site (the distribution switch is ~1000 instructions): remnant parameters first, then theta
        const bool isG = q < p.nGamma;
        const int k = isG ? q : q - p.nGamma;
        dist_logpdf(isG ? p.pgDist[k] : p.ptDist[k], (isG ? p.pgPar : p.ptPar) + (size_t)k * BAMGPU_PRIOR_NPAR,
                    xval(x, base, isG ? p.nFit + k : k, comp, dl), lp, feas, isnull);
        if (!feas) fail = min(fail, ((unsigned)order << 1));
        else if (isnull) fail = min(fail, ((unsigned)order << 1) | 1u);
        else pr += lp;
    }
    fail = __reduce_min_sync(0xffffffffu, fail);
    if (fail == 0xffffffffu && p.priorM != nullptr) {  // Gaussian copula (:3080-3112), v = qnorm(cdf(x)) (Gcop_getV :4246)
        const int k = p.kM;
        unsigned cfail = 0xffffffffu;
        for (int i = lane; i < k; i += 32) {
            double u;
            bool f;
            const bool isT = i < p.nFit;
            const int kk = isT ? i : i - p.nFit;
            dist_cdf(isT ? p.ptDist[kk] : p.pgDist[kk], (isT ? p.ptPar : p.pgPar) + (size_t)kk * BAMGPU_PRIOR_NPAR, xval(x, base, i, comp, dl), u, f);
            if (!f || !(u > 0.0 && u < 1.0)) { cfail = 0u; vcop[i] = 0.0; }
            else vcop[i] = normcdfinv(u);
        }
        cfail = __reduce_min_sync(0xffffffffu, cfail);
        __syncwarp();
        if (cfail == 0u) fail = ((unsigned)npar << 1);  // after every prior: infeasible
        else {
            double qf = 0.0;
            for (int cc = lane; cc < k; cc += 32) {
                double wv = 0.0;
                for (int r = 0; r < k; ++r) wv = wv + vcop[r] * p.priorM[r + (size_t)cc * k];
                qf += wv * vcop[cc];
            }
            pr -= 0.5 * qf;
        }
    }
    // N(0,1) priors of the input / output biases (:3116-3135): never fail
    for (int j = 0; j < p.nX; ++j)
        for (int i = lane; i < p.nXb[j]; i += 32) pr += gauss_logpdf(xval(x, base, p.offXb[j] + i, comp, dl), 0.0, 1.0);
    for (int j = 0; j < p.nY; ++j)
        for (int i = lane; i < p.nYb[j]; i += 32) pr += gauss_logpdf(xval(x, base, p.offYb[j] + i, comp, dl), 0.0, 1.0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, off);
    int st = 0;
    if (fail != 0xffffffffu) st = (fail & 1u) ? ST_PRIOR_NULL : ST_PRIOR_INFEAS;

    // ---- per-chain table
    double* row = tab + (size_t)c * p.tabStride;
    for (int g = lane; g < p.nGamma; g += 32) row[p.tabGamma + g] = xval(x, base, p.nFit + g, comp, dl);
    for (int j = 0; j < p.nX; ++j)
        for (int i = lane; i < p.nXb[j]; i += 32) row[p.tabOffXb[j] + i] = xval(x, base, p.offXb[j] + i, comp, dl);
    for (int j = 0; j < p.nY; ++j)
        for (int i = lane; i < p.nYb[j]; i += 32) row[p.tabOffYb[j] + i] = xval(x, base, p.offYb[j] + i, comp, dl);
    int stl = 0;
    if (SPEC != 0) {  // the generic instance leaves the model pre-pass to logpost_prep_model (thread per vector x combination)
        for (int k = lane; k < p.nCombo; k += 32) {  // ApplyModel_BaM theta expansion (:4133-4159), one combination per lane
            constexpr int MT = (SPEC == BAMGPU_MDL_BARATIN) ? 12 : BAM_MAXTHETA, MD = (SPEC == BAMGPU_MDL_BARATIN) ? 8 : BAM_MAXDER;
            double th[MT], der[MD];
            double* cr = row + p.tabCombo + (size_t)k * p.comboStride;
            for (int i = 0; i < p.ntheta; ++i) {
                const int src = p.comboSrc[(size_t)k * p.ntheta + i];
                th[i] = (src < 0) ? p.fixv[i] : xval(x, base, src, comp, dl);
                cr[i] = th[i];
            }
            for (int d = 0; d < p.nDer; ++d) der[d] = p.unfeas;
            const bool ok = baratin_prepare(p, th, der);
            if (!ok) stl |= ST_LKH_INFEAS;
            for (int d = 0; d < p.nDer; ++d) cr[p.ntheta + d] = der[d];
            if (dpar != nullptr)
                for (int d = 0; d < p.nDparModel; ++d) dpar[(size_t)c * p.nDpar + k * p.nDparModel + d] = ok ? der[d] : p.unfeas;
        }
    }
    st |= __reduce_or_sync(0xffffffffu, stl);
    if (lane == 0) {
        prior_out[c] = pr;
        status[c] = st;
    }
}

// K2m: model pre-pass of the generic path, THREAD per (vector, VAR combination), consecutive lanes = consecutive vectors
// of one combination.  Inside logpost_prep (one warp per vector, lanes over the combinations) a model with a single
// combination ran its pre-pass on one lane of the warp: the per-year degree-day scan of Vegetation (2 x 1095 rows) was
// 72 000 warp-instructions per vector, 0.80 ms of a 1.2 ms step at 4096 vectors.  Runs after logpost_prep on the same
// stream: ORs its flag into the status word written there.
#define PREPM_BLOCK 64
__global__ void __launch_bounds__(PREPM_BLOCK) logpost_prep_model(DevProblem p, const double* __restrict__ x, int K, int comp,
                                                                  const double* __restrict__ delta, double* __restrict__ tab,
                                                                  int* __restrict__ status, double* __restrict__ dpar, int stageSeries) {
    extern __shared__ double pms[];  // stageSeries: the three series the degree-day scan of Vegetation walks
    const long long t = (long long)blockIdx.x * PREPM_BLOCK + threadIdx.x;
    const double* xc[BAM_MAXX];
    for (int j = 0; j < p.nX; ++j) xc[j] = p.X + (size_t)j * p.n;
    if (stageSeries) {
        // every thread of the block walks the SAME series four rows (one 32-byte sector) per trip: from global memory each
        // trip waited for an L2 round trip (~300 cycles x 274 trips x 2 sweeps = most of the kernel's 0.22 ms)
        for (int i = threadIdx.x; i < p.n; i += PREPM_BLOCK) {
            pms[i] = xc[0][i];
            pms[p.n + i] = xc[2][i];
            pms[2 * (size_t)p.n + i] = xc[3][i];
        }
        __syncthreads();
        xc[0] = pms; xc[2] = pms + p.n; xc[3] = pms + 2 * (size_t)p.n;
    }
    if (t >= (long long)K * p.nCombo) return;
    const int c = (int)(t % K), k = (int)(t / K);
    const size_t base = (size_t)c * p.P;
    const double dl = (comp >= 0) ? delta[c] : 0.0;
    double th[BAM_MAXTHETA], der[BAM_MAXDER];
    double* cr = tab + (size_t)c * p.tabStride + p.tabCombo + (size_t)k * p.comboStride;
    for (int i = 0; i < p.ntheta; ++i) {  // ApplyModel_BaM theta expansion (:4133-4159)
        const int src = p.comboSrc[(size_t)k * p.ntheta + i];
        th[i] = (src < 0) ? p.fixv[i] : xval(x, base, src, comp, dl);
        cr[i] = th[i];
    }
    for (int d = 0; d < p.nDer; ++d) der[d] = p.unfeas;
    const bool ok = model_prepare(p, th, der, xc, p.n);
    if (p.isSeries) der[0] = (double)c;  // this vector's column of the series buffer (model_apply looks it up)
    if (!ok) atomicOr(&status[c], ST_LKH_INFEAS);
    for (int d = 0; d < p.nDer; ++d) cr[p.ntheta + d] = der[d];
    if (dpar != nullptr)
        for (int d = 0; d < p.nDparModel; ++d) dpar[(size_t)c * p.nDpar + k * p.nDparModel + d] = ok ? der[d] : p.unfeas;
}

using PrepKernel = void (*)(DevProblem, const double*, int, int, const double*, double*, double*, int*, double*);
static PrepKernel pick_prep(const DevProblem& p) {
    if (p.model == BAMGPU_MDL_BARATIN && p.ncontrol <= 4 && p.ntheta <= 12 && p.nDer <= 8) return logpost_prep<BAMGPU_MDL_BARATIN>;
    return logpost_prep<0>;
}
// K2 (+ K2m for the generic instance); returns the number of launches
static int launch_prep_on(cudaStream_t stream, const DevProblem& p, int K, const double* d_x, int comp, const double* d_delta,
                          double* tab, double* prior, int* status, double* dpar) {
    PrepKernel k = pick_prep(p);
    k<<<(K + PREP_WARPS - 1) / PREP_WARPS, 32 * PREP_WARPS, sizeof(double) * PREP_WARPS * std::max(p.kM, 1), stream>>>(
        p, d_x, K, comp, d_delta, tab, prior, status, dpar);
    if (k != logpost_prep<0>) return 1;
    const long long nt = (long long)K * p.nCombo;
    size_t smem = 0;
    if (p.model == BAMGPU_MDL_VEGETATION && p.vegGrowth != BAMGPU_VEG_YIN1_TEMPORAL && p.nX >= 4 && p.n > 0) {
        smem = sizeof(double) * 3 * (size_t)p.n;
        if (smem > 200 * 1024) smem = 0;
        else if (smem > 48 * 1024)  // (per device and cheap: set on every launch that needs it -- bam_gpu --gpus N drives several devices)
            cudaFuncSetAttribute((const void*)logpost_prep_model, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    logpost_prep_model<<<(unsigned)((nt + PREPM_BLOCK - 1) / PREPM_BLOCK), PREPM_BLOCK, smem, stream>>>(p, d_x, K, comp, d_delta, tab, status, dpar,
                                                                                                     smem ? 1 : 0);
    return 2;
}
static int launch_prep(bamgpu_handle* h, int K, const double* d_x, int comp, const double* d_delta) {
    return launch_prep_on(h->stream, h->dp, K, d_x, comp, d_delta, h->d_tab, h->d_prior, h->d_status, h->d_dpar);
}

// ------------------------------------------------------------------------------------------------
// K1s: time-recursive models (row f4).  thread <-> parameter vector: one sequential pass over the series, outputs (and
// state variables when nOut > nY) written to seriesY[(y * n + t) * K + vector] -- consecutive lanes = consecutive
// vectors, so every step of the recursion is one coalesced store.  An infeasible vector raises ST_LKH_INFEAS.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) series_run(DevProblem p, int K, const double* __restrict__ tab, int stride, int thOff,
                                                  int* __restrict__ status, int nOut) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (status[c] != 0) return;
    if (!model_series(p, tab + (size_t)c * stride + thOff, p.seriesY + c, (size_t)K, nOut)) atomicOr(&status[c], ST_LKH_INFEAS);
}

// ------------------------------------------------------------------------------------------------
// K1: fused model + likelihood (+ hyper)
// ------------------------------------------------------------------------------------------------
template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(BAM_BLOCK) logpost_main(DevProblem p, const double* __restrict__ x, int K, int CT, int TX,
                                                          const double* __restrict__ tab, const int* status,
                                                          double* __restrict__ part, int* statusOut, int comp,
                                                          const double* __restrict__ delta, int smemLogOff) {
    extern __shared__ double sm[];
    const int tabStride = p.tabStride;
    double* stab = sm;                        // CT x tabStride
    const int warpsPerRow = TX >> 5;
    double* red = sm + (size_t)CT * tabStride;  // [2][warpsPerRow][CT]
    int* sst = (int*)(red + 2 * warpsPerRow * CT);  // CT chain status
    double2* sLogT = reinterpret_cast<double2*>(sm + smemLogOff);  // table of the table-driven log (16-byte aligned)
    double* sExpT = sm + smemLogOff + 2 * BAM_LOGTAB_N;            // and of the table-driven exp (BaRatin)

    const int tid = threadIdx.x;
    for (int q = tid; q < BAM_LOGTAB_N; q += BAM_BLOCK) sLogT[q] = p.logTab[q];
    constexpr bool FM = MODEL == BAMGPU_MDL_VEGETATION || MODEL == BAMGPU_MDL_TEXTFILE;  // table-driven log / exp inside the model
    if (MODEL == BAMGPU_MDL_BARATIN || FM)
        for (int q = tid; q < BAM_EXPTAB_N; q += BAM_BLOCK) sExpT[q] = p.expTab[q];
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT);
    const FmTab fm = FM ? FmTab{lt, et} : FmTab{0u, 0u};
    const int tx = tid % TX, ty = tid / TX, TY = BAM_BLOCK / TX;
    const int lane = tid & 31, wrow = tx >> 5;
    const int c0 = blockIdx.y * CT;
    const int nc = min(CT, K - c0);
    for (int idx = tid; idx < nc * tabStride; idx += BAM_BLOCK) stab[idx] = tab[(size_t)c0 * tabStride + idx];
    for (int idx = tid; idx < nc; idx += BAM_BLOCK) sst[idx] = status[c0 + idx];
    __syncthreads();

    const int n = p.n;
    const int i = blockIdx.x * TX + tx;
    const bool act = i < n;
    // this thread's observation, kept in registers for all chains of the tile
    double xo[NX], xu[NX], xb[NX], yo[NY], yu[NY], yb[NY];
    int xbi[NX], ybi[NY], lat[NX], combo = 0;
#pragma unroll
    for (int j = 0; j < NX; ++j) { xo[j] = 0; xu[j] = 0; xb[j] = 0; xbi[j] = 0; lat[j] = -1; }
#pragma unroll
    for (int j = 0; j < NY; ++j) { yo[j] = 0; yu[j] = 0; yb[j] = 0; ybi[j] = 0; }
    if (act) {
        combo = (p.nCombo > 1) ? p.comboOf[i] : 0;
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const size_t q = i + (size_t)j * n;
            xo[j] = p.X[q];
            if (p.hasLatent) { xu[j] = p.Xu[q]; lat[j] = p.latIdx[q]; }
            if (p.hasXbias) { xb[j] = p.Xb[q]; xbi[j] = p.Xbindx[q]; }
        }
#pragma unroll
        for (int j = 0; j < NY; ++j) {
            const size_t q = i + (size_t)j * n;
            yo[j] = p.Y[q];
            yu[j] = p.Yu[q];
            if (p.hasYbias) { yb[j] = p.Yb[q]; ybi[j] = p.Ybindx[q]; }
        }
    }

    // hyper term of a latent input: log N(xo; mu, xu) = -log sqrt(2 pi) - log xu - 0.5 ((xo - mu) / xu)^2; log xu and
    // 1 / xu belong to the observation, not to the chain: computed once per thread
    double hyC[NX], hyR[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        hyC[j] = 0.0; hyR[j] = 0.0;
        if (p.hasLatent && xu[j] > 0.0) { hyC[j] = -BAM_LOG_SQRT_2PI - log(xu[j]); hyR[j] = 1.0 / xu[j]; }
    }
    // latent "true" inputs are per-chain state in HBM (16 B per chain x observation in cfg 3): the values of the NEXT
    // chain of the tile are requested before the current chain is evaluated, so the load latency overlaps the math
    double xlat[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) xlat[j] = 0.0;
    if (p.hasLatent && act && ty < nc) {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            if (lat[j] >= 0) xlat[j] = x[(size_t)(c0 + ty) * p.P + lat[j]];
    }
    for (int cc = ty; cc < nc; cc += TY) {
        const int c = c0 + cc;
        const double* __restrict__ row = stab + (size_t)cc * tabStride;
        double lk = 0.0, hy = 0.0;
        bool bad = false, badh = false;
        double xnext[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) xnext[j] = 0.0;
        if (p.hasLatent && act && cc + TY < nc) {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                if (lat[j] >= 0) xnext[j] = x[(size_t)(c + TY) * p.P + lat[j]];
        }
        if (act && sst[cc] == 0) {
            const double dl = (comp >= 0) ? delta[c] : 0.0;
            // UnfoldParVector (:3429-3452): latent cell from the vector, else un-bias the observed input
            double xt[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double v = xo[j];
                if (p.hasLatent && lat[j] >= 0) {
                    v = xlat[j];
                    if (lat[j] == comp) v += dl;
                } else if (p.hasXbias && xbi[j] > 0) {
                    v = xo[j] - xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
                }
                xt[j] = v;
            }
            const double* __restrict__ th = row + p.tabCombo + (size_t)combo * p.comboStride;
            double yh[NY];
            // BaRatin (few chains on many gaugings never reach the chain-parallel fast path): same exits as
            // baratin_apply, a (H-b)^c through the table-driven functions instead of libdevice log + exp
            const bool ok = (MODEL == BAMGPU_MDL_BARATIN) ? baratin_apply_t(p, xt[0], th, th + p.ntheta, yh[0], lt, et)
                                                          : model_apply<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh, fm);
            if (!ok) bad = true;
            else {
#pragma unroll
                for (int j = 0; j < NY; ++j) {  // GetLogLikelihood inner loop (:3204-3227)
                    if (p.hasMissing && yo[j] == p.mv) continue;
                    double sig = sigmafunk(p.sigfunc[j], row + p.tabGamma + p.gamOff[j], yh[j]);
                    if (!(sig > 0.0)) { bad = true; continue; }
                    const double var = fma(yu[j], yu[j], sig * sig);  // sigma_Y^2 + sigma_remnant^2 (:3213)
                    double mu = yh[j];
                    if (p.hasYbias && ybi[j] != 0) mu = yh[j] + yb[j] * row[p.tabOffYb[j] + ybi[j] - 1];
                    lk += gauss_logpdf_var_t(yo[j], mu, var, lt);
                }
                if (p.hasLatent) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) {  // GetHyperLogPdf (:3285-3303)
                        if (xu[j] == 0.0) continue;
                        if (!(xu[j] > 0.0)) { badh = true; continue; }
                        double mu = xt[j];
                        if (p.hasXbias && xbi[j] != 0) mu = xt[j] + xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
                        const double z = (xo[j] - mu) * hyR[j];
                        hy += fma(-0.5 * z, z, hyC[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) xlat[j] = xnext[j];
        // warp tree (fixed order) -> one value per (warp, chain)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lk += __shfl_down_sync(0xffffffffu, lk, off);
        if (p.hasLatent) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) hy += __shfl_down_sync(0xffffffffu, hy, off);
        }
        const unsigned anyBad = __ballot_sync(0xffffffffu, bad), anyBadH = __ballot_sync(0xffffffffu, badh);
        if (lane == 0) {
            red[(size_t)wrow * CT + cc] = lk;
            red[(size_t)(warpsPerRow + wrow) * CT + cc] = hy;
            if (anyBad | anyBadH) atomicOr(&statusOut[c], (anyBad ? ST_LKH_INFEAS : 0) | (anyBadH ? ST_HYPER_INFEAS : 0));
        }
    }
    __syncthreads();
    for (int cc = tid; cc < nc; cc += BAM_BLOCK) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < warpsPerRow; ++w) {
            a += red[(size_t)w * CT + cc];
            b += red[(size_t)(warpsPerRow + w) * CT + cc];
        }
        const size_t o = ((size_t)blockIdx.x * K + (c0 + cc)) * 2;
        part[o] = a;
        part[o + 1] = b;
    }
}

// ------------------------------------------------------------------------------------------------
// K1 fast path for cfg 3: Linear model, EVERY input cell latent (TextFile_InputUncertainty / Regression with input
// uncertainty), no biases, no missing Y, one sigma function.  Per (chain, observation) the path must read the chain's
// own "true" inputs (8 NX bytes: HBM-leaning), so the kernel is organised around those loads:
//   * block = 256 threads x LL_OPT observations each (stride 256: coalesced) x a tile of LL_CT chains;
//   * the observation's (Xobs, 1/Xu, Y, Yu^2) sit in registers for all chains of the tile; the chains' theta and gamma
//     are broadcast reads from shared memory; the 2 LL_CT (likelihood, hyper) sums stay in REGISTERS across the
//     thread's observations, so the cross-thread reduction (shuffle tree + fixed-order cross-warp sum) is paid once
//     per LL_OPT x LL_CT units instead of once per unit as in logpost_main;
//   * the LL_CT x NX loads of one observation are independent: 2 LL_CT requests in flight per thread;
//   * blockIdx.x walks the chain tiles, so the blocks that share an observation tile run together and the
//     observation tables are served by L2.
// Same arithmetic as logpost_main (GetLogLikelihood :3204-3227, GetHyperLogPdf :3285-3303) except that the
// chain-independent constants (-log sqrt(2 pi) - log Xu per cell) are summed once per thread and the factor -1/2 is
// applied to the thread's sum: results agree to ~1e-15 relative, not bitwise.
// Algorithmic bytes per (chain, observation): 8 NX (latent inputs); observation tables 8 (2 NX + 2 NY) / LL_CT.
// ------------------------------------------------------------------------------------------------
#ifndef LL_CT
#define LL_CT 4
#endif
#ifndef LL_OPT
#define LL_OPT 16
#endif
#ifndef LL_MINB
#define LL_MINB 2
#endif
template <int NX, int NY, int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK, LL_MINB) logpost_linear_latent(DevProblem p, const double* __restrict__ x, int K,
                                                                      const double* __restrict__ tab,
                                                                      const int* __restrict__ status,
                                                                      double* __restrict__ part, int* statusOut) {
    constexpr int NG = (SIGF == BAMGPU_SIG_LINEAR) ? 2 : 1;
    __shared__ double sth[LL_CT][NX * NY];
    __shared__ double sg[LL_CT][NG * NY];
    __shared__ int sst[LL_CT];
    __shared__ double red[BAM_BLOCK / 32][2 * LL_CT];
    __shared__ __align__(16) double2 sLogT[BAM_LOGTAB_N];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * LL_CT, nc = min(LL_CT, K - c0);
    for (int q = tid; q < BAM_LOGTAB_N; q += BAM_BLOCK) sLogT[q] = p.logTab[q];
    for (int q = tid; q < LL_CT * NX * NY; q += BAM_BLOCK) {
        const int cc = q / (NX * NY), i = q % (NX * NY);
        sth[cc][i] = (cc < nc) ? tab[(size_t)(c0 + cc) * p.tabStride + p.tabCombo + i] : 0.0;
    }
    for (int q = tid; q < LL_CT * NG * NY; q += BAM_BLOCK) {
        const int cc = q / (NG * NY), i = q % (NG * NY), j = i / NG, m = i % NG;
        sg[cc][i] = (cc < nc) ? tab[(size_t)(c0 + cc) * p.tabStride + p.tabGamma + p.gamOff[j] + m] : 0.0;
    }
    if (tid < LL_CT) sst[tid] = (tid < nc) ? status[c0 + tid] : 1;
    __syncthreads();
    const unsigned lt = smem_addr(sLogT);
    const int n = p.n;
    const size_t latBase = (size_t)p.nFit + p.nGamma;  // every cell latent: cell (i, j) sits at latBase + j n + i (Packer :4202-4242)
    const double* __restrict__ xrow[LL_CT];
#pragma unroll
    for (int cc = 0; cc < LL_CT; ++cc) xrow[cc] = x + (size_t)(c0 + min(cc, nc - 1)) * p.P + latBase;

    double accL[LL_CT], accH[LL_CT];
#pragma unroll
    for (int cc = 0; cc < LL_CT; ++cc) { accL[cc] = 0.0; accH[cc] = 0.0; }
    unsigned sgnAcc[LL_CT], zerAcc[LL_CT];
#pragma unroll
    for (int cc = 0; cc < LL_CT; ++cc) { sgnAcc[cc] = 0u; zerAcc[cc] = 0xffffffffu; }
    double hyConst = 0.0;  // sum over this thread's cells of -log sqrt(2 pi) - log Xu (DevProblem::hyC)
    int nObs = 0;
    unsigned badMask = 0;
    const int i0 = blockIdx.y * (BAM_BLOCK * LL_OPT) + tid;
    // software pipeline: the tables and the LL_CT x NX latent inputs of observation o+1 are requested before
    // observation o is evaluated, so the HBM latency overlaps ~LL_CT x 130 instructions of arithmetic
    double xoN[NX], hrN[NX], yoN[NY], yuN[NY], hcN = 0.0, xlN[LL_CT][NX];
    auto request = [&](int i) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            xoN[j] = p.X[i + (size_t)j * n];
            hrN[j] = p.XuInv[i + (size_t)j * n];
        }
#pragma unroll
        for (int j = 0; j < NY; ++j) {
            yoN[j] = p.Y[i + (size_t)j * n];
            yuN[j] = p.Yu[i + (size_t)j * n];
        }
        hcN = p.hyC[i];
#pragma unroll
        for (int cc = 0; cc < LL_CT; ++cc)
#pragma unroll
            for (int j = 0; j < NX; ++j) xlN[cc][j] = __ldcs(xrow[cc] + (size_t)j * n + i);
    };
    if (i0 < n) request(i0);
#pragma unroll 1
    for (int o = 0; o < LL_OPT; ++o) {
        const int i = i0 + o * BAM_BLOCK;
        if (i >= n) break;
        double xo[NX], hr[NX], yo[NY], yu2[NY], xl[LL_CT][NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) { xo[j] = xoN[j]; hr[j] = hrN[j]; }
#pragma unroll
        for (int j = 0; j < NY; ++j) { yo[j] = yoN[j]; yu2[j] = yuN[j] * yuN[j]; }
        hyConst += hcN;
#pragma unroll
        for (int cc = 0; cc < LL_CT; ++cc)
#pragma unroll
            for (int j = 0; j < NX; ++j) xl[cc][j] = xlN[cc][j];
        ++nObs;
        if (o + 1 < LL_OPT && i + BAM_BLOCK < n) request(i + BAM_BLOCK);
        // straight-line code for all LL_CT chains (no branches: the scheduler interleaves the independent chains); a
        // chain that is already infeasible is computed like the others and discarded at the end.
        // sigma > 0 is tracked per chain on the integer words of sigma (OR of the sign bits, minimum of hi|lo): a NaN
        // sigma needs no test, it ends in a NaN sum, which logpost_final flags like logpost_main's explicit test.
        // NY == 2: the two outputs share ONE log and ONE reciprocal through the product of their variances,
        //   log v0 + log v1 + r0^2/v0 + r1^2/v1 = log(v0 v1) + (r0^2 v1 + r1^2 v0) / (v0 v1);
        // a product outside the normal range (never with sane variances) re-evaluates the observation per output.
        double T[LL_CT];
        unsigned rng = 0;
        auto outputs = [&](int cc, double* v, double* r2) {
#pragma unroll
            for (int j = 0; j < NY; ++j) {
                double yh = 0.0;  // LM_Apply: Y = X . reshape(theta, (nX, nY))
#pragma unroll
                for (int k = 0; k < NX; ++k) yh = yh + xl[cc][k] * sth[cc][j * NX + k];
                double sig;
                if (SIGF == BAMGPU_SIG_CONSTANT) sig = sg[cc][j];
                else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(sg[cc][2 * j + 1], fabs(yh), sg[cc][2 * j]);
                else sig = sg[cc][j] * fabs(yh);
                sgnAcc[cc] |= (unsigned)__double2hiint(sig);
                zerAcc[cc] = min(zerAcc[cc], (unsigned)__double2hiint(sig) | (unsigned)__double2loint(sig));
                v[j] = fma(sig, sig, yu2[j]);
                const double r = yo[j] - yh;
                r2[j] = r * r;
            }
        };
#pragma unroll
        for (int cc = 0; cc < LL_CT; ++cc) {
            double v[NY], r2[NY];
            outputs(cc, v, r2);
            if (NY == 2) {
                const double vp = v[0] * v[1];
                const double rp = trcp(vp);
                rng = max(rng, (unsigned)__double2hiint(vp) - 0x00100000u);
                T[cc] = fma(fma(r2[0], v[1], r2[1] * v[0]), rp, tlog_pos(vp, lt));
            } else {
                double t = 0.0;
#pragma unroll
                for (int j = 0; j < NY; ++j) {
                    rng = max(rng, (unsigned)__double2hiint(v[j]) - 0x00100000u);
                    t += fma(r2[j], trcp(v[j]), tlog_pos(v[j], lt));
                }
                T[cc] = t;
            }
            double H = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const double z = (xo[j] - xl[cc][j]) * hr[j];
                H = fma(z, z, H);
            }
            accH[cc] += H;
        }
        if (rng >= 0x7fe00000u) {  // rare
#pragma unroll
            for (int cc = 0; cc < LL_CT; ++cc) {
                double v[NY], r2[NY], t = 0.0;
                outputs(cc, v, r2);
#pragma unroll
                for (int j = 0; j < NY; ++j) {
                    if ((unsigned)__double2hiint(v[j]) - 0x00100000u >= 0x7fe00000u) badMask |= 1u << cc;  // as a NaN sum in logpost_main
                    t += fma(r2[j], trcp(v[j]), tlog_pos(v[j], lt));
                }
                T[cc] = t;
            }
        }
#pragma unroll
        for (int cc = 0; cc < LL_CT; ++cc) accL[cc] += T[cc];
    }
#pragma unroll
    for (int cc = 0; cc < LL_CT; ++cc)
        if (nObs > 0 && ((sgnAcc[cc] & 0x80000000u) || zerAcc[cc] == 0u)) badMask |= 1u << cc;
    // thread sums -> log-densities, then the block reduction (fixed tree, fixed order)
#pragma unroll
    for (int cc = 0; cc < LL_CT; ++cc) {
        const bool live = sst[cc] == 0;
        double a = live ? fma(-0.5, accL[cc], -BAM_LOG_SQRT_2PI * (double)(NY * nObs)) : 0.0;
        double b = live ? fma(-0.5, accH[cc], hyConst) : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
        }
        if (lane == 0) { red[warp][2 * cc] = a; red[warp][2 * cc + 1] = b; }
    }
    const unsigned anyBad = __reduce_or_sync(0xffffffffu, badMask);
    if (lane == 0 && anyBad)
        for (int cc = 0; cc < nc; ++cc)
            if ((anyBad >> cc) & 1u) atomicOr(&statusOut[c0 + cc], ST_LKH_INFEAS);
    __syncthreads();
    if (tid < 2 * LL_CT) {
        const int cc = tid >> 1;
        if (cc < nc) {
            double a = 0.0;
            for (int w = 0; w < BAM_BLOCK / 32; ++w) a += red[w][tid];
            part[((size_t)blockIdx.y * K + (c0 + cc)) * 2 + (tid & 1)] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1b: final sum + early-exit priority (GetPostLogPdf :3358-3380)
// ------------------------------------------------------------------------------------------------
#define FINAL_STRIPES 8
__global__ void __launch_bounds__(32 * FINAL_STRIPES) logpost_final(DevProblem p, int K, int nTiles, const double* __restrict__ part,
                                                              const double* __restrict__ prior, int* __restrict__ status,
                                                              double* __restrict__ lkh, double* __restrict__ hyper,
                                                              double* __restrict__ fx) {
    // 32 chains per block; FINAL_STRIPES warps each sum every FINAL_STRIPES-th tile, then a fixed-order combine:
    // deterministic, and 8x more loads in flight than one thread per chain.
    __shared__ double sa[FINAL_STRIPES][32], sb[FINAL_STRIPES][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double a = 0.0, b = 0.0;
    if (c < K)
        for (int t = w; t < nTiles; t += FINAL_STRIPES) {
            const double2 v = *reinterpret_cast<const double2*>(part + ((size_t)t * K + c) * 2);
            a += v.x;
            b += v.y;
        }
    sa[w][lane] = a;
    sb[w][lane] = b;
    __syncthreads();
    if (w != 0 || c >= K) return;
    a = 0.0;
    b = 0.0;
#pragma unroll
    for (int s = 0; s < FINAL_STRIPES; ++s) { a += sa[s][lane]; b += sb[s][lane]; }
    int st = status[c];
    if (!st && !(a == a && b == b)) st |= ST_LKH_INFEAS;  // a NaN sum (e.g. NaN sigma) is an infeasible vector
    if (p.hyperInfeasible && !(st & (ST_PRIOR_INFEAS | ST_PRIOR_NULL | ST_LKH_INFEAS))) st |= ST_HYPER_INFEAS;
    status[c] = st;
    lkh[c] = a;
    hyper[c] = b;
    fx[c] = st ? BAMGPU_UNDEF_RN : (prior[c] + a + b);
}

// K1b for the fast path: per-combination sums (fixed stripes inside a combination, combinations in order), so that
//   * a full and a partial evaluation of the same state give bitwise the same log-posterior, and
//   * a proposal on a VAR parameter re-uses the cached sums of the combinations it does not touch.
__global__ void __launch_bounds__(32 * FINAL_STRIPES) logpost_final_combo(DevProblem p, int K, const int* __restrict__ comboTile0,
                                                                    const double* __restrict__ part,
                                                                    const double* __restrict__ prior, int* __restrict__ status,
                                                                    const int* __restrict__ affected,
                                                                    const double* __restrict__ lkCcur, double* __restrict__ lkCcand,
                                                                    double* __restrict__ lkh, double* __restrict__ hyper,
                                                                    double* __restrict__ fx, int* __restrict__ arrive) {
    // grid = (32-chain groups, VAR combinations): one block sums the tiles of ONE combination (FINAL_STRIPES warps, every
    // FINAL_STRIPES-th tile each, four loads in flight, fixed-order combine); the last block of a chain group to finish
    // adds the combinations in order.  With one block per chain group looping over the combinations the kernel was a
    // chain of dependent L2 loads on 16 blocks when chains are few (512 per GPU: 25 us of a 180 us step).
    __shared__ double sa[FINAL_STRIPES][32];
    __shared__ int sLast;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int v = blockIdx.y;
    const bool redo = affected == nullptr || affected[v] != 0;  // uniform over the block
    if (redo) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (c < K) {
            const int t1 = comboTile0[v + 1];
            int t = comboTile0[v] + w;
            // a fixed association: stripe w adds its tiles four at a time into four chains, then (a0 + a1) + (a2 + a3)
            for (; t + 3 * FINAL_STRIPES < t1; t += 4 * FINAL_STRIPES) {
                a0 += part[((size_t)t * K + c)];
                a1 += part[((size_t)(t + FINAL_STRIPES) * K + c)];
                a2 += part[((size_t)(t + 2 * FINAL_STRIPES) * K + c)];
                a3 += part[((size_t)(t + 3 * FINAL_STRIPES) * K + c)];
            }
            for (; t < t1; t += FINAL_STRIPES) a0 += part[((size_t)t * K + c)];
        }
        sa[w][lane] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (w == 0 && c < K) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < FINAL_STRIPES; ++q) s += sa[q][lane];
            lkCcand[(size_t)v * K + c] = s;
        }
    } else if (w == 0 && c < K) {
        lkCcand[(size_t)v * K + c] = lkCcur[(size_t)v * K + c];
    }
    // last block of this chain group: total over the combinations, in order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(&arrive[blockIdx.x], 1);
        sLast = (old == (int)gridDim.y - 1);
        if (sLast) arrive[blockIdx.x] = 0;  // re-armed for the next launch
    }
    __syncthreads();
    if (!sLast || w != 0 || c >= K) return;
    __threadfence();
    double total = 0.0;
    for (int q = 0; q < p.nCombo; ++q) total += __ldcg(&lkCcand[(size_t)q * K + c]);
    int st = status[c];
    if (!st && !(total == total)) st |= ST_LKH_INFEAS;
    status[c] = st;
    lkh[c] = total;
    hyper[c] = 0.0;
    fx[c] = st ? BAMGPU_UNDEF_RN : (prior[c] + total + 0.0);
}

// ------------------------------------------------------------------------------------------------
// K1 fast path: BaRatin, chain-parallel ("thread = chain") with observations broadcast from shared memory.
//
// Used for the headline configuration (many chains, no latent inputs / biases, Constant | Linear |
// Proportional remnant sigma).  A thread owns one chain: the parameters of the current VAR combination
// (k, b, c, log a) live in registers; a block of 256 chains walks one tile of observations that was staged
// ONCE in shared memory (H, Q, uQ^2): every shared-memory read is a broadcast.  The likelihood partial of a
// chain is accumulated in a register -- no shuffles, no cross-warp reduction -- and written coalesced.
// Lanes of a warp evaluate the same observation for neighbouring chains, so they agree on the stage range
// except within the posterior spread of the activation stages k: divergence is negligible.
// a_i (H-b_i)^c_i = texp(c_i tlog(H-b_i) + log a_i) and -0.5 log(v) via tlog: ~60 FP64-pipe instructions per
// chain x observation against ~330 for the libdevice pow/log/sqrt/div formulation.
// ------------------------------------------------------------------------------------------------
#ifndef BAM_CB_MINB
#define BAM_CB_MINB 2
#endif
#ifndef BAM_CB_TOBS
#define BAM_CB_TOBS 256  /* observations per tile (halved while the grid would leave SMs idle); 160 / 256 / 320 / 640: 1.467 / 1.464 / 1.460 / 1.482 ms */
#endif
#ifndef BAM_CB_ILP
#define BAM_CB_ILP 5  /* observations per loop iteration: 3 / 4 / 5 / 6 measured 1.535 / 1.496 / 1.467 / 1.491 ms (cfg 2) */
#endif
template <int NC, int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK, BAM_CB_MINB) logpost_baratin_cb(DevProblem p, int K, int TOBS, const int* __restrict__ tStart,
                                                                    const int* __restrict__ tEnd, const int* __restrict__ tCombo,
                                                                    const int* __restrict__ tileIds, const double* __restrict__ tab,
                                                                    const int* status, double* __restrict__ part,
                                                                    int* statusOut) {
    extern __shared__ __align__(16) double sm[];
    double2* sLog = reinterpret_cast<double2*>(sm);  // BAM_LOGTAB_N x 16 B
    double* sExp = sm + 2 * BAM_LOGTAB_N;
    double* sH = sExp + BAM_EXPTAB_N;
    double* sY = sH + TOBS;
    double* sV = sY + TOBS;                          // uQ^2

    const int tid = threadIdx.x;
    // tiles never straddle VAR combinations (host tile table); with a tile list only the listed tiles are evaluated
    const int tile = tileIds ? tileIds[blockIdx.x] : blockIdx.x;
    const int i0 = tStart[tile];
    const int nt = tEnd[tile] - i0;
    for (int i = tid; i < BAM_LOGTAB_N; i += BAM_BLOCK) sLog[i] = p.logTab[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExp[i] = p.expTab[i];
    for (int i = tid; i < nt; i += BAM_BLOCK) {
        const double u = p.Yu[i0 + i];
        sH[i] = p.X[i0 + i];
        sY[i] = p.Y[i0 + i];
        sV[i] = u * u;
    }
    __syncthreads();
    const int c = blockIdx.y * BAM_BLOCK + tid;
    if (c >= K) return;
    const size_t o = ((size_t)tile * K + c) * (p.partPacked ? 1 : 2);
    if (!p.partPacked) part[o + 1] = 0.0;
    if (status[c] != 0) { part[o] = 0.0; return; }
    const double* __restrict__ row = tab + (size_t)c * p.tabStride;
    const double g1 = row[p.tabGamma], g2 = (SIGF == BAMGPU_SIG_LINEAR) ? row[p.tabGamma + 1] : 0.0;
    // control matrix as nibbles: nibble r = row r-1 (nibble 0 = no control: H below every activation stage)
    unsigned mask4 = 0;
#pragma unroll
    for (int r = 1; r <= NC; ++r) mask4 |= ((unsigned)(p.ctrlMask >> ((r - 1) * 8)) & 0xfu) << (4 * r);
    const unsigned lt = smem_addr(sLog), et = smem_addr(sExp);
    double s2 = 0.0;           // sum of squared standardised residuals
    double pmu[BAM_CB_ILP];    // mantissa products of the variances
    int pe = 0;                // sum of their biased exponents
#pragma unroll
    for (int u = 0; u < BAM_CB_ILP; ++u) pmu[u] = 1.0;
    bool bad = false;
    int cnt = 0;

    {
        const int v = tCombo[tile];
        const int lo = 0, hi = nt;
        cnt = nt;
        const double* __restrict__ cr = row + p.tabCombo + (size_t)v * p.comboStride;
        double k[NC], b[NC], cc[NC], la[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) { k[j] = cr[3 * j]; cc[j] = cr[3 * j + 2]; b[j] = cr[3 * NC + j]; la[j] = cr[4 * NC + j]; }

        // GetHRange (BaRatin_model.f90:398-436): k is increasing, so range = #{k_j <= H}
        auto range = [&](double h) -> int {
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
        // BaRatin_computeQ (:236-249).  For a parameter set that passed ApplyContinuity every control j < range has
        // H >= k_(range) >= k_j >= b_j (k increasing, b_j <= k_j, and k_j >= b_i for i < j are exactly the feasibility
        // tests of :380-391), so H - b_j >= 0 holds in floating point as well: the "H-b<0" exits of the reference can
        // never fire here and the loop reduces to the sum over the active controls.
        auto q_checked = [&](double h, unsigned m) -> double {
            double q = 0.0;
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if ((m >> j) & 1u) q += texp(fma(cc[j], tlog(h - b[j], lt), la[j]), et);
            return q;
        };
        // Gaussian log-density of one observation given Q (GetLogLikelihood :3208-3225), variance form.  A
        // non-positive sigma makes the vector infeasible (:3210): flagged through the sign of (high word - 1), the
        // (garbage) contribution is still accumulated because the sum of an infeasible vector is never read.
        int badAcc = 0;
        // The sum of the log-variances is accumulated as ONE product of mantissas (BAM_CB_ILP independent chains, one per
        // observation slot of the unrolled loop) plus an integer sum of exponents: log(prod m_i) + ln2 * sum e_i costs
        // one DMUL and three integer instructions per observation instead of the eight FP64 instructions of a
        // table-driven log.  A mantissa product grows by < 2 per factor, i.e. < 2^64 per chain and tile: no overflow.
        auto tail = [&](double q, double y, double v2, double& pm) {
            double sig;
            if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
            else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
            else sig = g1 * fabs(q);
            { const int hs = __double2hiint(sig); badAcc |= hs | (hs - 1); }  // negative iff sig <= 0 (or denormal)
            const double var = fma(sig, sig, v2);
            const int vh = __double2hiint(var);
            pe += vh >> 20;  // biased exponent (var > 0); the bias is removed once per tile
            pm *= __hiloint2double((vh & 0x000fffff) | 0x3ff00000, __double2loint(var));
            const double res = y - q;
            s2 = fma(res * res, trcp(var), s2);
        };

        // observations are sorted by stage inside a combination, so the range only moves up: keep (row mask, next
        // activation stage) in registers and re-derive them only when an observation crosses the next stage.
        int r = range(sH[lo]);
        unsigned m = (mask4 >> (4 * r)) & 0xfu;
        double knext = next_stage(r);
        int i = lo;
        for (; i + (BAM_CB_ILP - 1) < hi; i += BAM_CB_ILP) {  // ILP observations per iteration: independent chains
            double hh[BAM_CB_ILP], qq[BAM_CB_ILP];
#pragma unroll
            for (int u = 0; u < BAM_CB_ILP; ++u) hh[u] = sH[i + u];
            if (hh[BAM_CB_ILP - 1] >= knext) {  // rare: a stage boundary inside the group
#pragma unroll
                for (int u = 0; u < BAM_CB_ILP; ++u) qq[u] = q_checked(hh[u], (mask4 >> (4 * range(hh[u]))) & 0xfu);
                r = range(hh[BAM_CB_ILP - 1]);
                m = (mask4 >> (4 * r)) & 0xfu;
                knext = next_stage(r);
            } else {
                unsigned oor = 0;
#pragma unroll
                for (int u = 0; u < BAM_CB_ILP; ++u) qq[u] = 0.0;
#pragma unroll
                for (int j = 0; j < NC; ++j)
                    if ((m >> j) & 1u) {
                        double xx[BAM_CB_ILP];
#pragma unroll
                        for (int u = 0; u < BAM_CB_ILP; ++u) {
                            xx[u] = fma(cc[j], tlog(hh[u] - b[j], lt), la[j]);
                            oor = max(oor, (unsigned)__double2hiint(xx[u]) & 0x7fffffffu);
                        }
#pragma unroll
                        for (int u = 0; u < BAM_CB_ILP; ++u) qq[u] += texp_inrange(xx[u], et);
                    }
                if (oor >= BAM_EXP_OOR) {  // rare: H == b exactly (log 0) or an overflowing power
#pragma unroll
                    for (int u = 0; u < BAM_CB_ILP; ++u) qq[u] = q_checked(hh[u], m);
                }
            }
#pragma unroll
            for (int u = 0; u < BAM_CB_ILP; ++u) tail(qq[u], sY[i + u], sV[i + u], pmu[u]);
        }
        for (; i < hi; ++i) {
            const double h0 = sH[i];
            tail(q_checked(h0, (mask4 >> (4 * range(h0))) & 0xfu), sY[i], sV[i], pmu[0]);
        }
        bad |= badAcc < 0;
    }
    // sum_i log N(y_i; q_i, sqrt(var_i)) = -cnt log sqrt(2 pi) - 0.5 [ sum log var_i + sum res_i^2 / var_i ]
    double pm = pmu[0];
#pragma unroll
    for (int u = 1; u < BAM_CB_ILP; ++u) pm *= pmu[u];
    const double lsum = fma((double)(pe - 1023 * cnt), kFM[4], tlog_pos(pm, lt));
    part[o] = fma(-0.5, lsum + s2, -BAM_LOG_SQRT_2PI * (double)cnt);
    if (bad) atomicOr(&statusOut[c], ST_LKH_INFEAS);
}

using FastKernel = void (*)(DevProblem, int, int, const int*, const int*, const int*, const int*, const double*, const int*,
                            double*, int*);

template <int NC>
FastKernel pick_fast_sig(int sigf) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return logpost_baratin_cb<NC, BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return logpost_baratin_cb<NC, BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return logpost_baratin_cb<NC, BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}


// ------------------------------------------------------------------------------------------------
// K1 fast path: Sediment (cfg 4 shapes: a few hundred rows, thousands of chains), chain-parallel.
//
// With pseudo-parameters FIX / IN and fixed css coefficients every formula of Sediment_Apply is
// t1 B(row) exp(t2 L1(row) + t3 L2(row)) (bam_sediment_fast.cuh): (B, L1, L2, row status) are computed ONCE per handle
// with the reference's libm-class functions, and a chain x row costs one table-driven exp + the Gaussian tail (~30 FP64
// instructions) instead of logpost_main's three pow, two exp, a cube root and a square root (~600) per chain x row:
// 812 rows x 4096 chains went from 0.29 ms to the latency of a launch.  thread = chain, rows staged in shared memory and
// read as broadcasts; one partial per (tile, chain), summed by logpost_final in fixed order.
// A row whose inputs are infeasible makes every vector infeasible (ModelLibrary_tools.f90:421-423), missing Y or not.
// ------------------------------------------------------------------------------------------------
#define SED_TOBS 64
template <int SIGF>
__global__ void __launch_bounds__(BAM_BLOCK) logpost_sediment_cb(DevProblem p, int K, const SedObs* __restrict__ obs,
                                                                 const double* __restrict__ tab, const int* status,
                                                                 double* __restrict__ part, int* statusOut) {
    __shared__ __align__(16) double sLog[2 * BAM_LOGTAB_N];
    __shared__ double sExp[BAM_EXPTAB_N];
    __shared__ SedObs sObs[SED_TOBS];
    __shared__ double sY[SED_TOBS], sV[SED_TOBS];
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int i0 = tile * SED_TOBS, nt = min(SED_TOBS, p.n - i0);
    for (int i = tid; i < BAM_LOGTAB_N; i += BAM_BLOCK) reinterpret_cast<double2*>(sLog)[i] = p.logTab[i];
    for (int i = tid; i < BAM_EXPTAB_N; i += BAM_BLOCK) sExp[i] = p.expTab[i];
    for (int i = tid; i < nt; i += BAM_BLOCK) {
        sObs[i] = obs[i0 + i];
        sY[i] = p.Y[i0 + i];
        const double yu = p.Yu[i0 + i];
        sV[i] = yu * yu;
    }
    __syncthreads();
    const int c = blockIdx.y * BAM_BLOCK + tid;
    if (c >= K) return;
    const size_t o = ((size_t)tile * K + c) * 2;
    part[o + 1] = 0.0;
    if (status[c] != 0) { part[o] = 0.0; return; }
    const double* __restrict__ row = tab + (size_t)c * p.tabStride;
    const double* __restrict__ th = row + p.tabCombo;
    const double t1 = th[0], e1 = th[1];
    const double e2 = (p.sedFormula == BAMGPU_SED_CAMENEN || p.sedFormula == BAMGPU_SED_SMART) ? th[2] : 0.0;
    const double g1 = row[p.tabGamma], g2 = (SIGF == BAMGPU_SIG_LINEAR) ? row[p.tabGamma + 1] : 0.0;
    const unsigned lt = smem_addr(sLog), et = smem_addr(sExp);
    double acc = 0.0;
    bool bad = false;
    for (int i = 0; i < nt; ++i) {
        const SedObs ob = sObs[i];
        if (ob.st == 2.0) { bad = true; continue; }  // infeasible row: the vector is infeasible
        const double q = (ob.st == 1.0) ? 0.0 : ob.B * t1 * texp(fma(e1, ob.L1, e2 * ob.L2), et);
        if (sY[i] == p.mv) continue;                   // missing output: no likelihood term (:3207)
        double sig;
        if (SIGF == BAMGPU_SIG_CONSTANT) sig = g1;
        else if (SIGF == BAMGPU_SIG_LINEAR) sig = fma(g2, fabs(q), g1);
        else sig = g1 * fabs(q);
        if (!(sig > 0.0)) { bad = true; continue; }   // :3211
        acc += gauss_logpdf_var_t(sY[i], q, fma(sig, sig, sV[i]), lt);
    }
    part[o] = acc;
    if (bad) atomicOr(&statusOut[c], ST_LKH_INFEAS);
}

using SedLogpostKernel = void (*)(DevProblem, int, const SedObs*, const double*, const int*, double*, int*);
SedLogpostKernel pick_sed_logpost(const DevProblem& p, int K) {
    if (!sed_fast_form(p) || p.hasLatent || p.hasXbias || p.hasYbias || p.nY != 1 || p.nCombo != 1 || K < 128 || p.n < 1) return nullptr;
    switch (p.sigfunc[0]) {
        case BAMGPU_SIG_CONSTANT: return logpost_sediment_cb<BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return logpost_sediment_cb<BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return logpost_sediment_cb<BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}

FastKernel pick_fast(const DevProblem& p, int K) {
    if (p.model != BAMGPU_MDL_BARATIN || p.hasLatent || p.hasXbias || p.hasYbias || p.hasMissing || K < 128 || p.nCombo > 4096) return nullptr;
    switch (p.ncontrol) {
        case 1: return pick_fast_sig<1>(p.sigfunc[0]);
        case 2: return pick_fast_sig<2>(p.sigfunc[0]);
        case 3: return pick_fast_sig<3>(p.sigfunc[0]);
        case 4: return pick_fast_sig<4>(p.sigfunc[0]);
    }
    return nullptr;
}

Fast2Kernel pick_fast2(const DevProblem& p, int K) {
    if (!pick_fast(p, K) || !p.logTab2) return nullptr;
    switch (p.ncontrol) {
        case 1: return pick_fast2_sig<1>(p.sigfunc[0]);
        case 2: return pick_fast2_sig<2>(p.sigfunc[0]);
        case 3: return pick_fast2_sig<3>(p.sigfunc[0]);
        case 4: return pick_fast2_sig<4>(p.sigfunc[0]);
    }
    return nullptr;
}

using MainKernel = void (*)(DevProblem, const double*, int, int, int, const double*, const int*, double*, int*, int,
                            const double*, int);

MainKernel pick_main(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return logpost_main<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}


// ------------------------------------------------------------------------------------------------
// GetSim_calibration (src/BaM_tools.f90:3619-3680): the model outputs on the calibration inputs for ONE parameter
// vector, in the caller's observation order (the device tables of the handle are sorted / compacted, so this
// kernel works on temporaries built from the caller's own arrays).  thread <-> observation.
// ------------------------------------------------------------------------------------------------
template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(BAM_BLOCK) calib_sim(DevProblem p, int n, const double* __restrict__ Xtrue,
                                                       const int* __restrict__ comboOf, const double* __restrict__ row,
                                                       double* __restrict__ Ysim, double* __restrict__ state,
                                                       int* __restrict__ nBad) {
    const int i = blockIdx.x * BAM_BLOCK + threadIdx.x;
    if (i >= n) return;
    constexpr int NS = model_nstate(MODEL);
    double xt[NX], yh[NY], st[NS > 0 ? NS : 1];
#pragma unroll
    for (int j = 0; j < NX; ++j) xt[j] = Xtrue[i + (size_t)j * n];
    const int combo = (p.nCombo > 1) ? comboOf[i] : 0;
    const double* th = row + p.tabCombo + (size_t)combo * p.comboStride;
    const bool ok = model_apply_s<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh, st);
    if (!ok) atomicAdd(nBad, 1);
#pragma unroll
    for (int j = 0; j < NY; ++j) Ysim[i + (size_t)j * n] = ok ? yh[j] : p.unfeas;  // ModelLibrary_tools.f90:419-432
    if (state != nullptr)
#pragma unroll
        for (int j = 0; j < NS; ++j) state[i + (size_t)j * n] = ok ? st[j] : p.unfeas;
}

using CalibKernel = void (*)(DevProblem, int, const double*, const int*, const double*, double*, double*, int*);
CalibKernel pick_calib(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return calib_sim<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

}  // namespace

namespace bam {

using LatentLogpostKernel = void (*)(DevProblem, const double*, int, const double*, const int*, double*, int*);
template <int NX, int NY>
static LatentLogpostKernel pick_latent_sig(int sigf) {
    switch (sigf) {
        case BAMGPU_SIG_CONSTANT: return logpost_linear_latent<NX, NY, BAMGPU_SIG_CONSTANT>;
        case BAMGPU_SIG_LINEAR: return logpost_linear_latent<NX, NY, BAMGPU_SIG_LINEAR>;
        case BAMGPU_SIG_PROPORTIONAL: return logpost_linear_latent<NX, NY, BAMGPU_SIG_PROPORTIONAL>;
    }
    return nullptr;
}
// eligibility of the cfg-3 fast path (see logpost_linear_latent); comp = component shifted by a proposal, or -1
static LatentLogpostKernel pick_latent(const bamgpu_handle* h, int comp) {
    const DevProblem& p = h->dp;
    const HostLayout& L = h->L;
    if (h->forceGeneric || p.model != BAMGPU_MDL_LINEAR || !p.hasLatent || p.n < 1 || !p.XuInv) return nullptr;
    if ((long long)L.nLatent != (long long)p.n * p.nX || p.n != L.n) return nullptr;  // every cell latent, table order = Packer order
    if (p.hasXbias || p.hasYbias || p.hasMissing || p.hyperInfeasible || p.nCombo != 1) return nullptr;
    if (comp >= p.nFit + p.nGamma) return nullptr;  // a shifted latent cell is applied by logpost_main
    for (int j = 1; j < p.nY; ++j)
        if (p.sigfunc[j] != p.sigfunc[0]) return nullptr;
    if (p.nX == 1 && p.nY == 1) return pick_latent_sig<1, 1>(p.sigfunc[0]);
    if (p.nX == 2 && p.nY == 1) return pick_latent_sig<2, 1>(p.sigfunc[0]);
    if (p.nX == 1 && p.nY == 2) return pick_latent_sig<1, 2>(p.sigfunc[0]);
    if (p.nX == 2 && p.nY == 2) return pick_latent_sig<2, 2>(p.sigfunc[0]);
    return nullptr;
}

static void make_plan(bamgpu_handle* h, int K) {
    LogpostPlan& pl = h->plan;
    const int n = std::max(h->dp.n, 1);
    pl.TX = n > 128 ? 256 : (n > 64 ? 128 : (n > 32 ? 64 : 32));
    pl.TY = BAM_BLOCK / pl.TX;
    pl.nTiles = (n + pl.TX - 1) / pl.TX;
    const int stride = h->dp.tabStride;
    // chains per tile: as many as fit in ~40 KB of shared memory, at most 64, a multiple of TY,
    // but keep at least ~4 blocks per SM in flight when K is small
    int ct = 64;
    while (ct > pl.TY && (size_t)ct * stride * 8 > 40 * 1024) ct >>= 1;
    while (ct > pl.TY && (long long)pl.nTiles * ((K + ct - 1) / ct) < 148 * 4) ct >>= 1;
    ct = std::max(ct, pl.TY);
    pl.CT = ct;
    // [chain tables | cross-warp partials | chain status (int)] then, 16-byte aligned, the 4 KB log table
    const size_t head = ((size_t)ct * stride + 2 * (pl.TX / 32) * ct) * 8 + (size_t)ct * 4;
    pl.logOff = (int)((head + 15) / 16 * 2);  // in doubles
    pl.smem = (size_t)pl.logOff * 8 + sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N);
}

#define BAM_FP_MAXCOMBO 64

bool fast_path_keeps_combo_sums(const bamgpu_handle* h, int K) {
    return !h->forceGeneric && h->dp.n > 0 && pick_fast(h->dp, K) != nullptr && h->dp.nCombo <= BAM_FP_MAXCOMBO;
}

// Tile size of the persistent fast path: a whole number of Taylor sub-tiles (BAM_CB2_TSUB groups of BAM_CB2_ILP gaugings).
// Items = tiles x chain blocks are handed out dynamically to `slots` resident blocks, so the idle tail at the end of the
// launch is about half an item per block: large tiles amortise the per-item work (parameter loads, barrier, final log)
// when there are many rounds, small tiles keep the tail short when chains are few (512 per GPU in the sharded cfg 2).
#ifndef BAM_CB2_ROUNDS
#define BAM_CB2_ROUNDS 8 /* items per resident block below which the tiles get smaller */
#endif
static int pick_tile_size(const bamgpu_handle* h, int chainBlocks, int slots) {
    const int nCombo = h->dp.nCombo;
    const int sub = BAM_CB2_TSUB * BAM_CB2_ILP;
    int best = sub;
    for (int tobs = (BAM_CB_TOBS / sub) * sub; tobs >= sub; tobs -= sub) {
        long long tiles = 0;
        for (int v = 0; v < nCombo; ++v) tiles += (h->comboStartHost[v + 1] - h->comboStartHost[v] + tobs - 1) / tobs;
        if (tiles * chainBlocks >= (long long)BAM_CB2_ROUNDS * slots || tobs == sub) { best = tobs; break; }
    }
    return best;
}

// tiles of at most `tobs` observations that never straddle two VAR combinations (observations are stored sorted by
// combination): tile table on the device, tile range of every combination on host and device
static void ensure_fast_tiles(bamgpu_handle* h, int tobs) {
    if (h->fpTobs == tobs && h->d_fpStart) return;
    auto fr = [](auto*& q) { if (q) cudaFree(q); q = nullptr; };
    fr(h->d_fpStart); fr(h->d_fpEnd); fr(h->d_fpCombo); fr(h->d_fpComboTile0); fr(h->d_fpPacked); fr(h->d_fpOff); fr(h->d_fpBytes);
    for (auto*& q : h->d_compTiles) fr(q);
    for (auto*& q : h->d_compAffected) fr(q);
    h->d_compTiles.clear(); h->d_compAffected.clear(); h->compTileCnt.clear();
    const int nCombo = h->dp.nCombo;
    std::vector<int> ts, te, tc;
    h->fpComboTile0.assign(nCombo + 1, 0);
    for (int v = 0; v < nCombo; ++v) {
        h->fpComboTile0[v] = (int)ts.size();
        for (int i = h->comboStartHost[v]; i < h->comboStartHost[v + 1]; i += tobs) {
            ts.push_back(i); te.push_back(std::min(i + tobs, h->comboStartHost[v + 1])); tc.push_back(v);
        }
    }
    h->fpComboTile0[nCombo] = (int)ts.size();
    h->fpTiles = (int)ts.size();
    h->fpTobs = tobs;
    h->fpOrderItems = 0; h->fpOrderValid = false;
    auto up = [&](int*& d, const std::vector<int>& v) {
        BAM_CUDA(cudaMalloc(&d, sizeof(int) * std::max<size_t>(v.size(), 1)));
        BAM_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, h->stream));
    };
    up(h->d_fpStart, ts); up(h->d_fpEnd, te); up(h->d_fpCombo, tc); up(h->d_fpComboTile0, h->fpComboTile0);
#if BAM_CB_GEN == 2
    {   // the tiles as logpost_baratin_cb2 stages them (one bulk copy each)
        std::vector<long long> off(ts.size() + 1, 0);
        std::vector<int> bytes(ts.size(), 0);
        int mx = 0;
        for (size_t t = 0; t < ts.size(); ++t) {
            const int d = cb2_tile_doubles(te[t] - ts[t]);
            bytes[t] = d * 8;
            off[t + 1] = off[t] + d;
            mx = std::max(mx, d);
        }
        h->fpMaxDoubles = mx;
        BAM_CUDA(cudaMalloc(&h->d_fpPacked, sizeof(double) * std::max<long long>(off.back(), 2)));
        BAM_CUDA(cudaMalloc(&h->d_fpOff, sizeof(long long) * off.size()));
        BAM_CUDA(cudaMalloc(&h->d_fpBytes, sizeof(int) * std::max<size_t>(bytes.size(), 1)));
        BAM_CUDA(cudaMemcpyAsync(h->d_fpOff, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, h->stream));
        BAM_CUDA(cudaMemcpyAsync(h->d_fpBytes, bytes.data(), sizeof(int) * bytes.size(), cudaMemcpyHostToDevice, h->stream));
        if (!h->d_fpSched) {
            BAM_CUDA(cudaMalloc(&h->d_fpSched, 2 * sizeof(int)));
            BAM_CUDA(cudaMemsetAsync(h->d_fpSched, 0, 2 * sizeof(int), h->stream));
        }
        if (!ts.empty())
            cb2_pack_tiles<<<(int)ts.size(), 128, 0, h->stream>>>(h->dp, (int)ts.size(), h->d_fpStart, h->d_fpEnd, h->d_fpCombo, h->d_fpOff,
                                                                  h->d_fpPacked);
        BAM_CUDA(cudaGetLastError());
    }
#endif
    BAM_CUDA(cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
    h->lkCvalid = false;
}

// tile list and affected-combination flags of one vector component (built on first use).  A fitted theta value
// touches the combinations whose parameter set reads it (HostLayout::comboSrc); gamma touches all of them.
static void ensure_comp_tiles(bamgpu_handle* h, int comp) {
    const HostLayout& L = h->L;
    if (h->d_compTiles.empty()) {
        h->d_compTiles.assign(L.P, nullptr); h->d_compAffected.assign(L.P, nullptr); h->compTileCnt.assign(L.P, -1);
    }
    if (h->compTileCnt[comp] >= 0) return;
    std::vector<int> aff(L.nCombo, comp >= L.nFit ? 1 : 0), tiles;
    if (comp < L.nFit)
        for (int k = 0; k < L.nCombo; ++k)
            for (int i = 0; i < L.ntheta; ++i)
                if (L.comboSrc[(size_t)k * L.ntheta + i] == comp) aff[k] = 1;
    for (int k = 0; k < L.nCombo; ++k)
        if (aff[k])
            for (int t = h->fpComboTile0[k]; t < h->fpComboTile0[k + 1]; ++t) tiles.push_back(t);
    BAM_CUDA(cudaMalloc(&h->d_compTiles[comp], sizeof(int) * std::max<size_t>(tiles.size(), 1)));
    BAM_CUDA(cudaMalloc(&h->d_compAffected[comp], sizeof(int) * L.nCombo));
    BAM_CUDA(cudaMemcpyAsync(h->d_compTiles[comp], tiles.data(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice, h->stream));
    BAM_CUDA(cudaMemcpyAsync(h->d_compAffected[comp], aff.data(), sizeof(int) * L.nCombo, cudaMemcpyHostToDevice, h->stream));
    BAM_CUDA(cudaStreamSynchronize(h->stream));
    h->compTileCnt[comp] = (int)tiles.size();
}

void ensure_series_capacity(bamgpu_handle* h, size_t count) {
    if (count <= h->seriesCap) return;
    if (h->d_series) cudaFree(h->d_series);
    BAM_CUDA(cudaMalloc(&h->d_series, sizeof(double) * count));
    h->seriesCap = count;
}

void launch_series(bamgpu_handle* h, const DevProblem& p, int K, const double* tab, int stride, int thOff, int* status, int nOut) {
    series_run<<<(K + 127) / 128, 128, 0, h->stream>>>(p, K, tab, stride, thOff, status, nOut);
    h->launches += 1;
}

cudaEvent_t take_event(bamgpu_handle* h) {
    if (!h->evPool.empty()) { cudaEvent_t e = h->evPool.back(); h->evPool.pop_back(); return e; }
    cudaEvent_t e;
    BAM_CUDA(cudaEventCreate(&e));
    return e;
}

void ensure_chain_capacity(bamgpu_handle* h, int K) {
    if (K <= h->Kcap) { make_plan(h, K); return; }
    auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    fr(h->d_x); fr(h->d_tab); fr(h->d_part); fr(h->d_prior); fr(h->d_lkh); fr(h->d_hyper); fr(h->d_fx); fr(h->d_dpar);
    fr(h->d_status); fr(h->d_delta); fr(h->d_lkCcur); fr(h->d_lkCcand); fr(h->d_finalArrive);
    h->lkCvalid = false;
    const DevProblem& p = h->dp;
    make_plan(h, K);
    BAM_CUDA(cudaMalloc(&h->d_x, sizeof(double) * (size_t)K * p.P));
    BAM_CUDA(cudaMalloc(&h->d_tab, sizeof(double) * (size_t)K * p.tabStride));
    BAM_CUDA(cudaMalloc(&h->d_part, sizeof(double) * 2 * (size_t)K * (std::max(std::max(std::max(h->plan.nTiles, (h->dp.n + 63) / 64), logpost_chain_tiles(h->dp.n)), 1) + p.nCombo)));
    if (p.nCombo <= BAM_FP_MAXCOMBO) {
        BAM_CUDA(cudaMalloc(&h->d_lkCcur, sizeof(double) * (size_t)K * p.nCombo));
        BAM_CUDA(cudaMalloc(&h->d_lkCcand, sizeof(double) * (size_t)K * p.nCombo));
    }
    BAM_CUDA(cudaMalloc(&h->d_prior, sizeof(double) * K));
    BAM_CUDA(cudaMalloc(&h->d_lkh, sizeof(double) * K));
    BAM_CUDA(cudaMalloc(&h->d_hyper, sizeof(double) * K));
    BAM_CUDA(cudaMalloc(&h->d_fx, sizeof(double) * K));
    BAM_CUDA(cudaMalloc(&h->d_dpar, sizeof(double) * (size_t)K * std::max(p.nDpar, 1)));
    BAM_CUDA(cudaMalloc(&h->d_status, sizeof(int) * K));
    BAM_CUDA(cudaMalloc(&h->d_delta, sizeof(double) * K));
    BAM_CUDA(cudaMalloc(&h->d_finalArrive, sizeof(int) * ((K + 31) / 32)));
    BAM_CUDA(cudaMemsetAsync(h->d_finalArrive, 0, sizeof(int) * ((K + 31) / 32), h->stream));
    h->Kcap = K;
}

// one vector through the model on n caller-ordered observations (d_xtrue: n x nX, d_combo: n); returns #infeasible rows
// or -1 when the parameter set itself is infeasible / null
int launch_calibration_run(bamgpu_handle* h, const double* d_x1, int n, const double* d_xtrue, const int* d_combo,
                           double* d_ysim, double* d_state) {
    const DevProblem& p0 = h->dp;
    CalibKernel kern = pick_calib(p0.model, p0.nX, p0.nY);
    if (!kern) throw CudaError{"bamgpu: no kernel instantiated for this (model, nX, nY)"};
    double *tab = nullptr, *prior = nullptr, *dpar = nullptr;
    int *status = nullptr;
    BAM_CUDA(cudaMalloc(&tab, sizeof(double) * std::max(p0.tabStride, 1)));
    BAM_CUDA(cudaMalloc(&prior, sizeof(double)));
    BAM_CUDA(cudaMalloc(&dpar, sizeof(double) * std::max(p0.nDpar, 1)));
    BAM_CUDA(cudaMalloc(&status, 2 * sizeof(int)));
    BAM_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int), h->stream));
    // whole-series models (Vegetation) scan the input series in prep: point the problem image at the caller's series
    DevProblem p = p0;
    p.X = d_xtrue;
    p.n = n;
    launch_prep_on(h->stream, p, 1, d_x1, -1, nullptr, tab, prior, status, dpar);
    double* d_idx = nullptr;
    if (p.isSeries) {  // the series of this one vector (outputs + states), then the lookup kernel on the row indices
        const int nOut = p.nY + model_nstate(p.model);
        ensure_series_capacity(h, (size_t)n * nOut);
        p.seriesX = d_xtrue; p.seriesN = n; p.seriesK = 1; p.seriesY = h->d_series;
        series_run<<<1, 128, 0, h->stream>>>(p, 1, tab, p.tabStride, p.tabCombo, status, nOut);
        std::vector<double> idx((size_t)n * p.nX, 0.0);
        for (int i = 0; i < n; ++i) idx[i] = (double)i;
        BAM_CUDA(cudaMalloc(&d_idx, sizeof(double) * idx.size()));
        BAM_CUDA(cudaMemcpyAsync(d_idx, idx.data(), sizeof(double) * idx.size(), cudaMemcpyHostToDevice, h->stream));
        BAM_CUDA(cudaStreamSynchronize(h->stream));
        d_xtrue = d_idx;
    }
    kern<<<(n + BAM_BLOCK - 1) / BAM_BLOCK, BAM_BLOCK, 0, h->stream>>>(p, n, d_xtrue, d_combo, tab, d_ysim, d_state, status + 1);
    int st[2] = {0, 0};
    BAM_CUDA(cudaMemcpyAsync(st, status, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BAM_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(tab); cudaFree(prior); cudaFree(dpar); cudaFree(status);
    if (d_idx) cudaFree(d_idx);
    h->launches += 2;
    BAM_CUDA(cudaGetLastError());
    return (st[0] & ST_LKH_INFEAS) ? -1 : st[1];
}

// only K2: the per-chain tables (gamma, biases, expanded theta + derived values) of the vectors at d_x
void launch_table(bamgpu_handle* h, int K, const double* d_x) {
    h->launches += launch_prep(h, K, d_x, -1, nullptr);
    BAM_CUDA(cudaGetLastError());
}

void launch_logpost(bamgpu_handle* h, int K, const double* d_x, int comp, const double* d_delta, bool timeMain, bool partial) {
    const DevProblem& p = h->dp;
    const LogpostPlan& pl = h->plan;
    MainKernel kern = pick_main(p.model, p.nX, p.nY);
    if (!kern) throw CudaError{"bamgpu: no kernel instantiated for this (model, nX, nY)"};
    FastKernel fast = (p.n > 0 && !h->forceGeneric) ? pick_fast(p, K) : nullptr;
    cudaEvent_t p0 = nullptr;
    if (timeMain && h->recordKernels && h->prepEv.size() < 8192) { p0 = take_event(h); BAM_CUDA(cudaEventRecord(p0, h->stream)); }
    h->launches += launch_prep(h, K, d_x, comp, d_delta) - 1;  // (the first launch is counted with the step below)
    if (p0) { cudaEvent_t p1 = take_event(h); BAM_CUDA(cudaEventRecord(p1, h->stream)); h->prepEv.emplace_back(p0, p1); }
    DevProblem pser = p;
    if (p.isSeries && p.n > 0) {  // time-recursive model: the outputs of all K vectors, then the generic likelihood kernel
        ensure_series_capacity(h, (size_t)p.seriesN * p.nY * K);
        pser.seriesY = h->d_series;
        pser.seriesK = K;
        series_run<<<(K + 127) / 128, 128, 0, h->stream>>>(pser, K, h->d_tab, p.tabStride, p.tabCombo, h->d_status, p.nY);
        h->launches += 1;
    }
    cudaEvent_t k0 = nullptr, k1 = nullptr;
    if (timeMain) {
        BAM_CUDA(cudaEventRecord(h->ev0, h->stream));
        if (h->recordKernels && h->kernEv.size() < 8192) { k0 = take_event(h); k1 = take_event(h); BAM_CUDA(cudaEventRecord(k0, h->stream)); }
    }
    int nTiles = pl.nTiles;
    bool comboFinal = false;
    const int* d_affected = nullptr;
    if (fast) {
        const int chainBlocks = (K + BAM_BLOCK - 1) / BAM_BLOCK;
        const int slots = BAM_CB2_MINB * h->smCount;  // persistent blocks
#if BAM_CB_GEN == 2
        Fast2Kernel fast2 = h->fpYuOk ? pick_fast2(p, K) : nullptr;
        int tobs = BAM_CB_TOBS;
        if (fast2) tobs = pick_tile_size(h, chainBlocks, slots);
#else
        Fast2Kernel fast2 = nullptr;
        // observations per block: large tiles amortise the staging, but keep >= ~6 blocks per SM in flight
        int tobs = BAM_CB_TOBS;
        while (tobs > 128 && (long long)chainBlocks * ((p.n + tobs - 1) / tobs) < 148 * 6) tobs >>= 1;
#endif
        ensure_fast_tiles(h, tobs);
        nTiles = h->fpTiles;
        comboFinal = p.nCombo <= BAM_FP_MAXCOMBO;
        DevProblem pf = p;
        pf.partPacked = comboFinal ? 1 : 0;  // logpost_final_combo reads one double per (tile, chain)
        // a proposal that touches only some VAR combinations re-evaluates only their tiles (the cached sums of the
        // current state cover the rest); needs the cache to describe the CURRENT state: the sampler guarantees it
        const int* d_tiles = nullptr;
        int launchTiles = nTiles;
        if (partial && comboFinal && comp >= 0 && comp < h->L.P && h->lkCvalid && p.nCombo > 1) {
            ensure_comp_tiles(h, comp);
            if (h->compTileCnt[comp] < nTiles) {
                d_tiles = h->d_compTiles[comp];
                launchTiles = h->compTileCnt[comp];
                d_affected = h->d_compAffected[comp];
            }
        }
        if (fast2) {
            // tables + {2 mbarriers, 2 item ids, 2 arrival counters} + two packed-tile buffers
            auto smem_for = [](int tileDoubles) { return sizeof(double) * (2 * (size_t)BAM_LOG2_N + BAM_EXPTAB_N + 2 * BAM_LOGTAB_N + 4 + 2 * (size_t)tileDoubles); };
            const size_t smem = smem_for(h->fpMaxDoubles);
            if (!h->fp2Attr) {
                // every instantiation that may be picked later for this handle is this one (model options are fixed)
                BAM_CUDA(cudaFuncSetAttribute((const void*)fast2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_for(cb2_tile_doubles(BAM_CB_TOBS + BAM_CB2_ILP))));
                h->fp2Attr = true;
            }
            const int nItems = launchTiles * chainBlocks;
            // profile-guided queue order of the FULL launch (partial re-evaluations are short: natural order)
            const int* d_order = nullptr;
            float* d_cost = nullptr;
            const bool full = d_tiles == nullptr;
            if (full && nItems <= CB2_REORDER_MAX && nItems > 2 * slots) {
                if (h->fpOrderItems != nItems) {
                    auto fr2 = [](auto*& q) { if (q) cudaFree(q); q = nullptr; };
                    fr2(h->d_fpOrder); fr2(h->d_fpCost);
                    BAM_CUDA(cudaMalloc(&h->d_fpOrder, sizeof(int) * nItems));
                    BAM_CUDA(cudaMalloc(&h->d_fpCost, sizeof(float) * nItems));
                    h->fpOrderItems = nItems; h->fpOrderValid = false; h->fpFullLaunches = 0;
                }
                // re-sort after the first launch of a plan, then every 32 launches (chains move slowly in a sampler)
                if (h->fpFullLaunches == 1 || (h->fpFullLaunches > 1 && h->fpFullLaunches % 32 == 0)) {
                    int m = 1;
                    while (m < nItems) m <<= 1;
                    if (m * 8 > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)cb2_reorder, cudaFuncAttributeMaxDynamicSharedMemorySize, m * 8));
                    cb2_reorder<<<1, 1024, (size_t)m * 8, h->stream>>>(h->d_fpCost, nItems, h->d_fpOrder);
                    h->fpOrderValid = true;
                }
                d_order = h->fpOrderValid ? h->d_fpOrder : nullptr;
                d_cost = h->d_fpCost;
                ++h->fpFullLaunches;
            }
            if (nItems > 0)
                fast2<<<std::min(nItems, slots), BAM_BLOCK, smem, h->stream>>>(pf, K, nItems, chainBlocks, h->fpMaxDoubles, h->d_fpPacked, h->d_fpOff,
                                                                            h->d_fpBytes, d_tiles, h->d_tab, h->d_status, h->d_part, h->d_status,
                                                                            h->d_fpSched, d_order, d_cost);
        } else {
            const size_t smem = sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N + 3 * (size_t)tobs);
            dim3 grid(std::max(launchTiles, 1), chainBlocks);
            if (launchTiles > 0)
                fast<<<grid, BAM_BLOCK, smem, h->stream>>>(pf, K, tobs, h->d_fpStart, h->d_fpEnd, h->d_fpCombo, d_tiles, h->d_tab,
                                                           h->d_status, h->d_part, h->d_status);
        }
    } else if (SedLogpostKernel sk = (h->forceGeneric || comp >= p.nFit + p.nGamma) ? nullptr : pick_sed_logpost(p, K)) {
        if (!h->d_sedObs) {  // per-row invariants of the calibration inputs, once per handle
            std::vector<double*> cols(p.nX);
            for (int j = 0; j < p.nX; ++j) cols[j] = const_cast<double*>(p.X) + (size_t)j * p.n;
            double** d_cols = nullptr;
            BAM_CUDA(cudaMalloc(&d_cols, sizeof(double*) * p.nX));
            BAM_CUDA(cudaMemcpyAsync(d_cols, cols.data(), sizeof(double*) * p.nX, cudaMemcpyHostToDevice, h->stream));
            BAM_CUDA(cudaMalloc(&h->d_sedObs, sizeof(SedObs) * (size_t)p.n));
            pick_sed_prep(p.nX)<<<(p.n + 127) / 128, 128, 0, h->stream>>>(p, p.n, 0, p.n, d_cols, static_cast<SedObs*>(h->d_sedObs));
            BAM_CUDA(cudaStreamSynchronize(h->stream));
            cudaFree(d_cols);
            h->launches += 1;
        }
        nTiles = (p.n + SED_TOBS - 1) / SED_TOBS;
        dim3 grid(nTiles, (K + BAM_BLOCK - 1) / BAM_BLOCK);
        sk<<<grid, BAM_BLOCK, 0, h->stream>>>(p, K, static_cast<const SedObs*>(h->d_sedObs), h->d_tab, h->d_status, h->d_part, h->d_status);
    } else if (LatentLogpostKernel lk = pick_latent(h, comp)) {
        nTiles = (p.n + BAM_BLOCK * LL_OPT - 1) / (BAM_BLOCK * LL_OPT);
        dim3 grid((K + LL_CT - 1) / LL_CT, nTiles);
        lk<<<grid, BAM_BLOCK, 0, h->stream>>>(p, d_x, K, h->d_tab, h->d_status, h->d_part, h->d_status);
    } else if (int tt = launch_logpost_tile(h, pser, K, d_x, comp)) {
        nTiles = tt;
    } else if (int ct = launch_logpost_chain(h, pser, K, comp)) {
        nTiles = ct;
    } else if (p.n > 0) {
        if (pl.smem > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        dim3 grid(pl.nTiles, (K + pl.CT - 1) / pl.CT);
        kern<<<grid, BAM_BLOCK, pl.smem, h->stream>>>(pser, d_x, K, pl.CT, pl.TX, h->d_tab, h->d_status, h->d_part, h->d_status,
                                                       comp, d_delta, pl.logOff);
    }
    if (timeMain) {
        BAM_CUDA(cudaEventRecord(h->ev1, h->stream));
        if (k1) { BAM_CUDA(cudaEventRecord(k1, h->stream)); h->kernEv.emplace_back(k0, k1); }
    }
    cudaEvent_t f0 = nullptr;
    if (p0) { f0 = take_event(h); BAM_CUDA(cudaEventRecord(f0, h->stream)); }
    if (comboFinal)
        logpost_final_combo<<<dim3((K + 31) / 32, p.nCombo), 32 * FINAL_STRIPES, 0, h->stream>>>(p, K, h->d_fpComboTile0, h->d_part, h->d_prior,
                                                                                                  h->d_status, d_affected, h->d_lkCcur,
                                                                                                  h->d_lkCcand, h->d_lkh, h->d_hyper, h->d_fx,
                                                                                                  h->d_finalArrive);
    else
        logpost_final<<<(K + 31) / 32, 32 * FINAL_STRIPES, 0, h->stream>>>(p, K, p.n > 0 ? nTiles : 0, h->d_part, h->d_prior,
                                                                           h->d_status, h->d_lkh, h->d_hyper, h->d_fx);
    if (f0) {
        cudaEvent_t f1 = take_event(h);
        BAM_CUDA(cudaEventRecord(f1, h->stream));
        h->finalEv.emplace_back(f0, f1);
    }
    h->launches += (p.n > 0) ? 3 : 2;
    BAM_CUDA(cudaGetLastError());
}

}  // namespace bam

#ifdef BAM_CB2_TRACE
extern "C" int bamgpu_debug_cb2_trace(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, g_cb2Trace, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
extern "C" int bamgpu_debug_cb2_counts(unsigned long long* out, int reset) {
    int e = cudaMemcpyFromSymbol(out, g_cb2Cnt, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : 1;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_cb2Cnt, z, sizeof(z)); }
    return e;
}
#endif
