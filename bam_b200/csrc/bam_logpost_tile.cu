// bam_logpost_tile.cu -- K1t: generic likelihood + hyper kernel for LATENT inputs on many observations.
//
// logpost_main (bam_logpost.cu) pays, per (chain, observation), a shuffle tree for the likelihood and another for the hyper
// term, a ballot pair, loop and prefetch bookkeeping: 440 thread-instructions per unit around a 5-opcode TextFile formula
// (profiles/r2_textfile_main_ncu.json).  With 10^6 observations per chain (configs[2] and its TextFile twin) the
// organisation of logpost_linear_latent pays for any pointwise model:
//   * block = 256 threads x LT_OPT observations each (stride 256: coalesced) x a tile of LT_CT chains;
//   * the chains' table rows are broadcast reads from shared memory, the 2 LT_CT (likelihood, hyper) sums stay in
//     registers across the thread's observations: ONE block reduction per LT_OPT x LT_CT units;
//   * the chain-independent part of the hyper term (-log sqrt(2 pi) - log Xu, 1 / Xu) comes from the per-handle tables
//     DevProblem::hyC / XuInv instead of a log and a division per (thread, observation).
// Same arithmetic per unit as logpost_main (GetLogLikelihood :3204-3227, GetHyperLogPdf :3285-3303, UnfoldParVector
// :3429-3452) except that hyper terms are summed as -1/2 sum z^2 + sum constants: ~1e-15 relative, not bitwise.
// Eligible (launch_logpost_tile): latent inputs, n >= 16384, no negative Xu, no shifted latent / bias component.
#include <algorithm>

#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_baratin_fast.cuh"

#define LT_CT 4
#define LT_OPT 8
#define LT_BLOCK 256

namespace {

template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(LT_BLOCK) logpost_tile(DevProblem p, const double* __restrict__ x, int K,
                                                         const double* __restrict__ tab, const int* __restrict__ status,
                                                         double* __restrict__ part, int* statusOut) {
    extern __shared__ __align__(16) double sm[];  // [log table | exp table | LT_CT rows | cross-warp partials]
    double2* sLogT = reinterpret_cast<double2*>(sm);
    double* sExpT = sm + 2 * BAM_LOGTAB_N;
    double* rows = sm + 2 * BAM_LOGTAB_N + BAM_EXPTAB_N;
    const int stride = p.tabStride;
    double* red = rows + (size_t)LT_CT * stride;  // [LT_BLOCK / 32][2 LT_CT]
    __shared__ int sst[LT_CT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FM = MODEL == BAMGPU_MDL_VEGETATION || MODEL == BAMGPU_MDL_TEXTFILE;
    for (int q = tid; q < BAM_LOGTAB_N; q += LT_BLOCK) sLogT[q] = p.logTab[q];
    if (MODEL == BAMGPU_MDL_BARATIN || FM)
        for (int q = tid; q < BAM_EXPTAB_N; q += LT_BLOCK) sExpT[q] = p.expTab[q];
    const int c0 = blockIdx.x * LT_CT, nc = min(LT_CT, K - c0);
    for (int idx = tid; idx < LT_CT * stride; idx += LT_BLOCK) {
        const int cc = idx / stride;
        rows[idx] = tab[(size_t)(c0 + min(cc, nc - 1)) * stride + (idx - cc * stride)];
    }
    if (tid < LT_CT) sst[tid] = (tid < nc) ? status[c0 + tid] : 1;
    __syncthreads();
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT);
    const FmTab fm = FM ? FmTab{lt, et} : FmTab{0u, 0u};
    const int n = p.n;
    const double* __restrict__ xrow[LT_CT];
#pragma unroll
    for (int cc = 0; cc < LT_CT; ++cc) xrow[cc] = x + (size_t)(c0 + min(cc, nc - 1)) * p.P;
    bool live[LT_CT];
#pragma unroll
    for (int cc = 0; cc < LT_CT; ++cc) live[cc] = sst[cc] == 0;

    double accL[LT_CT], accH[LT_CT];
#pragma unroll
    for (int cc = 0; cc < LT_CT; ++cc) { accL[cc] = 0.0; accH[cc] = 0.0; }
    double hyConst = 0.0;
    unsigned badMask = 0;
    const int i0 = blockIdx.y * (LT_BLOCK * LT_OPT) + tid;
#pragma unroll 1
    for (int o = 0; o < LT_OPT; ++o) {
        const int i = i0 + o * LT_BLOCK;
        if (i >= n) break;
        const int combo = (p.nCombo > 1) ? p.comboOf[i] : 0;
        double xo[NX], hr[NX], xb[NX], xl[LT_CT][NX];
        int lat[NX], xbi[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const size_t q = i + (size_t)j * n;
            xo[j] = p.X[q];
            hr[j] = p.XuInv[q];
            lat[j] = p.latIdx[q];
            xb[j] = 0.0; xbi[j] = 0;
            if (p.hasXbias) { xb[j] = p.Xb[q]; xbi[j] = p.Xbindx[q]; }
        }
        // the LT_CT x NX chain-private cells of this observation: independent loads, requested together
#pragma unroll
        for (int cc = 0; cc < LT_CT; ++cc)
#pragma unroll
            for (int j = 0; j < NX; ++j) xl[cc][j] = (lat[j] >= 0) ? __ldcs(xrow[cc] + lat[j]) : 0.0;
        double yo[NY], yu[NY], yb[NY];
        int ybi[NY];
#pragma unroll
        for (int j = 0; j < NY; ++j) {
            const size_t q = i + (size_t)j * n;
            yo[j] = p.Y[q];
            yu[j] = p.Yu[q];
            yb[j] = 0.0; ybi[j] = 0;
            if (p.hasYbias) { yb[j] = p.Yb[q]; ybi[j] = p.Ybindx[q]; }
        }
        hyConst += p.hyC[i];
        // UnfoldParVector for the LT_CT chains: latent cell from the vector, else un-bias the observed input
        double xt[LT_CT][NX], yhAll[LT_CT][NY];
        unsigned okMask = 0u;
#pragma unroll
        for (int cc = 0; cc < LT_CT; ++cc) {
            const double* __restrict__ row = rows + (size_t)cc * stride;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                double v = xo[j];
                if (lat[j] >= 0) v = xl[cc][j];
                else if (p.hasXbias && xbi[j] > 0) v = xo[j] - xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
                xt[cc][j] = v;
            }
        }
        if (MODEL == BAMGPU_MDL_TEXTFILE) {  // one pass of the bytecode interpreter serves the LT_CT chains
            const double* thp[LT_CT];
#pragma unroll
            for (int cc = 0; cc < LT_CT; ++cc) thp[cc] = rows + (size_t)cc * stride + p.tabCombo + (size_t)combo * p.comboStride;
            unsigned fail = 0u;
#pragma unroll
            for (int k = 0; k < NY; ++k) {
                double r[LT_CT];
                fail |= textfile_eval_multi<NX, LT_CT>(p, k, xt, thp, r, fm);
#pragma unroll
                for (int cc = 0; cc < LT_CT; ++cc) yhAll[cc][k] = r[cc];
            }
            okMask = ~fail;
        } else {
#pragma unroll
            for (int cc = 0; cc < LT_CT; ++cc) {
                if (!live[cc]) continue;  // uniform over the block
                const double* __restrict__ th = rows + (size_t)cc * stride + p.tabCombo + (size_t)combo * p.comboStride;
                const bool ok = (MODEL == BAMGPU_MDL_BARATIN) ? baratin_apply_t(p, xt[cc][0], th, th + p.ntheta, yhAll[cc][0], lt, et)
                                                              : model_apply<MODEL, NX, NY>(p, xt[cc], th, th + p.ntheta, yhAll[cc], fm);
                if (ok) okMask |= 1u << cc;
            }
        }
#pragma unroll
        for (int cc = 0; cc < LT_CT; ++cc) {
            if (!live[cc]) continue;
            if (!((okMask >> cc) & 1u)) { badMask |= 1u << cc; continue; }
            const double* __restrict__ row = rows + (size_t)cc * stride;
            const double* yh = yhAll[cc];
            double l = 0.0;
#pragma unroll
            for (int j = 0; j < NY; ++j) {
                if (p.hasMissing && yo[j] == p.mv) continue;
                const double sig = sigmafunk(p.sigfunc[j], row + p.tabGamma + p.gamOff[j], yh[j]);
                if (!(sig > 0.0)) { badMask |= 1u << cc; continue; }
                const double var = fma(yu[j], yu[j], sig * sig);
                double mu = yh[j];
                if (p.hasYbias && ybi[j] != 0) mu = yh[j] + yb[j] * row[p.tabOffYb[j] + ybi[j] - 1];
                l += gauss_logpdf_var_t(yo[j], mu, var, lt);
            }
            accL[cc] += l;
            double H = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {  // cells without input error have hr = 0: no term
                double mu = xt[cc][j];
                if (p.hasXbias && xbi[j] != 0) mu = xt[cc][j] + xb[j] * row[p.tabOffXb[j] + xbi[j] - 1];
                const double z = (xo[j] - mu) * hr[j];
                H = fma(z, z, H);
            }
            accH[cc] += H;
        }
    }
    // thread sums -> log-densities, then the block reduction (fixed tree, fixed order)
#pragma unroll
    for (int cc = 0; cc < LT_CT; ++cc) {
        double a = live[cc] ? accL[cc] : 0.0;
        double b = live[cc] ? fma(-0.5, accH[cc], hyConst) : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
        }
        if (lane == 0) { red[warp * 2 * LT_CT + 2 * cc] = a; red[warp * 2 * LT_CT + 2 * cc + 1] = b; }
    }
    const unsigned anyBad = __reduce_or_sync(0xffffffffu, badMask);
    if (lane == 0 && anyBad)
        for (int cc = 0; cc < nc; ++cc)
            if ((anyBad >> cc) & 1u) atomicOr(&statusOut[c0 + cc], ST_LKH_INFEAS);
    __syncthreads();
    if (tid < 2 * LT_CT) {
        const int cc = tid >> 1;
        if (cc < nc) {
            double a = 0.0;
            for (int w = 0; w < LT_BLOCK / 32; ++w) a += red[w * 2 * LT_CT + tid];
            part[((size_t)blockIdx.y * K + (c0 + cc)) * 2 + (tid & 1)] = a;
        }
    }
}

using TileKernel = void (*)(DevProblem, const double*, int, const double*, const int*, double*, int*);
TileKernel pick_tile(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return logpost_tile<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

}  // namespace

namespace bam {

// returns the number of observation tiles written to h->d_part, or 0 when the launch is not eligible (caller falls back)
int launch_logpost_tile(bamgpu_handle* h, const DevProblem& p, int K, const double* d_x, int comp) {
    if (h->forceGeneric || !p.hasLatent || p.hyperInfeasible || !p.XuInv || !p.hyC || p.n < 16384 || p.isSeries) return 0;
    if (comp >= p.nFit + p.nGamma) return 0;  // a shifted latent cell / bias is applied by logpost_main
    const size_t smem = sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N + (size_t)LT_CT * p.tabStride + 2 * LT_CT * (LT_BLOCK / 32));
    if (smem > 64 * 1024) return 0;
    TileKernel kern = pick_tile(p.model, p.nX, p.nY);
    if (!kern) return 0;
    if (smem > 48 * 1024 && h->tileAttr < smem) {
        BAM_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->tileAttr = smem;
    }
    const int nTiles = (p.n + LT_BLOCK * LT_OPT - 1) / (LT_BLOCK * LT_OPT);
    dim3 grid((K + LT_CT - 1) / LT_CT, nTiles);
    kern<<<grid, LT_BLOCK, smem, h->stream>>>(p, d_x, K, h->d_tab, h->d_status, h->d_part, h->d_status);
    return nTiles;
}

}  // namespace bam
