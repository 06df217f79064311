// bam_logpost_chain.cu -- K1c: the generic likelihood kernel in CHAIN-PARALLEL layout (thread <-> chain).
//
// logpost_main (bam_logpost.cu) gives a thread one observation and walks the chains of a tile: right for few chains on
// many observations, but at K >= 1024 chains it pays a shuffle tree per (chain, observation), and a model whose
// parameters are selected by the observation (the per-year parameters of Vegetation, VAR combinations) reads its
// shared-memory row at 32 different places per warp.  Here, as in the prediction kernel (pred_main, bam_predict.cu):
//   * block = 128 chains; every chain's table row [gamma | biases | theta, derived per combination] sits in shared
//     memory at an ODD stride, so "entry k of my row" is conflict-free across the lanes;
//   * the block walks a tile of observations; X, Y, Yu, combination and bias indices are the same address for all
//     lanes (one broadcast load each, L1 / L2 resident);
//   * the thread accumulates its chain's log-likelihood in a register, in observation order: no shuffles, no barriers
//     after the staging; one partial per (tile, chain), summed over tiles by logpost_final in fixed order.
// Same arithmetic per (chain, observation) as logpost_main: GetLogLikelihood, src/BaM_tools.f90:3204-3227 with
// UnfoldParVector's un-biasing of the inputs (:3429-3452).  The tiling depends on n only, so a chain's value does not
// depend on how many chains share the launch.
// Eligible (launch_logpost): K >= 1024, no latent inputs (their hyper term and per-chain cells stay with logpost_main),
// rows fit in shared memory.  Proposals of the launch-per-proposal sampler (one shifted component) take it as well.
#include <algorithm>

#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_baratin_fast.cuh"

#define LC_BLOCK 128
#define LC_TILES 32   /* observation tiles (upper bound): n only, never K */
#define LC_MINOBS 8

namespace {

template <int MODEL, int NX, int NY>
__global__ void __launch_bounds__(LC_BLOCK) logpost_chain(DevProblem p, int K, int tobs, const double* __restrict__ tab,
                                                          const int* __restrict__ status, double* __restrict__ part,
                                                          int* statusOut) {
    extern __shared__ __align__(16) double sm[];  // [log table | exp table | LC_BLOCK rows of (tabStride | 1) doubles]
    double2* sLogT = reinterpret_cast<double2*>(sm);
    double* sExpT = sm + 2 * BAM_LOGTAB_N;
    double* rows = sm + 2 * BAM_LOGTAB_N + BAM_EXPTAB_N;
    const int tid = threadIdx.x;
    constexpr bool FM = MODEL == BAMGPU_MDL_VEGETATION || MODEL == BAMGPU_MDL_TEXTFILE;  // table-driven log / exp inside the model
    for (int q = tid; q < BAM_LOGTAB_N; q += LC_BLOCK) sLogT[q] = p.logTab[q];
    if (MODEL == BAMGPU_MDL_BARATIN || FM)
        for (int q = tid; q < BAM_EXPTAB_N; q += LC_BLOCK) sExpT[q] = p.expTab[q];
    const int stride = p.tabStride, strideP = stride | 1;
    const int c0 = blockIdx.x * LC_BLOCK, nc = min(LC_BLOCK, K - c0);
    // coalesced staging: consecutive threads read consecutive doubles of the block's nc contiguous rows
    for (int idx = tid; idx < nc * stride; idx += LC_BLOCK) {
        const int cc = idx / stride, k = idx - cc * stride;
        rows[(size_t)cc * strideP + k] = tab[(size_t)c0 * stride + idx];
    }
    __syncthreads();
    if (tid >= nc) return;
    const int c = c0 + tid;
    const unsigned lt = smem_addr(sLogT), et = smem_addr(sExpT);
    const FmTab fm = FM ? FmTab{lt, et} : FmTab{0u, 0u};
    const double* __restrict__ row = rows + (size_t)tid * strideP;
    const int n = p.n;
    const int iBeg = blockIdx.y * tobs, iEnd = min(n, iBeg + tobs);
    double lk = 0.0;
    bool bad = false;
    if (status[c] == 0) {
        for (int i = iBeg; i < iEnd; ++i) {
            const int combo = (p.nCombo > 1) ? p.comboOf[i] : 0;
            double xt[NX], yh[NY];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const size_t q = i + (size_t)j * n;
                double v = p.X[q];
                if (p.hasXbias) {
                    const int bi = p.Xbindx[q];
                    if (bi > 0) v = v - p.Xb[q] * row[p.tabOffXb[j] + bi - 1];
                }
                xt[j] = v;
            }
            const double* __restrict__ th = row + p.tabCombo + (size_t)combo * p.comboStride;
            const bool ok = (MODEL == BAMGPU_MDL_BARATIN) ? baratin_apply_t(p, xt[0], th, th + p.ntheta, yh[0], lt, et)
                                                          : model_apply<MODEL, NX, NY>(p, xt, th, th + p.ntheta, yh, fm);
            if (!ok) { bad = true; continue; }
#pragma unroll
            for (int j = 0; j < NY; ++j) {
                const size_t q = i + (size_t)j * n;
                const double yo = p.Y[q];
                if (p.hasMissing && yo == p.mv) continue;
                const double sig = sigmafunk(p.sigfunc[j], row + p.tabGamma + p.gamOff[j], yh[j]);
                if (!(sig > 0.0)) { bad = true; continue; }
                const double yu = p.Yu[q];
                const double var = fma(yu, yu, sig * sig);
                double mu = yh[j];
                if (p.hasYbias) {
                    const int bi = p.Ybindx[q];
                    if (bi != 0) mu = yh[j] + p.Yb[q] * row[p.tabOffYb[j] + bi - 1];
                }
                lk += gauss_logpdf_var_t(yo, mu, var, lt);
            }
        }
    }
    const size_t o = ((size_t)blockIdx.y * K + c) * 2;
    part[o] = lk;
    part[o + 1] = 0.0;
    if (bad) atomicOr(&statusOut[c], ST_LKH_INFEAS);
}

using ChainKernel = void (*)(DevProblem, int, int, const double*, const int*, double*, int*);
ChainKernel pick_chain(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return logpost_chain<M, A, B>;
    BAM_FOR_EACH_INSTANCE(X)
#undef X
    return nullptr;
}

}  // namespace

namespace bam {

int logpost_chain_tiles(int n) {
    const int tobs = std::max(LC_MINOBS, (n + LC_TILES - 1) / LC_TILES);
    return (n + tobs - 1) / tobs;
}

// returns the number of observation tiles written to h->d_part, or 0 when the launch is not eligible (caller falls back)
int launch_logpost_chain(bamgpu_handle* h, const DevProblem& p, int K, int comp) {
    // (a single-component shift `comp` needs nothing here: without latent inputs every component is a theta, a gamma or a
    // bias, and logpost_prep has already written the shifted value into the chain's table row)
    (void)comp;
    if (h->forceGeneric || K < 1024 || p.n < 1 || p.hasLatent) return 0;
    const size_t smem = sizeof(double) * (2 * BAM_LOGTAB_N + BAM_EXPTAB_N + (size_t)LC_BLOCK * (p.tabStride | 1));
    if (smem > 200 * 1024) return 0;
    ChainKernel kern = pick_chain(p.model, p.nX, p.nY);
    if (!kern) return 0;
    if (smem > 48 * 1024 && h->chainAttr < smem) {
        BAM_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->chainAttr = smem;
    }
    const int tobs = std::max(LC_MINOBS, (p.n + LC_TILES - 1) / LC_TILES);
    const int nTiles = (p.n + tobs - 1) / tobs;
    dim3 grid((K + LC_BLOCK - 1) / LC_BLOCK, nTiles);
    kern<<<grid, LC_BLOCK, smem, h->stream>>>(p, K, tobs, h->d_tab, h->d_status, h->d_part, h->d_status);
    return nTiles;
}

}  // namespace bam
