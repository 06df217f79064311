// bam_sampler_kernels.h -- the two big sampler kernels live in their own translation units (they are instantiated for
// every (model, nX, nY) and dominate the build time); bam_sampler.cu only needs their launch signatures.
#pragma once
#include "bam_device.cuh"

#define WS_WARPS 4       /* chains per block */
#define WS_MAXP 64       /* vector length limit of this path (no latent inputs) */
#define WS_TABLES (2 * BAM_LOGTAB_N + BAM_EXPTAB_N) /* doubles of shared memory in front of the per-warp areas: log / exp tables */

struct WarpSamplerArgs {
    int K, nIterChunk, nAdapt;
    long long it0, burn, keepEvery, nKeep;
    double minMR, maxMR, downMult, upMult;
    unsigned long long seed;
    long long chainOffset;
    double *x, *sd, *fxcur, *dparcur, *rows, *mcount, *mmean, *mm2;
    int* moves;
};

struct bamgpu_handle;

namespace bam {
using WarpSamplerKernel = void (*)(DevProblem, WarpSamplerArgs);
WarpSamplerKernel pick_warp_sampler(int model, int nX, int nY);  // bam_sampler_warp.cu
using LatentKernel = void (*)(DevProblem, int, unsigned long long, long long, unsigned, double*, const double*, int*,
                              const double*);
LatentKernel pick_latent_sweep(int model, int nX, int nY);       // bam_latent.cu
LatentKernel pick_latent_sweep_linear(const bamgpu_handle* h);   // bam_latent.cu: cfg-3 fast path or nullptr
}  // namespace bam
