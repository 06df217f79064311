// bam_sampler_warp_a.cu -- warp sampler instances, group A of bam_dispatch.cuh
#include "bam_sampler_warp.cuh"

namespace bam {
WarpSamplerKernel pick_warp_sampler_b(int model, int nX, int nY);
WarpSamplerKernel pick_warp_sampler(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return sampler_warp<M, A, B>;
    BAM_INSTANCES_A(X)
#undef X
    return pick_warp_sampler_b(model, nX, nY);
}
}  // namespace bam
