// bam_sampler_warp_b.cu -- warp sampler instances, group B of bam_dispatch.cuh
#include "bam_sampler_warp.cuh"

namespace bam {
WarpSamplerKernel pick_warp_sampler_c(int model, int nX, int nY);
WarpSamplerKernel pick_warp_sampler_b(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return sampler_warp<M, A, B>;
    BAM_INSTANCES_B(X)
#undef X
    return pick_warp_sampler_c(model, nX, nY);
}
}  // namespace bam
