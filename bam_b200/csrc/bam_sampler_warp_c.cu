// bam_sampler_warp_c.cu -- warp sampler instances, group C of bam_dispatch.cuh
#include "bam_sampler_warp.cuh"

namespace bam {

WarpSamplerKernel pick_warp_sampler_c(int model, int nX, int nY) {
#define X(M, A, B) \
    if (model == M && nX == A && nY == B) return sampler_warp<M, A, B>;
    BAM_INSTANCES_C(X)
#undef X
    return nullptr;
}
}  // namespace bam
