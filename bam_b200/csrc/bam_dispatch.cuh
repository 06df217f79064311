// bam_dispatch.cuh -- the (model, nX, nY) combinations the templated kernels are instantiated for.
// One list, used by the log-posterior, sampler and prediction translation units.
#pragma once
#include "../../include/bam_gpu.h"

// three groups so that the heaviest kernel (the warp sampler) can be compiled as three translation units in parallel
#define BAM_INSTANCES_A(X)                                                                                  \
    X(BAMGPU_MDL_BARATIN, 1, 1)                                                                             \
    X(BAMGPU_MDL_LINEAR, 1, 1) X(BAMGPU_MDL_LINEAR, 2, 1) X(BAMGPU_MDL_LINEAR, 3, 1) X(BAMGPU_MDL_LINEAR, 4, 1)   \
    X(BAMGPU_MDL_LINEAR, 5, 1) X(BAMGPU_MDL_LINEAR, 6, 1) \
    X(BAMGPU_MDL_LINEAR, 1, 2) X(BAMGPU_MDL_LINEAR, 2, 2) X(BAMGPU_MDL_LINEAR, 3, 2) X(BAMGPU_MDL_LINEAR, 4, 2)   \
    X(BAMGPU_MDL_SEDIMENT, 2, 1) X(BAMGPU_MDL_SEDIMENT, 3, 1) X(BAMGPU_MDL_SEDIMENT, 4, 1)                   \
    X(BAMGPU_MDL_SEDIMENT, 5, 1) X(BAMGPU_MDL_SEDIMENT, 6, 1)
#define BAM_INSTANCES_B(X)                                                                                  \
    X(BAMGPU_MDL_TEXTFILE, 1, 1) X(BAMGPU_MDL_TEXTFILE, 2, 1) X(BAMGPU_MDL_TEXTFILE, 3, 1)                   \
    X(BAMGPU_MDL_TEXTFILE, 4, 1) X(BAMGPU_MDL_TEXTFILE, 1, 2) X(BAMGPU_MDL_TEXTFILE, 2, 2)                   \
    X(BAMGPU_MDL_TEXTFILE, 3, 2) X(BAMGPU_MDL_TEXTFILE, 4, 2)
#define BAM_INSTANCES_C(X)                                                                                  \
    X(BAMGPU_MDL_VEGETATION, 2, 4) X(BAMGPU_MDL_VEGETATION, 2, 5)                                           \
    X(BAMGPU_MDL_VEGETATION, 4, 4) X(BAMGPU_MDL_VEGETATION, 4, 5)                                           \
    X(BAMGPU_MDL_SUSPENDEDLOAD, 1, 1) X(BAMGPU_MDL_SWOT, 2, 1) X(BAMGPU_MDL_SWOT, 3, 1)                     \
    X(BAMGPU_MDL_SGD, 2, 1) X(BAMGPU_MDL_ORTHO, 3, 2) X(BAMGPU_MDL_RECESSION_H, 1, 1)                       \
    X(BAMGPU_MDL_HCS, 1, 1) X(BAMGPU_MDL_SEGMENTATION, 1, 1) X(BAMGPU_MDL_BARATINBAC, 1, 1)                 \
    X(BAMGPU_MDL_SFD, 2, 1)                                                                                 \
    X(BAMGPU_MDL_MIXTURE, 4, 1) X(BAMGPU_MDL_MIXTURE, 5, 1) X(BAMGPU_MDL_MIXTURE, 6, 1)                     \
    X(BAMGPU_MDL_MIXTURE, 7, 1) X(BAMGPU_MDL_MIXTURE, 8, 1)                                                 \
    X(BAMGPU_MDL_SFDTIDAL_QMEC, 2, 1) X(BAMGPU_MDL_SFDTIDAL_QMEC2, 2, 1) X(BAMGPU_MDL_ALGAE, 5, 2) \
    X(BAMGPU_MDL_DYNVEG, 5, 3) X(BAMGPU_MDL_SFDTIDAL_QMEC0, 2, 1) X(BAMGPU_MDL_SFDTIDAL_QMECMS, 2, 1) \
    X(BAMGPU_MDL_SFDTIDAL_LAG, 2, 1) X(BAMGPU_MDL_SFDTIDAL_LAG, 3, 1)
#define BAM_FOR_EACH_INSTANCE(X) BAM_INSTANCES_A(X) BAM_INSTANCES_B(X) BAM_INSTANCES_C(X)
