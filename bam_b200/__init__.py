"""bam_b200 -- B200-native (sm_100a) implementation of BaM's data-parallel hot path.

The product is the C-ABI CUDA library ``bam_b200/libbam_gpu.so`` (sources in ``bam_b200/csrc``,
interface in ``include/bam_gpu.h``) and the C++ host driver ``bam_b200/bam_gpu``.  This Python
package is only the ctypes binding used by the tests and by ``bench.py``; it contains no
compute path and no CPU fallback: every call fails loudly if the CUDA library is missing.
"""
from .capi import (  # noqa: F401
    BamGpu,
    Problem,
    Sizes,
    load_library,
    library_path,
    DIST,
    SIGMAFUNC,
    PARTYPE,
)

__all__ = ["BamGpu", "Problem", "Sizes", "load_library", "library_path", "DIST", "SIGMAFUNC", "PARTYPE"]
