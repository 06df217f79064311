import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu, load_library
import os
LIB = load_library(os.environ['BAM_LIB']) if os.environ.get('BAM_LIB') else None
Nrep, nobs = 1_000_000, 1024
prob, _ = synth.baratin_2c(nobs=38)
mc, mode, Hs = synth.baratin_spaghetti(Nrep=Nrep, nobs=nobs)
g = BamGpu(prob, lib=LIB)
g.predict_upload(mc, mode, [Hs])
for _ in range(3):
    g.step_begin(); g.predict_resident(1, [1], seed=1); g.step_end()
g.sync()
print(os.environ.get("BAM_LIB","main"), "ms per step", g.step_times(), "kernels", g.kernel_times())
