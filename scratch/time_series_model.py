import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14401
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prob, truth = synth.sfdtidal_qmec(nobs=n)
x = synth.feasible_starts(truth, K, rel=0.005, seed=2)
g = BamGpu(prob)
g.chains_upload(x)
for _ in range(2): g.logpost_resident()
g.sync(); g.step_times()
for _ in range(5):
    g.step_begin(); g.logpost_resident(); g.step_end()
g.sync()
ms = g.step_times()
fx, feas, _ = g.chains_download()
print(f"SFDTidal_Qmec n={n} K={K}: {ms.mean():.3f} ms per batch -> {K*n/ms.mean()/1e-3:.3e} chain*obs/s; feasible {feas.mean():.2f}")
