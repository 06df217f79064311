import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
n, K = 1_000_000, 64
prob, truth = synth.linear_latent(nobs=n)
g = BamGpu(prob)
s = np.tile(truth, (K, 1))
sd = np.tile(np.concatenate([0.01 * np.abs(truth[:8]), np.full(2 * n, 0.1)]), (K, 1))
g.fit(s, sd, nAdapt=3, nCycles=1, seed=1, want_rows=False)
print("ok", g.step_times())
