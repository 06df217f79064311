import sys, os; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu, load_library
LIB = load_library(os.environ['BAM_LIB']) if os.environ.get('BAM_LIB') else None
n, K = 1_000_000, 64
prob, truth = synth.linear_latent(nobs=n)
g = BamGpu(prob, lib=LIB)
s = np.tile(truth, (K, 1))
sd = np.tile(np.concatenate([0.01 * np.abs(truth[:8]), np.full(2 * n, 0.1)]), (K, 1))
g.fit(s, sd, nAdapt=1, nCycles=1, seed=1, want_rows=False)
g.step_times()
g.fit(s, sd, nAdapt=4, nCycles=1, seed=2, want_rows=False)
print(os.environ.get('BAM_LIB', 'main'), "ms per iteration", float(g.step_times()[-1]) / 4)
