import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
K, n = 4096, 100_000
prob, truth = synth.baratin_spd(nobs=n)
s = synth.feasible_starts(truth, K, check=synth.spd_start_check)
sd = 0.002 * np.abs(s) + 1e-6
for nopart in (1, 0):
    g = BamGpu(prob)
    g.set_option("no_partial", nopart)
    g.fit(s, sd, nAdapt=1, nCycles=1, seed=1, want_rows=False)
    g.step_times()
    g.fit(s, sd, nAdapt=5, nCycles=1, seed=2, want_rows=False)
    ms = g.step_times()[-1] / 5
    print(f"cfg 2 sampler, {K} chains x {n} obs, P={g.P}: no_partial={nopart}: {ms:.2f} ms per iteration ({K*g.P/(ms*1e-3):.3e} proposals/s, {K*g.P*n/(ms*1e-3):.3e} chain*obs/s equivalent)")
