import sys, time; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prob, truth = synth.linear_latent(nobs=n)
g = BamGpu(prob)
P = g.P
s = np.tile(truth, (K, 1))
sd = np.concatenate([0.01 * np.abs(truth[:8]), np.full(2 * n, 0.1)])
sdev = np.tile(sd, (K, 1))
g.fit(s, sdev, nAdapt=1, nCycles=1, seed=1, want_rows=False)
for nA in (4, 8):
    t0 = time.perf_counter()
    g.fit(s, sdev, nAdapt=nA, nCycles=1, seed=2, want_rows=False)
    dt = time.perf_counter() - t0
    print(f"n={n} K={K} P={P}: {nA} iterations in {dt:.3f} s (includes {2*K*P*8/1e9:.1f} GB host upload)")
fx, feas, isnull = g.chains_download()
print("feasible", feas.all(), "fx", fx[:3])
cnt, mean, m2 = g.moments()
print("moves ok", np.isfinite(mean).all())
