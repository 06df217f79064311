import sys, time; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
Nrep = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 812
prob, truth = synth.sediment_camenen(nobs=nobs)
mc = synth.sediment_posterior(Nrep)
X = [np.asarray(prob.X)[:, j].copy() for j in range(4)]
g = BamGpu(prob)
g.predict_upload(mc, truth, X)
for dor in ([0], [1]):
    g.predict_resident(1, dor, seed=1); g.sync(); g.step_times(); g.kernel_times()
    g.step_begin(); g.predict_resident(1, dor, seed=1); g.step_end(); g.sync()
    ms = g.step_times(); km = g.kernel_times()
    env = g.predict_download()
    print(f"Sediment Camenen+Soulsby Nrep={Nrep} x {nobs} inputs, remnant={dor[0]}: {ms[0]:.1f} ms -> {Nrep*nobs/ms[0]/1e-3:.3e} pairs/s; propagation {km[0::2].sum():.1f} ms, envelopes {km[1::2].sum():.1f} ms; finite {np.isfinite(env).all()}")
