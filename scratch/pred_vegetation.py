import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
Nrep = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 1095
growth = sys.argv[3] if len(sys.argv) > 3 else "Growth_Yin3"
prob, truth = synth.vegetation(nobs=nobs, growth=growth, nY=5)
rng = np.random.default_rng(3)
mc = truth * (1.0 + 0.002 * rng.standard_normal((Nrep, truth.size)))
X = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
g = BamGpu(prob)
g.predict_upload(mc, truth, X)
dor = [1, 0, 0, 0, 0]
g.predict_resident(1, dor, seed=1); g.sync(); g.step_times(); g.kernel_times()
g.step_begin(); g.predict_resident(1, dor, seed=1); g.step_end(); g.sync()
ms = g.step_times(); km = g.kernel_times()
env = g.predict_download()
print(f"Vegetation {growth} Nrep={Nrep} x {nobs} inputs x 5 outputs: {ms[0]:.1f} ms -> {Nrep*nobs/ms[0]/1e-3:.3e} sample*input/s; propagation {km[0::2].sum():.1f} ms, envelopes {km[1::2].sum():.1f} ms; finite {np.isfinite(env).all()}")
