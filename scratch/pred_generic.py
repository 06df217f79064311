import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
Nrep = 1_000_000
for model, kw in (("SWOT", dict(variant="SWOT_hSB")), ("SuspendedLoad", {})):
    prob, truth = synth.second_wave(model, nobs=1024, **kw)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((Nrep, truth.size)))
    X = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    import os
    from bam_b200.capi import load_library
    g = BamGpu(prob, lib=(load_library(os.environ["BAM_LIB"]) if os.environ.get("BAM_LIB") else None))
    g.predict_upload(mc, truth, X)
    g.predict_resident(1, [1], seed=1); g.sync(); g.step_times(); g.kernel_times()
    g.step_begin(); g.predict_resident(1, [1], seed=1); g.step_end(); g.sync()
    ms = g.step_times(); km = g.kernel_times()
    print(f"{model}: 1e6 x 1024 with noise: step {ms[0]:.2f} ms, propagation {km[0::2].sum():.2f} ms, envelopes {km[1::2].sum():.2f} ms")
