import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu, load_library
import os
LIB = load_library(os.environ['BAM_LIB']) if os.environ.get('BAM_LIB') else None
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prob, truth = synth.linear_latent(nobs=n)
rng = np.random.default_rng(1)
x = np.tile(truth, (K, 1))
x[:, :8] *= 1.0 + 0.01 * rng.standard_normal((K, 8))
x[:, 8:] += 0.05 * rng.standard_normal((K, 2 * n))
res = {}
for mode in (("fast",) if os.environ.get("BAM_LIB") else ("generic", "fast")):
    g = BamGpu(prob, lib=LIB)
    g.set_option("force_generic", 1 if mode == "generic" else 0)
    g.chains_upload(x)
    for _ in range(3):
        g.logpost_resident()
    g.sync(); g.step_times(); g.kernel_times()
    for _ in range(10):
        g.flush_l2()
        g.step_begin(); g.logpost_resident(); g.step_end()
    g.sync()
    ms = g.step_times(); km = g.kernel_times()
    fx, feas, isnull = g.chains_download()
    res[mode] = fx
    print(f"{mode}: n={n} K={K} step {ms.mean():.3f} ms kernel {km.mean():.3f} ms -> {K*n/km.mean()/1e-3:.3e} chain*obs/s, "
          f"{16.0*K*n/km.mean()/1e-3/1e9:.0f} GB/s of latent inputs; feas {feas.all()}")
    del g
if "generic" in res: print("max rel diff fast vs generic", np.max(np.abs(res["fast"] - res["generic"]) / np.abs(res["generic"])))
