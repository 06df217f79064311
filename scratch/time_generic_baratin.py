import sys, os; sys.path.insert(0, '.')
import numpy as np
from bam_b200 import synth
from bam_b200.capi import BamGpu
K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prob, truth = synth.baratin_spd(nobs=100_000, copula=True)
x = synth.feasible_starts(truth, K, check=synth.spd_start_check)
from bam_b200.capi import load_library
g = BamGpu(prob, lib=(load_library(os.environ["BAM_LIB"]) if os.environ.get("BAM_LIB") else None))
g.chains_upload(x)
for _ in range(3): g.logpost_resident()
g.sync(); g.step_times(); g.kernel_times()
for _ in range(10):
    g.flush_l2(); g.step_begin(); g.logpost_resident(); g.step_end()
g.sync()
ms = g.step_times(); km = g.kernel_times()
print(f"generic BaRatin_SPD K={K} x 100000: step {ms.mean():.4f} ms kernel {km.mean():.4f} ms -> {K*1e5/km.mean()/1e-3:.3e} chain*obs/s")
