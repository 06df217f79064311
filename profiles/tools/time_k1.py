import sys, os; sys.path.insert(0, '.')
import numpy as np
from workloads import synth
from bam_b200.capi import BamGpu, load_library
LIB = load_library(os.environ['BAM_LIB']) if os.environ.get('BAM_LIB') else None
Ks = [int(k) for k in os.environ.get('BAM_K', '4096').split(',')]
prob, truth = synth.baratin_spd(nobs=100_000, copula=True)
x = synth.feasible_starts(truth, max(Ks), check=synth.spd_start_check)
for K in Ks:
    g = BamGpu(prob, lib=LIB)
    g.chains_upload(x[:K])
    for _ in range(3): g.logpost_resident()
    g.sync(); g.step_times(); g.kernel_times()
    for _ in range(20):
        g.flush_l2(); g.step_begin(); g.logpost_resident(); g.step_end()
    g.sync()
    ms = g.step_times(); km = g.kernel_times()
    print(os.environ.get('BAM_LIB', 'main'), f"K {K} step {ms.mean():.4f} ms (min {ms.min():.4f}) kernel {km.mean():.4f} ms (min {km.min():.4f})", flush=True)
