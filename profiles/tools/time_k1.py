import sys, os; sys.path.insert(0, '.')
import numpy as np
from workloads import synth
from bam_b200.capi import BamGpu, load_library
LIB = load_library(os.environ['BAM_LIB']) if os.environ.get('BAM_LIB') else None
Ks = [int(k) for k in os.environ.get('BAM_K', '4096').split(',')]
gens = [int(k) for k in os.environ.get('BAM_GEN', '4').split(',')]
prob, truth = synth.baratin_spd(nobs=100_000, copula=True)
x = synth.feasible_starts(truth, max(Ks), check=synth.spd_start_check)
ref = {}
for gen in gens:
  for K in Ks:
    g = BamGpu(prob, lib=LIB)
    g.chains_upload(x[:K])
    for _ in range(3): g.logpost_resident()
    g.sync(); g.step_times(); g.kernel_times(); g.phase_times()
    for _ in range(20):
        g.flush_l2(); g.step_begin(); g.logpost_resident(); g.step_end()
    g.sync()
    ms = g.step_times(); km = g.kernel_times(); pm, fm = g.phase_times()
    fx = g.chains_download()[0]
    d = ""
    if K in ref: d = " max rel diff vs first gen %.2e" % float(np.max(np.abs(fx - ref[K]) / np.abs(ref[K])))
    else: ref[K] = fx
    print(f"gen {gen} K {K} step {ms.mean():.4f} ms (min {ms.min():.4f}) kernel {km.mean():.4f} ms (min {km.min():.4f}) prep {pm.mean():.4f} final {fm.mean():.4f}{d}", flush=True)
