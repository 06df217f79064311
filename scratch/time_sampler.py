import sys, time; sys.path.insert(0, '.')
import numpy as np
from bam_b200.capi import BamGpu
from tests.helpers import problem_from_fixture
prob, d = problem_from_fixture("baratin")
for K, nA, nC, generic in ((4096, 100, 20, 0), (4096, 100, 2, 1), (1, 100, 100, 0)):
    g = BamGpu(prob)
    g.set_option("force_generic", generic)
    s = np.tile(np.asarray(d["start"]), (K, 1)); sd = 0.1*np.abs(s)
    g.fit(s, sd, nAdapt=10, nCycles=1, want_rows=False)
    t0 = time.perf_counter()
    rows = g.fit(s, sd, nAdapt=nA, nCycles=nC, seed=3, want_rows=False)
    dt = time.perf_counter() - t0
    nprop = K * nA * nC * g.P
    print(f"K={K} iters={nA*nC} generic={generic}: {dt:.3f}s  {nprop/dt:.3e} proposals/s  {nprop*38/dt:.3e} chain*obs/s, launches {g.launch_count()}")
    cnt, mean, m2 = g.moments()
    if K > 1: print("   rhat", np.round(g.rhat(cnt, mean, m2), 3))
