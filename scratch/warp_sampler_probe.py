import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200.capi import BamGpu
from tests.helpers import problem_from_fixture
prob, d = problem_from_fixture("baratin")
K = 4096
g = BamGpu(prob)
s = np.tile(np.asarray(d["start"]), (K, 1)); sd = 0.1*np.abs(s)
g.fit(s, sd, nAdapt=10, nCycles=1, want_rows=False)
g.step_times()
g.fit(s, sd, nAdapt=100, nCycles=3, seed=3, want_rows=False)
print("ms", g.step_times())
