"""One ensemble (10 chains = 10 threads) for a few iterations: the latency-bound extreme, for ncu."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(270, 270)
s = C.Series(t, y, e)
for _ in range(2):
    t0 = time.perf_counter()
    r = s.pt_run(C.KIND_CARMA, 5, 3, 100, 100, ntemps=10, n_ensembles=1, seed=3, init=synth.readme_theta(3))
    print("200 iterations: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
