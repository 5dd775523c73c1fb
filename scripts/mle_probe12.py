"""The grid fit (carma_mle_grid_device) against the number of starts per model and the blocks per SM of the launch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
model = C.CarmaModel(t, y, e)
pqlist = [(p, q) for p in range(1, 8) for q in range(p)]
print({k: v for k, v in os.environ.items() if k.startswith("CARMA_MLE")})
for ntr in (100, 300):
    jobs = []
    for k, (p, q) in enumerate(pqlist):
        j = model.mle_starts(p, q, ntr, seed=500 + k)
        jobs.append((j[0], p, q) + tuple(j[1:]))
    model.series.mle_grid(jobs[:3], maxiter=3)
    for rep in range(2):
        t0 = time.perf_counter()
        res = model.series.mle_grid(jobs)
        print("starts per model", ntr, "rep", rep, "wall_s", round(time.perf_counter() - t0, 3), "best (7,6)", round(float(res[-1][1].min()), 4), flush=True)
