"""choose_order(7) x 100 with the on-device optimiser, three times in one process, with a per-model timeline
(start of the fit, end of the starting-value phase, end of the fit) of the last repetition."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
print({k: v for k, v in os.environ.items() if k.startswith("CARMA_") or k.startswith("CUDA_MODULE")})
model = C.CarmaModel(t, y, e)
model.mle_optimizer = sys.argv[1] if len(sys.argv) > 1 else "device"
log, T0 = [], [0.0]
orig_starts = model.mle_starts
def starts(p, q, *a, **k):
    t0 = time.perf_counter() - T0[0]
    r = orig_starts(p, q, *a, **k)
    log.append([p, q, round(t0, 3), round(time.perf_counter() - T0[0], 3)])
    return r
model.mle_starts = starts
for rep in range(3):
    log.clear()
    T0[0] = time.perf_counter()
    mle, pq, aicc = model.choose_order(7, ntrials=100, seed=500, verbose=False)
    print("rep", rep, "wall_s", round(time.perf_counter() - T0[0], 3), flush=True)
print("starting-value phases (p, q, begin, end):")
for row in sorted(log, key=lambda r: r[2]):
    print("  ", row)
