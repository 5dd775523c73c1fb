"""choose_order(7) x 100 starts, cold and repeated in the same process, host-loop vs on-device optimiser."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
print("CUDA_MODULE_LOADING", os.environ.get("CUDA_MODULE_LOADING"), "workers", os.environ.get("CARMA_ORDER_WORKERS"))
for opt in sys.argv[1:] or ("native", "device"):
    model = C.CarmaModel(t, y, e)
    model.mle_optimizer = opt
    for rep in range(3):
        t0 = time.perf_counter()
        mle, pq, aicc = model.choose_order(7, ntrials=100, seed=500, verbose=False)
        print(opt, "rep", rep, "wall_s", round(time.perf_counter() - t0, 3), "selected", [model.p, model.q], round(float(np.min(aicc)), 4), flush=True)
