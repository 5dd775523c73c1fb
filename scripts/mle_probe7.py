"""choose_order(7) x 100 starts with the on-device optimiser: wall time against the number of concurrent model fits
and with the series read from shared memory or from global memory (CARMA_MLE_SERIES_SMEM, a switch that only the build
of that experiment had; with one launch per model, as choose_order worked before carma_mle_grid_device)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
out = []
for smem in ("1", "0"):
    os.environ["CARMA_MLE_SERIES_SMEM"] = smem
    for opt, workers in (("native", 16), ("device", 16), ("device", 8), ("device", 4), ("device", 2), ("device", 28)):
        if opt == "native" and smem == "0":
            continue
        os.environ["CARMA_ORDER_WORKERS"] = str(workers)
        model = C.CarmaModel(t, y, e)
        model.mle_optimizer = opt
        model.choose_order(2, ntrials=4, seed=1, verbose=False)
        t0 = time.perf_counter()
        mle, pq, aicc = model.choose_order(7, ntrials=100, seed=500, verbose=False)
        out.append({"optimizer": opt, "series_smem": smem, "workers": workers, "wall_s": round(time.perf_counter() - t0, 3),
                    "selected": [model.p, model.q], "best_aicc": float(np.min(aicc))})
        print(out[-1], flush=True)
print(json.dumps(out))
