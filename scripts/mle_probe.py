"""Host-loop optimiser (carma_mle_batch) against the on-device optimiser (carma_mle_batch_device) on the config-4
series: per model the best and median -loglik over the starts, iteration counts, evaluations and wall time; then the
whole choose_order(7) grid with either.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C  # noqa: E402
from carma_pack_b200 import synth  # noqa: E402

t, y, e = synth.readme_series(500, 500)
model = C.CarmaModel(t, y, e)
out = {"models": [], "choose_order": {}}
for p, q in ((2, 0), (3, 1), (5, 3), (7, 4)):
    kind, x0, lo, hi, prior, flags = model.mle_starts(p, q, 100, seed=11)
    row = {"p": p, "q": q}
    for name, dev in (("native", False), ("device", True)):
        model.series.mle_batch(kind, p, q, x0[:4], lo, hi, prior=prior, flags=flags, maxiter=3, on_device=dev)
        t0 = time.perf_counter()
        x, f, nit, nfev = model.series.mle_batch(kind, p, q, x0, lo, hi, prior=prior, flags=flags, on_device=dev)
        row[name] = {"wall_s": time.perf_counter() - t0, "best": float(f.min()), "median": float(np.median(f)), "nit": nit,
                     "nfev": nfev, "n_within_1e-3_of_best": int((f < f.min() + 1e-3).sum())}
        row[name + "_f"] = f
    d = np.abs(row["native_f"] - row["device_f"])
    row["per_start_abs_diff_quantiles_50_90_100"] = [float(np.quantile(d, z)) for z in (0.5, 0.9, 1.0)]
    row["starts_equal_to_1e-6"] = int((d < 1e-6).sum())
    del row["native_f"], row["device_f"]
    out["models"].append(row)
for name in ("native", "device"):
    model.mle_optimizer = name
    model.choose_order(2, ntrials=4, seed=1, verbose=False)
    t0 = time.perf_counter()
    mle, pq, aicc = model.choose_order(7, ntrials=100, seed=500, verbose=False)
    out["choose_order"][name] = {"wall_s": time.perf_counter() - t0, "selected": [model.p, model.q], "best_aicc": float(np.min(aicc)),
                                 "aicc": [float(a) for a in aicc]}
print(json.dumps(out))
