"""KalmanFilterp::Filter (GetMean/GetVar) and Predict through the C ABI: the real-half fast path (time-parallel forward
filter, queries resume from stored states) against the general complex kernels (CARMA_PREDICT_GENERAL=1: Filter on one
thread, every query re-filters from the first point).  Wall time of the host call, H2D/D2H included.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C  # noqa: E402
from carma_pack_b200 import synth  # noqa: E402

ar, ma, sigsqr = synth.carma31_truth()
ma = np.concatenate([ma, np.zeros(3 - len(ma))])[:3]
rng = np.random.default_rng(0)
out = {"model": "CARMA(3,1)", "filter": [], "predict": []}


def timed(fn, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


for ny in (270, 1000, 10000, 100000, 1000000):
    t = np.cumsum(rng.uniform(0.5, 1.5, ny))
    y = rng.standard_normal(ny)
    e = np.full(ny, 0.3)
    s = C.Series(t, y, e)
    row = {"ny": ny}
    for mode, name in (("0", "fast_ms"), ("1", "general_ms")):
        if mode == "1" and ny > 100000:
            continue
        os.environ["CARMA_PREDICT_GENERAL"] = mode
        row[name] = 1e3 * timed(lambda: s.filter(sigsqr, ar, ma))
    out["filter"].append(row)
    for nq in (512, 8192):
        if ny > 100000:
            continue
        tq = np.sort(rng.uniform(t[0] - 10, t[-1] + 10, nq))
        row = {"ny": ny, "nq": nq}
        res = {}
        for mode, name in (("0", "fast_ms"), ("1", "general_ms")):
            if mode == "1" and ny * nq > 2e9:
                continue
            os.environ["CARMA_PREDICT_GENERAL"] = mode
            row[name] = 1e3 * timed(lambda: res.__setitem__(mode, s.predict(sigsqr, ar, ma, tq)), reps=2)
        if "1" in res:
            row["max_rel_diff_mean"] = float(np.max(np.abs(res["0"][0] - res["1"][0]) / np.maximum(np.abs(res["1"][0]), 1e-3)))
            row["max_rel_diff_var"] = float(np.max(np.abs(res["0"][1] - res["1"][1]) / res["1"][1]))
        out["predict"].append(row)
    s.close()
os.environ.pop("CARMA_PREDICT_GENERAL", None)
print(json.dumps(out))
