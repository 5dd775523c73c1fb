"""Per-model solo fit time of the config-4 grid (seed as choose_order uses it), host-loop against on-device optimiser."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
model = C.CarmaModel(t, y, e)
pqlist = [(p, q) for p in range(1, 8) for q in range(p)]
rows = []
for k, (p, q) in enumerate(pqlist):
    kind, x0, lo, hi, prior, flags = model.mle_starts(p, q, 100, seed=500 + k)
    row = {"p": p, "q": q}
    for name, dev in (("native", False), ("device", True)):
        model.series.mle_batch(kind, p, q, x0[:4], lo, hi, prior=prior, flags=flags, maxiter=3, on_device=dev)
        t0 = time.perf_counter()
        x, f, nit, nfev = model.series.mle_batch(kind, p, q, x0, lo, hi, prior=prior, flags=flags, on_device=dev)
        row[name] = [round(time.perf_counter() - t0, 4), nit, nfev, round(float(f.min()), 4)]
    rows.append(row)
    print(row, flush=True)
print("sum native", sum(r["native"][0] for r in rows), "sum device", sum(r["device"][0] for r in rows))
