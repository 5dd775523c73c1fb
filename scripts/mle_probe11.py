"""One on-device fit launch against the number of starts (the 100 starts of a model tiled k times): does the kernel
itself scale when the GPU fills up?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
model = C.CarmaModel(t, y, e)
for (p, q) in ((7, 4), (5, 2), (3, 1)):
    kind, x0, lo, hi, prior, flags = model.mle_starts(p, q, 100, seed=500)
    for dev in (True, False):
        model.series.mle_batch(kind, p, q, x0[:4], lo, hi, prior=prior, flags=flags, maxiter=3, on_device=dev)
        for k in (1, 2, 4, 8, 16, 28):
            xx = np.tile(x0, (k, 1))
            t0 = time.perf_counter()
            x, f, nit, nfev = model.series.mle_batch(kind, p, q, xx, lo, hi, prior=prior, flags=flags, on_device=dev)
            print((p, q), "device" if dev else "native", "starts", 100 * k, "wall_s", round(time.perf_counter() - t0, 3), "nfev", nfev, flush=True)
