"""BASELINE config 1 on the GPU path: README CARMA(5,3) light curve (ny=270), run_mcmc(50000) =
75,000 iterations x 10 temperatures, ONE ensemble (the reference's own CPU-runnable case)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth

nsamples = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
t, y, e = synth.readme_series(270, 270)
model = C.CarmaModel(t, y, e, p=5, q=3)
model.series  # create the device series outside the timed region
t0 = time.perf_counter()
sample = model.run_mcmc(nsamples, seed=11)
wall = time.perf_counter() - t0
tr = sample._samples
truth = synth.readme_theta(3)
out = {"config": "README CARMA(5,3) ny=270 run_mcmc(%d): %d iterations x 10 temperatures, 1 ensemble" % (nsamples, nsamples * 3 // 2),
       "wall_s": wall, "evals_per_s": nsamples * 1.5 * 10 / wall,
       "accept_rate_cool": float(sample.accept_rates[0, 0]), "exchange_rates": [float(x) for x in sample.exchange_rates[0]],
       "post_mean_sigma_y": float(np.sqrt(tr["var"]).mean()), "post_sd_sigma_y": float(np.sqrt(tr["var"]).std()),
       "post_mean_mu": float(tr["mu"].mean()), "truth_sigma_y": 2.3, "truth_mu": 17.0,
       "max_logpost": float(tr["logpost"].max()), "logpost_at_truth": float(model.series.loglik(C.KIND_CARMA, 5, 3, truth)[0])}
print(json.dumps(out))
