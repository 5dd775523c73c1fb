"""Dump the rows of the config-2 batch where GPU and oracle disagree most (diagnostics)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
from oracle import oracle as O

t, y, e = synth.readme_series(270, 270)
th = synth.theta_batch(65536, t, y, seed=2)
s = C.Series(t, y, e)
pr = s.default_prior()
got = s.loglik(C.KIND_CARMA, 5, 3, th, prior=pr)
opr = O.default_prior(t, y)
want = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=opr)
ld = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=opr, long_double=True)
fin = np.isfinite(want) & np.isfinite(got)
rel = np.zeros_like(want)
rel[fin] = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1)
noise = np.zeros_like(want)
noise[fin] = np.abs(ld[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1)
relld = np.zeros_like(want)
relld[fin] = np.abs(got[fin] - ld[fin]) / np.maximum(np.abs(want[fin]), 1)
order = np.argsort(-rel)[:25]
out = {"n_over_1e-9": int((rel > 1e-9).sum()), "n_over_1e-10": int((rel > 1e-10).sum()),
       "n_over_1e-11": int((rel > 1e-11).sum()), "median_rel": float(np.median(rel[fin])),
       "p99_rel": float(np.quantile(rel[fin], 0.99)),
       "class_mismatch": int((np.isfinite(want) != np.isfinite(got)).sum()),
       "oracle_noise_over_1e-9": int((noise > 1e-9).sum()),
       "rows": [dict(i=int(i), got=float(got[i]), want=float(want[i]), ld=float(ld[i]), rel=float(rel[i]),
                     noise=float(noise[i]), rel_vs_ld=float(relld[i]), theta=[float(x) for x in th[i]]) for i in order]}
json.dump(out, open("gpurun_out/debug_parity.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "rows"}))
for r in out["rows"][:8]:
    print(r["i"], r["got"], r["want"], r["rel"], r["noise"], r["rel_vs_ld"])
