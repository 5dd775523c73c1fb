"""Which starts of CARMA(6,5) are slow, and what do the evaluations around their end points look like?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
t, y, e = synth.readme_series(500, 500)
model = C.CarmaModel(t, y, e)
p, q = 6, 5
k = [(pp, qq) for pp in range(1, 8) for qq in range(pp)].index((p, q))
kind, x0, lo, hi, prior, flags = model.mle_starts(p, q, 100, seed=500 + k)
s = model.series
s.mle_batch(kind, p, q, x0[:2], lo, hi, prior=prior, flags=flags, maxiter=2, on_device=True)
rows = []
for i in range(100):
    t0 = time.perf_counter()
    x, f, nit, nfev = s.mle_batch(kind, p, q, x0[i:i + 1], lo, hi, prior=prior, flags=flags, on_device=True)
    rows.append((time.perf_counter() - t0, i, nit, nfev, float(f[0]), x[0]))
rows.sort(key=lambda r: -r[0])
for r in rows[:6]:
    print("start %3d: %.3f s, nit %4d, nfev %7d, ms/iter %.3f, evals/iter %.1f, f %.4f" % (r[1], r[0], r[2], r[3], 1e3 * r[0] / max(r[2], 1), r[3] / max(r[2], 1), r[4]))
print("median time", np.median([r[0] for r in rows]), "median ms/iter", np.median([1e3 * r[0] / max(r[2], 1) for r in rows]))
# the neighbourhood of the slowest start's end point: how many evaluations are infeasible?
xs = rows[0][5]
rng = np.random.default_rng(0)
for scale in (1e-8, 1e-4, 1e-2):
    pts = xs + scale * rng.standard_normal((2000, xs.size))
    lp = s.loglik(kind, p, q, pts, prior=prior, flags=flags)
    print("scale", scale, "finite fraction", np.isfinite(lp).mean(), "nan", np.isnan(lp).mean(), "-inf", np.isneginf(lp).mean())
print("x of slowest", np.array2string(xs, precision=6))
print("bounds lo", lo, "hi", hi)
