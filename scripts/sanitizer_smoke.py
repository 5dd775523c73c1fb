"""Small end-to-end run of every kernel family, sized for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth

rng = np.random.default_rng(0)
t, y, e = synth.readme_series(601, 3)          # > 512 points: double-buffered TMA path in K1
s = C.Series(t, y, e)
pr = s.default_prior()
th = synth.theta_batch(200, t, y, seed=1)
lp = s.loglik(C.KIND_CARMA, 5, 3, th, prior=pr)
lp2 = s.loglik_scan(C.KIND_CARMA, 5, 3, th[:3], prior=pr, chunk=16)
m, v = s.filter(1.0, [-0.1 + 0.3j, -0.1 - 0.3j, -0.05], [1.0, 0.5, 0.0])
qm, qv = s.predict(1.0, [-0.1 + 0.3j, -0.1 - 0.3j, -0.05], [1.0, 0.5, 0.0], [t[0] - 1, t[5] + 0.1, t[-1] + 3])
r = s.pt_run(C.KIND_CARMA, 5, 3, 4, 4, ntemps=10, n_ensembles=7, seed=2, prior=pr, record_trace=True)
r1 = s.pt_run(C.KIND_CAR1, 1, 0, 4, 4, ntemps=1, n_ensembles=3, seed=2)
# the three launch shapes of the PT kernel: warp-specialised (default for small launches), plain, time-sliced
shapes = {}
for name, env in (("help", {"CARMA_PT_HELP": "1"}), ("plain", {"CARMA_PT_HELP": "0", "CARMA_PT_SLICE": "0"}),
                  ("sliced", {"CARMA_PT_HELP": "0", "CARMA_PT_SLICE": "2"})):
    if os.environ.get("SANITIZER_SKIP_HELP") and name == "help":
        continue
    os.environ.update(env)
    shapes[name] = s.pt_run(C.KIND_CARMA, 5, 3, 6, 10, ntemps=10, n_ensembles=20, seed=5, prior=pr)
    for k in env:
        os.environ.pop(k)
names = list(shapes)
for a in names[1:]:
    assert np.array_equal(shapes[names[0]]["samples"], shapes[a]["samples"]), a
ysim = s.simulate(1.0, [-0.1 + 0.3j, -0.1 - 0.3j, -0.05], [1.0, 0.5, 0.0], [t[0] - 1, t[5] + 0.1, t[-1] + 3], npaths=3, seed=1)
ts, ys, es, off = [], [], [], [0]
for c in range(9):
    n = int(rng.integers(5, 60))
    tt = np.cumsum(rng.uniform(0.5, 2, n)); ts.append(tt); ys.append(rng.standard_normal(n)); es.append(np.full(n, 0.2)); off.append(off[-1] + n)
ms = C.MultiSeries(np.concatenate(ts), np.concatenate(ys), np.concatenate(es), off)
thm = np.vstack([synth.prior_draws(1, 3, 1, ts[c], ys[c], rng) for c in range(9)])
lm = ms.loglik(C.KIND_CARMA, 3, 1, thm)
rm = ms.pt_run(C.KIND_CARMA, 3, 1, 3, 3, ntemps=4, n_ensembles=2, seed=4)
# all orders through K1 (shared-memory LU scratch of every size, incl. the > 48 KiB opt-in at P = 7)
for p_, q_ in ((1, 0), (2, 1), (3, 0), (4, 2), (6, 3), (7, 5)):
    kind_ = C.KIND_CAR1 if p_ == 1 else (C.KIND_CARMA if q_ else C.KIND_CARP)
    thp = np.tile(np.array([1.0, 1.0, 0.0, np.log(0.1)]), (70, 1)) if p_ == 1 else synth.prior_draws(70, p_, q_, t, y, rng)
    s.loglik(kind_, p_, q_, thp, prior=pr)
sim = C.MultiSeries.simulate(130, 40, C.KIND_CARMA, 3, 1, synth.carma31_theta(), seed=3)
ls = sim.loglik(C.KIND_CARMA, 3, 1, np.tile(synth.carma31_theta(), (130, 1)))
tc = sim.curve(129)
lo_, hi_ = np.full(7, -10.0), np.full(7, 10.0)
xm, fm, nit, nfev = s.mle_batch(C.KIND_CARMA, 3, 1, synth.prior_draws(6, 3, 1, t, y, rng), lo_, hi_, prior=pr,
                                flags=C.IGNORE_BOUNDS, maxiter=3)
# the on-device optimiser: one model, and a grid of models of different orders served from the queue in one launch
xd, fd, nitd, nfevd = s.mle_batch(C.KIND_CARMA, 3, 1, synth.prior_draws(6, 3, 1, t, y, rng), lo_, hi_, prior=pr,
                                  flags=C.IGNORE_BOUNDS, maxiter=3, on_device=True)
grid_jobs = []
for p_, q_, n_ in ((2, 0, 5), (5, 2, 9), (7, 6, 4), (1, 0, 3)):
    kind_ = C.KIND_CAR1 if p_ == 1 else (C.KIND_CARMA if q_ else C.KIND_CARP)
    d_ = C.model_dim(kind_, p_, q_)
    x0_ = np.tile(np.array([1.0, 1.0, 0.0, np.log(0.1)]), (n_, 1)) if p_ == 1 else synth.prior_draws(n_, p_, q_, t, y, rng)
    grid_jobs.append((kind_, p_, q_, x0_, np.full(d_, -10.0), np.full(d_, 10.0), pr, 0 if p_ == 1 else C.IGNORE_BOUNDS))
grid = s.mle_grid(grid_jobs, maxiter=2)
assert len(grid) == 4 and all(np.all(np.isfinite(g[0])) for g in grid)
print("ok", np.isfinite(lp).mean(), lp2[:2], m[:2], qv, r["logposts"].shape, r1["logposts"].shape, lm[:3], rm["logposts"].shape, ls[:2], tc[0][:2], fm[:2], nit, nfev, fd[:2], nitd, nfevd, [g[2] for g in grid])
