"""K4 (one light curve per thread) throughput against the model order: simulated surveys of 200,000 curves x 500 points,
one theta per curve (the generating theta), three timed launches each."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth
import torch
for (p, q) in ((2, 0), (3, 1), (4, 2), (5, 3), (6, 2), (7, 4)):
    t, y, e = synth.readme_series(300, 7)
    rng = np.random.default_rng(p)
    th = synth.prior_draws(1, p, q, t, y, rng)[0]
    th[1] = 1.0
    kind = C.KIND_CARMA if q else C.KIND_CARP
    ms = C.MultiSeries.simulate(200000, 500, kind, p, q, th, seed=3)
    thm = np.tile(th, (200000, 1))
    ms.loglik(kind, p, q, thm)
    d_th = torch.from_numpy(thm).cuda()
    d_out = torch.empty(200000, dtype=torch.float64, device="cuda")
    pri = torch.from_numpy(ms.default_priors().view(np.float64).reshape(-1, 6).copy()).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(4):
        ev[0].record()
        ms.loglik_dev(kind, p, q, pri.data_ptr(), d_th.data_ptr(), d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        ev[1].record(); ev[1].synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    print("CARMA(%d,%d): %.3f ms, %.3e curves/s, %.1f TFLOP/s algorithmic, finite %.3f" % (
        p, q, best, 200000 / best * 1e3, 200000 * 499 * (20 * p * p + 36 * p + 7) / best * 1e3 / 1e12, float(torch.isfinite(d_out).float().mean())), flush=True)
    ms.close()
