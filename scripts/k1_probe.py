"""K1 cost model probe: time loglik_batch_kernel<P> (65,536 theta, device-resident) on the README-style series
truncated to several lengths; a linear fit  t(ny) = prologue + (ny - 1) * step  separates the theta-transform
prologue (+ launch, tail) from the per-Kalman-step cost.  Also reports the all-conjugate batch against a batch in
which half of the rows have one real root pair (generic loop); both with CARMA_IGNORE_BOUNDS so that every row runs
the whole recursion.  One JSON line.

    python scripts/k1_probe.py [p] [q]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import carma_pack_b200 as C  # noqa: E402
from carma_pack_b200 import synth  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 5
q = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 65536
kind = C.KIND_CARMA if q > 0 else C.KIND_CARP
t, y, e = synth.readme_series(1080, 270)
th = synth.theta_batch(N, t[:270], y[:270], p=p, q=q, seed=0)
rng = np.random.default_rng(1)
th_mixed = th.copy()
half = rng.uniform(size=N) < 0.5
q1 = np.exp(th_mixed[half, 3])
th_mixed[half, 4] = np.log(np.sqrt(4.0 * q1) * rng.uniform(1.05, 3.0, half.sum()))   # first quadratic factor: two real roots
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
out = {"p": p, "q": q, "n_theta": N, "ms": {}, "ms_mixed": {}}
for name, batch in (("ms", th), ("ms_mixed", th_mixed)):
    d_theta = torch.from_numpy(batch).cuda()
    d_out = torch.empty(N, dtype=torch.float64, device="cuda")
    for ny in (2, 10, 34, 90, 270, 540, 1080):
        s = C.Series(t[:ny], y[:ny], e[:ny])
        pr = C.Series(t[:270], y[:270], e[:270]).default_prior()   # same prior bounds at every length
        for _ in range(3):
            s.loglik_dev(kind, p, q, d_theta.data_ptr(), d_out.data_ptr(), N, pr, C.IGNORE_BOUNDS, stream)
        torch.cuda.synchronize()
        best = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.loglik_dev(kind, p, q, d_theta.data_ptr(), d_out.data_ptr(), N, pr, C.IGNORE_BOUNDS, stream)
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1))
        out[name][str(ny)] = float(np.median(best))
        if ny == 270:
            out[name + "_finite_frac_270"] = float(torch.isfinite(d_out).double().mean().item())
        s.close()
for name in ("ms", "ms_mixed"):
    nys = np.array([int(k) for k in out[name]], float)
    ms = np.array(list(out[name].values()))
    sel = nys >= 34
    A = np.vstack([np.ones(sel.sum()), nys[sel] - 1]).T
    c, *_ = np.linalg.lstsq(A, ms[sel], rcond=None)
    out[name + "_fit"] = {"prologue_ms": float(c[0]), "step_us": float(1e3 * c[1]),
                          "cycles_per_warp_step_at_3.5_warps_per_scheduler": float(c[1] * 1e-3 * 1.965e9 / 3.5)}
print(json.dumps(out))
