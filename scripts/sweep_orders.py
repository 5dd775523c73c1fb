"""K1 throughput for every AR order p = 1..7 (q = p-1, CAR1 for p = 1): 65,536 theta on one ny=270 series."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth

t, y, e = synth.readme_series(270, 270)
s = C.Series(t, y, e)
pr = s.default_prior()
N = 65536
peak = C._lib.fp64_peak_tflops(0)
out = []
for p in range(1, 8):
    q = max(p - 1, 0)
    kind = C.KIND_CAR1 if p == 1 else C.KIND_CARMA
    rng = np.random.default_rng(p)
    if p == 1:
        th = np.column_stack([np.full(N, y.std()), np.ones(N), np.full(N, y.mean()), rng.uniform(-5, -1, N)])
    else:
        th = synth.prior_draws(N, p, q, t, y, rng)
    d_th = torch.from_numpy(np.ascontiguousarray(th)).cuda()
    d_o = torch.empty(N, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        s.loglik_dev(kind, p, q, d_th.data_ptr(), d_o.data_ptr(), N, pr, 0, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        s.loglik_dev(kind, p, q, d_th.data_ptr(), d_o.data_ptr(), N, pr, 0, st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    fin = float(torch.isfinite(d_o).float().mean())
    fe = 269 * (20 * p * p + 36 * p + 7) + (2.0 / 3.0) * 8 * p ** 3 + 22 * p * p
    out.append(dict(p=p, q=q, ms=ms, evals_per_s=N / (ms * 1e-3), finite_frac=fin,
                    tflops_algorithmic=N * fe / (ms * 1e-3) / 1e12, frac_of_fp64_peak=N * fe / (ms * 1e-3) / 1e12 / peak))
    print(out[-1])
json.dump(dict(fp64_peak_tflops=peak, rows=out), open("gpurun_out/sweep_orders.json", "w"), indent=1)
