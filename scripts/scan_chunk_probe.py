"""Time of one CARMA(3,1) LogDensity on a single long series through the associative-scan kernels against the chunk
length (points folded per thread).  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C  # noqa: E402
from carma_pack_b200 import synth  # noqa: E402

out = {}
for ny in (100000, 1000000, 4000000):
    rng = np.random.default_rng(5)
    t = np.cumsum(rng.uniform(0.5, 1.5, ny))
    y = rng.normal(0, 1, ny)
    e = np.full(ny, 0.3)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    ar = synth.roots_to_logquad(np.array([-0.05 + 0.3j, -0.05 - 0.3j, -0.2]))
    th = torch.tensor([[1.0, 1.0, 0.0] + list(ar) + [np.log(1.0 / 3.0)]], dtype=torch.float64).cuda()
    o = torch.empty(1, dtype=torch.float64, device="cuda")
    row = {}
    for chunk in (16, 32, 48, 64, 96, 128, 192, 256, 512):
        for _ in range(3):
            s.loglik_scan_dev(C.KIND_CARMA, 3, 1, th.data_ptr(), o.data_ptr(), 1, pr, C.IGNORE_BOUNDS, chunk, 0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            s.loglik_scan_dev(C.KIND_CARMA, 3, 1, th.data_ptr(), o.data_ptr(), 1, pr, C.IGNORE_BOUNDS, chunk, 0)
        b.record()
        torch.cuda.synchronize()
        row[str(chunk)] = {"ms": a.elapsed_time(b) / 10, "value": float(o.item())}
    out[str(ny)] = row
    s.close()
print(json.dumps(out))
