"""How the PT kernel's time depends on how many 64-thread blocks each SM hosts (config 3 shape: CARMA(5,3), ny = 1000,
10 temperatures = 6 ensembles per block).  148 SMs x 4 sub-partitions: 592 blocks put exactly two warps on every
scheduler, 683 (config 3) put three on a third of them.  One JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C  # noqa: E402
from carma_pack_b200 import synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t, y, e = synth.readme_series(1000, 1000)
s = C.Series(t, y, e)
out = {"iterations": iters, "rows": []}
groups = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (148, 296, 444, 592, 683, 740, 888, 1000, 1366)
out["CARMA_PT_PIPE"] = os.environ.get("CARMA_PT_PIPE", "auto")
for blocks in groups:
    n_ens = blocks * 6 if blocks not in (683, 1366) else (4096 if blocks == 683 else 8192)
    for slice_env in ("0", None, "592", "888"):
        if slice_env in ("592", "888") and blocks <= int(slice_env):
            continue
        if slice_env is None:
            os.environ.pop("CARMA_PT_SLICE", None)
        else:
            os.environ["CARMA_PT_SLICE"] = slice_env
        s.pt_run(C.KIND_CARMA, 5, 3, 1, 1, ntemps=10, n_ensembles=n_ens, seed=3)
        t0 = time.perf_counter()
        s.pt_run(C.KIND_CARMA, 5, 3, iters // 2, iters - iters // 2, ntemps=10, n_ensembles=n_ens, seed=3)
        w = time.perf_counter() - t0
        out["rows"].append({"groups": blocks, "ensembles": n_ens, "CARMA_PT_SLICE": slice_env or "auto", "wall_ms": 1e3 * w,
                            "us_per_tick": 1e6 * w / (iters + 9), "ens_iters_per_s": n_ens * iters / w})
os.environ.pop("CARMA_PT_SLICE", None)
s.close()
print(json.dumps(out))
