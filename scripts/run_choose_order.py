"""BASELINE config 4: choose_order grid pmax=7 (28 (p,q) models) x 100 random starts on an OGLE-like ny=500
series; optionally sharded over the GPUs of one box (torchrun).  Prints the AICc table and the wall time."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carma_pack_b200 as C
from carma_pack_b200 import synth


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pmax = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    ntrials = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    t, y, e = synth.readme_series(500, 500)
    model = C.CarmaModel(t, y, e, device=local_rank)
    t0 = time.perf_counter()
    mle, pqlist, aicc = model.choose_order(pmax, ntrials=ntrials, seed=500, verbose=False, dist=dist)
    wall = time.perf_counter() - t0
    if rank == 0:
        out = {"config": "choose_order pmax=%d (%d models) x %d starts, ny=500, %d GPU(s)" % (pmax, len(pqlist), ntrials, world),
               "wall_s": wall, "best_pq": [model.p, model.q], "aicc": dict(("%d,%d" % pq, a) for pq, a in zip(pqlist, aicc)),
               "best_fun": mle.fun}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
