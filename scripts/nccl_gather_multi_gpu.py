"""carma_gather_summaries (C ABI, NCCL all-gather of per-rank summaries) on N GPUs of one box WITHOUT torch.distributed:
rank 0 makes the NCCL id (carma_comm_unique_id) and hands it to the other ranks through a file; every rank fits its
cost-weighted share of a small choose_order grid with carma_mle_batch and all-gathers [-loglik, theta-hat] per model.
Run under gpurun --gpus N:   python scripts/nccl_gather_multi_gpu.py N     (prints one JSON line from rank 0)"""
import ctypes
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, idfile, q):
    import carma_pack_b200 as C
    from carma_pack_b200 import synth, sharding
    lib = C._lib.lib
    ident = ctypes.create_string_buffer(128)
    if rank == 0:
        C._lib.check(lib.carma_comm_unique_id(ident), "carma_comm_unique_id")
        with open(idfile + ".tmp", "wb") as f:
            f.write(ident.raw)
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.01)
        ident = ctypes.create_string_buffer(open(idfile, "rb").read(), 128)
    comm = ctypes.c_void_p()
    C._lib.check(lib.carma_comm_init_rank(world, rank, ident, rank, ctypes.byref(comm)), "carma_comm_init_rank")
    t, y, e = synth.readme_series(300, 300)
    model = C.CarmaModel(t, y, e, device=rank)
    pq = [(2, 0), (2, 1), (3, 0), (3, 1), (3, 2), (4, 1)]
    ntrials, dmax = 16, 8
    costs = np.repeat([(20 * p * p + 36 * p + 7) * (4 + p + q) for p, q in pq], ntrials).astype(float)
    mine = sharding.partition_weighted(costs, world, rank)
    table = np.full((len(pq), 1 + dmax), np.nan)
    table[:, 0] = np.inf
    t0 = time.perf_counter()
    for k in sorted(set(int(u) // ntrials for u in mine)):
        tr = [int(u) % ntrials for u in mine if int(u) // ntrials == k]
        p, qq = pq[k]
        mle = model.get_mle(p, qq, ntrials=len(tr), seed=77 + k, trial_offset=min(tr))
        table[k, 0] = mle.fun
        table[k, 1:1 + len(mle.x)] = mle.x
    fit_s = time.perf_counter() - t0
    flat = np.ascontiguousarray(np.nan_to_num(table, nan=0.0, posinf=1e300).ravel())
    allv = np.empty(world * flat.size)
    dp = ctypes.POINTER(ctypes.c_double)
    t0 = time.perf_counter()
    C._lib.check(lib.carma_gather_summaries(comm, flat.ctypes.data_as(dp), flat.size, allv.ctypes.data_as(dp), None), "carma_gather_summaries")
    gather_s = time.perf_counter() - t0
    allt = allv.reshape(world, len(pq), 1 + dmax)
    best = allt[np.argmin(allt[:, :, 0], axis=0), np.arange(len(pq))]
    C._lib.check(lib.carma_comm_destroy(comm), "carma_comm_destroy")
    q.put((rank, best[:, 0].tolist(), fit_s, gather_s))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    idfile = os.path.join(tempfile.mkdtemp(), "nccl_id")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, idfile, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=900) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
    # single-process reference on GPU 0
    import carma_pack_b200 as C
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(300, 300)
    model = C.CarmaModel(t, y, e, device=0)
    pq = [(2, 0), (2, 1), (3, 0), (3, 1), (3, 2), (4, 1)]
    single = [model.get_mle(p, qq, ntrials=16, seed=77 + k).fun for k, (p, qq) in enumerate(pq)]
    same = all(r[1] == single for r in res)
    print(json.dumps({"world": world, "gathered_equals_single_process_bitwise": same, "neg_loglik": single,
                      "fit_s_per_rank": [r[2] for r in res], "gather_s_per_rank": [r[3] for r in res]}))
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
