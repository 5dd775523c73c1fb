"""N>1 path on CPU: world_size-2 gloo processes exercise the partitioning and the summary gather
(the only collective of the system)."""
import os
import socket

import numpy as np
import pytest

from carma_pack_b200 import sharding


def test_partition_covers_everything():
    for n in (0, 1, 7, 64, 65536, 1000003):
        for w in (1, 2, 3, 8):
            parts = [sharding.partition(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for a, b in zip(parts[:-1], parts[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_partition_weighted_balances_choose_order_grid():
    # 28 (p,q) models x 100 starts, cost ~ (20p^2+36p+7)(d+1)
    costs = []
    for p in range(1, 8):
        for q in range(p):
            costs += [(20 * p * p + 36 * p + 7) * (3 + p + q + 1)] * 100
    costs = np.array(costs, dtype=float)
    owned = [sharding.partition_weighted(costs, 8, r) for r in range(8)]
    assert sum(len(o) for o in owned) == costs.size
    assert len(set(np.concatenate(owned))) == costs.size
    loads = np.array([costs[o].sum() for o in owned])
    assert loads.max() / loads.mean() < 1.05


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_models, dmax = 5, 4
        start, stop = sharding.partition(n_models * 10, world, rank)
        table = np.full((n_models, 2 + dmax), np.nan)
        table[:, 0] = np.inf
        rng = np.random.default_rng(100 + rank)
        for u in range(start, stop):  # unit = (model, start index)
            mdl = u // 10
            aicc = 100.0 + mdl + rng.uniform()
            if aicc < table[mdl, 0]:
                table[mdl, 0] = aicc
                table[mdl, 1] = aicc / 2
                table[mdl, 2:] = u
        best = sharding.best_aicc(table, dist)
        g = sharding.gather_summaries(np.array([float(rank), float(stop - start)]), dist)
        q.put((rank, best, g))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    best0, best1 = res[0][1], res[1][1]
    assert np.array_equal(best0, best1, equal_nan=True)   # every rank ends with the same table
    assert np.all(np.isfinite(best0[:, 0]))              # every model was fitted by somebody
    g = res[0][2]
    assert g.shape == (2, 2) and g[:, 1].sum() == 50 and list(g[:, 0]) == [0.0, 1.0]
