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
        best = sharding.best_per_model(table, dist)
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


# ---- CarmaModel.choose_order(dist=...) end to end on CPU: the sharding and gather logic with a stand-in fit
class _FakeSeries:
    def __init__(self, *a, **k):
        pass

    def close(self):
        pass


def _fake_get_mle(self, p, q, ntrials=100, njobs=1, seed=None, maxiter=200, trial_offset=0, series=None, optimizer="native"):
    """Deterministic stand-in for a fit: the value of trial j of model (p,q) depends only on (p, q, j), so the
    best over any partition of the trials must equal the best over all of them."""
    from scipy.optimize import OptimizeResult
    d = 4 if p == 1 else 3 + p + q
    vals = [200.0 + 3.0 * abs(p - 3) + 1.5 * q + ((7 * (trial_offset + j) + 3 * p + q) % 11) * 0.1 for j in range(ntrials)]
    j = int(np.argmin(vals))
    return OptimizeResult(x=np.full(d, float(trial_offset + j)), fun=float(vals[j]), success=True, nit=7, nfev=99)


def _choose_order_worker(rank, world, port, q):
    import torch.distributed as dist
    import carma_pack_b200.carma_pack as cp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cp.Series = _FakeSeries
        cp.CarmaModel.get_mle = _fake_get_mle
        t = np.arange(120.0)
        model = cp.CarmaModel(t, np.sin(t), np.full(t.size, 0.1))
        model.mle_optimizer = "native"   # the per-model path, whose get_mle the stand-in replaces
        mle, pqlist, aicc = model.choose_order(4, ntrials=10, seed=1, verbose=False, dist=dist)
        q.put((rank, list(pqlist), list(aicc), float(mle.fun), np.asarray(mle.x).tolist(), (model.p, model.q)))
    finally:
        dist.destroy_process_group()


def test_choose_order_sharded_over_two_ranks_equals_single_process():
    import torch.multiprocessing as mp
    import carma_pack_b200.carma_pack as cp
    # single process, same stand-in fit
    saved = (cp.Series, cp.CarmaModel.get_mle)
    try:
        cp.Series = _FakeSeries
        cp.CarmaModel.get_mle = _fake_get_mle
        t = np.arange(120.0)
        model = cp.CarmaModel(t, np.sin(t), np.full(t.size, 0.1))
        model.mle_optimizer = "native"
        mle1, pq1, aicc1 = model.choose_order(4, ntrials=10, seed=1, verbose=False)
        best1 = (model.p, model.q)
    finally:
        cp.Series, cp.CarmaModel.get_mle = saved
    assert len(pq1) == 10   # p = 1..4, q < p
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_choose_order_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] == [tuple(x) for x in pq1] or r[1] == pq1
        np.testing.assert_allclose(r[2], aicc1, rtol=0, atol=1e-12)   # same AICc table on every rank
        assert abs(r[3] - mle1.fun) < 1e-12 and r[5] == best1
        assert r[4] == np.asarray(mle1.x).tolist()                     # and the same theta-hat (from the best trial)
