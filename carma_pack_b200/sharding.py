"""Multi-GPU plumbing: one process per GPU (torch.distributed), independent units (theta rows,
PT ensembles, light curves, (p,q,start) fits) are block-partitioned over ranks with no data-path
collective; only per-model summaries (AICc, log-likelihood, theta-hat) are gathered at the end
(SURVEY 8e).  Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
import numpy as np


def partition(n_units, world_size, rank):
    """Contiguous block partition: unit i -> rank floor(i * world / n).  Returns (start, stop)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    start = -((-rank * n_units) // world_size)
    stop = -((-(rank + 1) * n_units) // world_size)
    return start, stop


def partition_weighted(costs, world_size, rank):
    """Cost-weighted contiguous partition (choose_order: cost ~ F_step(p) (d+1) per start): rank r
    owns the units whose cumulative-cost midpoint falls in [r, r+1) * total / world."""
    costs = np.asarray(costs, dtype=float)
    mid = np.cumsum(costs) - 0.5 * costs
    owner = np.minimum((mid * world_size / costs.sum()).astype(int), world_size - 1)
    idx = np.nonzero(owner == rank)[0]
    return idx


def gather_summaries(local, dist=None):
    """All-gather a small float64 summary array (same shape on every rank) -> (world, ...) array.
    The only collective of the system (K6): a few kB over NVLink."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local[None, ...]
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def best_per_model(local_table, dist=None):
    """local_table: (n_models, 1 + k) rows [key, payload...] with key = -loglik of the best fit this rank found for
    the model (+inf for models it did not fit); the payload travels with it (theta-hat, nit, nfev, nan padding).
    Returns, per model, the row of the rank with the smallest key.  For a fixed model AICc is an increasing function
    of -loglik, so this is also the rank with the best AICc; the AICc itself is computed by the caller afterwards."""
    allt = gather_summaries(local_table, dist)
    pick = np.argmin(allt[:, :, 0], axis=0)
    return allt[pick, np.arange(allt.shape[1])]

