"""Site sharding across ranks (one process per GPU) and the single gather of output rows.

Sites are independent, so the path shards with no data-path collective: rank r scores a
contiguous range of sites, balanced by evidence rows rather than by site count,
and ONE collective at the end brings the fixed-width 80-byte output rows to rank 0
(`torch.distributed` gather: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import evidence as ev


def shard_bounds(batch, world):
    """[lo_0, lo_1, ..., lo_world]: contiguous site ranges with ~equal row counts."""
    n = batch.n_sites
    if n == 0:
        return [0] * (world + 1)
    if hasattr(batch, "work"):                        # CompactBatch
        work = batch.work() + 4
    else:
        work = batch.sites[:, 12].astype(np.int64) + batch.sites[:, 15].astype(np.int64) + 4   # + fixed per-site cost
    csum = np.cumsum(work)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, csum[-1] * r / world, side="left")))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def local_shard(batch, rank, world):
    b = shard_bounds(batch, world)
    shard = batch.slice_sites(b[rank], b[rank + 1])
    shard.order = shard.length_order() if shard.n_sites else None
    return shard, b


def gather_rows(local_rows, bounds, rank, world, device=None):
    """Gather per-rank OUT_DTYPE rows to rank 0 in site order.  `local_rows` is a numpy
    OUT_DTYPE array or a [n, 80] uint8 tensor (device tensors gather over NCCL)."""
    import torch
    import torch.distributed as dist
    if isinstance(local_rows, np.ndarray):
        t = torch.from_numpy(local_rows.view(np.uint8).reshape(-1, ev.OUT_BYTES).copy())
        if device is not None:
            t = t.to(device)
    else:
        t = local_rows
    if world == 1:
        return t
    sizes = [bounds[i + 1] - bounds[i] for i in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad, ev.OUT_BYTES), dtype=torch.uint8, device=t.device)
    buf[:t.shape[0]] = t
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return torch.cat([out[i][:sizes[i]] for i in range(world)], dim=0)


def rows_from_tensor(t):
    return t.cpu().numpy().reshape(-1).view(ev.OUT_DTYPE).copy()
