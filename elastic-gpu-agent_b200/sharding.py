"""Row-sharded snapshot scoring across ranks.

Rank g scores request rows [lo_g, hi_g); the table is replicated; the only exchange is
one all-gather of the per-rank demand vectors (int64[2*D]: core sums then mem sums),
after which every rank applies the summed demand and holds the same table'.

The collective and the bookkeeping live here, above the C ABI; the scan and the
table update are the CUDA library's (`BestFitAllocator.bestfit_dev`,
`apply_deltas_dev`).  `combine_demands` is the host restatement of
`apply_deltas_kernel` used where the vectors are already on the host.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first total_rows % world ranks get one extra row."""
    if world < 1 or not (0 <= rank < world) or total_rows < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_demands(delta, world: int, group=None):
    """all_gather of one rank's demand vector (torch tensor int64[2*D], any device the
    process group supports) -> tensor [world, 2*D], identical on every rank."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * delta.numel(), dtype=delta.dtype, device=delta.device)
    if world == 1:
        out.copy_(delta)
    else:
        dist.all_gather_into_tensor(out, delta.contiguous(), group=group)
    return out.view(world, delta.numel())


def combine_demands(free_core, free_mem, gathered) -> np.ndarray:
    """table' (int32[3*D]: free_core', free_mem', oversub) from the gathered demand
    vectors [world, 2*D] — spec §2.4 applied to the sum over ranks."""
    fc = np.asarray(free_core, dtype=np.int64)
    fm = np.asarray(free_mem, dtype=np.int64)
    g = np.asarray(gathered, dtype=np.int64)
    D = fc.size
    if g.ndim != 2 or g.shape[1] != 2 * D:
        raise ValueError("gathered must be [world, 2*D]")
    tot = g.sum(axis=0)
    c = fc - tot[:D]
    m = fm - tot[D:]
    lo, hi = -(2 ** 31), 2 ** 31 - 1
    over = ((c < 0) | (m < 0)).astype(np.int64)
    return np.concatenate([np.clip(c, lo, hi), np.clip(m, lo, hi), over]).astype(np.int32)


def sharded_step(alloc, d_core: int, d_mem: int, rows: int, d_idx: int, delta, gathered_flat, table_out,
                 world: int, stream: int, commit: bool = False, group=None):
    """One multi-GPU step on device buffers: scan the local shard, all-gather the demand
    vectors, apply their sum.  `delta`, `gathered_flat` (int64[world*2*D]) and `table_out`
    (int32[3*D]) are CUDA tensors; returns nothing (outputs are in the tensors)."""
    import torch.distributed as dist
    alloc.bestfit_dev(d_core, d_mem, rows, d_idx, delta.data_ptr(), 0, False, stream)
    if world > 1:
        dist.all_gather_into_tensor(gathered_flat, delta, group=group)
    else:
        gathered_flat.copy_(delta)
    alloc.apply_deltas_dev(gathered_flat.data_ptr(), world, table_out.data_ptr(), commit, stream)

