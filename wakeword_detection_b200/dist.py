"""Multi-GPU plumbing: streams/clips shard trivially (every posterior depends on its own
L mel frames only), so there is no data-path collective; the only exchange is one
all-reduce (sum) of the int64 FAR/FRR counters — a few KB over NVLink (NCCL) on the GPU
box, gloo in the CPU tests.  (The reference's inference path is single-process; its only
parallel construct is tf.distribute.MirroredStrategy in WaveNet *training*,
wwdetect/wavenet/train_wavenet.py:39, out of scope.)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def rank_world() -> Tuple[int, int]:
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n_items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_round_robin(items: Sequence, rank: int, world: int) -> List:
    return list(items[rank::world])


def time_chunks(n_windows: int, world: int, halo_lo: int = 16, halo_hi: int = 14):
    """Splits one long posterior trajectory of n_windows into `world` chunks for the FAR
    edge count: chunk r counts windows [b, e) and needs posteriors [b - lo, e + hi) where
    lo/hi are clipped at the true ends (the 30-tap mean looks 15 back / 14 ahead and the
    edge detector one further back).  Returns a list of
    (begin, end, lo, hi)."""
    out = []
    for r in range(world):
        b, e = shard_range(n_windows, r, world)
        lo = min(halo_lo, b)
        hi = min(halo_hi, n_windows - e)
        out.append((b, e, lo, hi))
    return out


def all_reduce_counters(*tensors):
    """Sum int64 counter tensors over all ranks with ONE all-reduce; returns the reduced
    tensors (same shapes).  No-op when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors if len(tensors) > 1 else tensors[0]
    flat = torch.cat([t.reshape(-1).to(torch.int64) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    out, o = [], 0
    for t in tensors:
        out.append(flat[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    return tuple(out) if len(out) > 1 else out[0]
