"""Multi-GPU host logic: reads are independent after the pre-pass, so a file is dealt to the ranks
as contiguous batches; nothing crosses GPUs on the data path.  Only the counter
block (DropInfo, quality histograms, per-100 bp and 5'/3' tables: all sums) is combined, with one
allreduce(sum) at the end of the file.  One process per GPU, `torch.distributed` for the plumbing
(NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def split_batches(offsets: np.ndarray, target_bases: int) -> List[Tuple[int, int]]:
    """Cut reads [0, n) into contiguous batches of about target_bases bases (never empty, never
    splitting a read).  Returns [(lo, hi)] read-index ranges in input order."""
    n = len(offsets) - 1
    out = []
    lo = 0
    while lo < n:
        limit = int(offsets[lo]) + max(1, int(target_bases))
        hi = int(np.searchsorted(offsets, limit, side="right")) - 1
        hi = min(n, max(hi, lo + 1))
        out.append((lo, hi))
        lo = hi
    return out


def rank_batches(batches: List[Tuple[int, int]], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Round-robin deal in input order: batch i goes to rank i % world.  Returns
    [(batch_index, lo, hi)] for this rank; output order is restored by batch_index."""
    return [(i, lo, hi) for i, (lo, hi) in enumerate(batches) if i % world == rank]


def allreduce_counters(flat, group=None):
    """Sum the uint64 counter block over all ranks.  `flat` is a torch int64 tensor viewing the
    block (uint64 sums wrap identically in two's complement).  In place; returns flat."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def device_counter_tensor(engine, device):
    """Zero-copy torch view (int64) of an engine's device-resident counter block."""
    import torch
    ptr, nwords = engine.counters_device_ptr()

    class _Block:
        __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False),
                                    "version": 3}
    return torch.as_tensor(_Block(), device=device)
