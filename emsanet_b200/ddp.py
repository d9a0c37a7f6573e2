"""Data parallelism for the EMSANet path: one process per GPU, batches sharded across ranks, and ONE exchange per
step — a mean all-reduce of the fp32 parameter gradients over NCCL (NVLink 5 / NVSwitch), SURVEY.md §8(e).

The engine keeps all parameter gradients in one flat fp32 buffer laid out in the order backward finishes them, last
first: [encoder stem + stages 1-2 | encoder stages 3-4 | context module + decoders].  It is reduced as three buckets:
[decoders + context] is launched asynchronously the moment backward crosses the encoder boundary, [encoder stages 3-4]
when backward passes the marker after stage 2 — both overlap the backward kernels that follow — and the small rest at
the end.  BatchNorm statistics stay local to each rank (the reference has no SyncBN); running buffers are not exchanged.

The reference has no distributed code at all; this is new capability behind the same nn.Module.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    """Attach with `GradAllReducer(engine)`; call `finish()` after backward (before the optimizer step)."""

    def __init__(self, engine=None, group=None, average: bool = True):
        self.group = group
        self.average = average
        self.pending: List = []
        self.ranges: List = []
        if engine is not None:
            engine.on_grads_ready = self.on_grads_ready

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def on_grads_ready(self, flat: torch.Tensor, start: int, end: int) -> None:
        """flat[start:end] is final on the current stream: start its all-reduce without blocking the host."""
        if end <= start:
            return
        self.ranges.append((start, end))
        if self.world == 1:
            return
        chunk = flat[start:end]
        if self.average and chunk.is_cuda:
            work = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self.pending.append((work, None))
        else:   # gloo has no AVG: sum, then scale when waiting
            work = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((work, chunk if self.average else None))

    def finish(self) -> None:
        """make the current stream wait for all outstanding bucket reductions"""
        for work, scale_chunk in self.pending:
            work.wait()
            if scale_chunk is not None:
                scale_chunk.mul_(1.0 / self.world)
        self.pending.clear()
        self.ranges.clear()


def shard_batch(global_batch: int, rank: int, world: int):
    """contiguous per-rank slice [lo, hi) of a global batch; sizes differ by at most one image"""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
