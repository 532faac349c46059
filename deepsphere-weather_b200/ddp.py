"""Batch-sharded data parallelism: one process per GPU, gradients all-reduced once per step.

The hot path shards over the batch only (every sample is independent through ``conv_cheb``, the
pools and the whole U-Net — SURVEY.md §8e); Laplacians, pool matrices and weights are replicated.
The only exchange is one all-reduce of the flat fp32 gradient vector (1 771 082 elements =
7.08 MB for ``UNetSpherical``) per step, over NCCL (NVLink 5 / NVSwitch) on GPUs, or gloo in the
CPU tests.  All parameter gradients are views into ONE contiguous bucket, so no pack / unpack
kernels surround the collective.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """Makes every ``p.grad`` a view into a single flat buffer and all-reduces that buffer."""

    def __init__(self, module: torch.nn.Module, process_group: Optional[dist.ProcessGroup] = None,
                 broadcast_parameters: bool = True):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.group = process_group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off : off + n].view_as(p)
            off += n
        self.world_size = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if broadcast_parameters and self.world_size > 1:
            for p in self.params:
                dist.broadcast(p.data, src=0, group=self.group)
            for b in module.buffers():
                if not b.is_sparse:
                    dist.broadcast(b.data, src=0, group=self.group)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero_(self):
        """Replaces ``optimizer.zero_grad()`` (which would detach the views when set_to_none=True)."""
        self.flat.zero_()

    def check_views(self):
        base = self.flat.untyped_storage().data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.untyped_storage().data_ptr() != base:
                raise RuntimeError("a parameter gradient was detached from the flat bucket "
                                   "(use bucket.zero_() instead of zero_grad(set_to_none=True))")

    def allreduce_mean(self, async_op: bool = False, local_batch: Optional[int] = None, global_batch: Optional[int] = None):
        """Average the per-rank gradients into the global-batch mean gradient.

        With equal shards (the default) every rank's loss is a mean over the same number of samples and the mean of the
        per-rank gradients IS the global mean: sum, then divide by the world size.  ``shard_batch`` makes uneven shards
        when the global batch does not divide; pass ``local_batch`` / ``global_batch`` then and each rank's gradient is
        weighted by its share of the samples before the sum (a mean of per-shard means would over-weight the samples of
        the smaller shards)."""
        if (local_batch is None) != (global_batch is None):
            raise ValueError("pass both local_batch and global_batch, or neither")
        if self.world_size == 1:
            return None
        self._post_scale = 1.0 / self.world_size
        if local_batch is not None:
            self.flat.mul_(float(local_batch) / float(global_batch))
            self._post_scale = 1.0
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        if self._post_scale != 1.0:
            self.flat.mul_(self._post_scale)
        return None

    def finish(self, work):
        if work is not None:
            work.wait()
            if getattr(self, "_post_scale", 1.0 / self.world_size) != 1.0:
                self.flat.mul_(getattr(self, "_post_scale", 1.0 / self.world_size))


def shard_batch(global_batch: int, rank: int, world_size: int):
    """Contiguous, near-even split of ``range(global_batch)``; returns ``(start, stop)``."""
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
