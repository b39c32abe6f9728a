"""Data parallelism for the hot path: one process per GPU, batch split by sample, every rank
holds all tables.  The only exchange of a step is the gradient all-reduce.

The reference is single-process (no torch.distributed anywhere); this is new plumbing, kept to
what the path needs: all registered gradients — the dense `[V, D]` table gradients produced by
the segment reduction and the tower gradients — are packed into ONE flat fp32 buffer, reduced
with a single NCCL all-reduce over NVLink/NVSwitch (gloo on CPU in the tests), averaged, and
handed back as views of that buffer (no copy back).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, model: torch.nn.Module, group=None):
        self.group = group
        # row-sharded tables are owned by one rank each: their gradients are not replicated
        self.params = [p for p in model.parameters()
                       if p.requires_grad and not getattr(p, "_rank_local", False)]
        if not self.params:
            raise ValueError("model has no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=dtype, device=dev)
        self.views, pos = [], 0
        for p, n in zip(self.params, sizes):
            self.views.append(self.flat[pos:pos + n].view_as(p))
            pos += n

    @property
    def world_size(self):
        return dist.get_world_size(self.group)

    def allreduce(self, grads=None):
        """Average every parameter's gradient over the ranks (call after backward).  `grads`
        (default: each parameter's .grad) lets a CUDA-graph step pass its static tensors."""
        grads = [p.grad for p in self.params] if grads is None else list(grads)
        have = [(v, g) for v, g in zip(self.views, grads) if g is not None]
        missing = [v for v, g in zip(self.views, grads) if g is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v in missing:
            v.zero_()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world_size)
        for v, p in zip(self.views, self.params):
            p.grad = v
        return self.flat


def shard_batch(batch, rank: int, world: int):
    """Rows [rank*B/world, (rank+1)*B/world) of every tensor of a (nested dict) batch."""
    if torch.is_tensor(batch):
        n = batch.shape[0]
        per = n // world
        return batch[rank * per:(rank + 1) * per]
    return {k: shard_batch(v, rank, world) for k, v in batch.items()}
