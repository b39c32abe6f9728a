"""Row-sharded embedding table with NCCL all-to-all (BASELINE config 5 "scaled").

The reference is single-process; its BST model keeps the whole feedid table on one device.  In the
scaled configuration (1e8 rows x 16 floats = 6.4 GB) the table is block-partitioned by row: rank r
owns rows [r*Vs, (r+1)*Vs), Vs = ceil(V / world).  One lookup of a local batch of n indices is

    owner/route kernels -> all_to_all(indices) -> local gather kernel -> all_to_all(rows)

with FIXED-CAPACITY messages: every rank sends every peer a block of n slots (the worst case: all of
its indices owned by one peer), live requests first, the rest marked dead (-1).  The split sizes are
therefore known without looking at the data: no device->host copy, no host synchronisation anywhere
in the step, and the whole exchange can be captured in a CUDA graph.  The price is world x the wire
bytes of an exact-size exchange — 10 MB per direction at 8 ranks x 20 480 indices x 64-byte rows,
~12 us of NVLink time — against two host round trips per step before.  The consumer (the first BST
block kernel) reads the received rows through the inverse permutation, so no un-permute pass exists.

The backward mirrors it: per-occurrence gradients land in the padded slots, go back to the owners
through the same all-to-all and are reduced there with the sorted segment reduction — into a dense
`[Vs, D]` gradient, or (sparse_grad=True, the only feasible choice at 1e8 rows) into touched rows:
`weight.touched_grad = sparse.TouchedRows(distinct rows, summed gradient rows, count on the device)`,
which `optim.RowwiseAdam` consumes without a host sync either.

Device-specific steps sit behind a small ops object so that the exchange logic can be exercised on
CPU with gloo in the tests (the product default is the CUDA implementation; there is no CPU path
in the product).
"""
from __future__ import annotations

import types

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .sparse import GradSource, OccurrencePlan, TouchedRows, gather_concat


class CudaShardOps:
    """The CUDA implementation of the device-side steps (csrc/shard.cu, plan_sort, segment_reduce)."""

    def route(self, idx, rows_total, rows_per_rank, world):
        """idx [n] int64 -> (send_local [n] in owner-sorted order, inv [n], counts [world]) (int64, on device)."""
        lib = _lib.load()
        idx = _lib.require_cuda(idx, "sharded index", torch.int64)
        n, dev = int(idx.numel()), idx.device
        owner = torch.empty(n, dtype=torch.int64, device=dev)
        rc = lib.rk_shard_owner(idx.data_ptr(), n, rows_total, rows_per_rank, owner.data_ptr(),
                                _lib.err_flag(dev).data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "rk_shard_owner")
        plan = OccurrencePlan([owner], [world], direct=False)
        plan.join()
        send_local = torch.empty(n, dtype=torch.int64, device=dev)
        inv = torch.empty(n, dtype=torch.int64, device=dev)
        counts = torch.empty(world, dtype=torch.int64, device=dev)
        rc = lib.rk_shard_route(idx.data_ptr(), plan.sorted_keys.data_ptr(), plan.perm.data_ptr(), n, rows_total,
                                rows_per_rank, world, send_local.data_ptr(), inv.data_ptr(), counts.data_ptr(),
                                _lib.stream_ptr())
        _lib.check(rc, "rk_shard_route")
        return send_local, inv, counts

    def gather(self, table, rows_idx):
        """table [V, D], rows_idx [m] -> [m, D]."""
        if rows_idx.numel() == 0:
            return torch.empty(0, table.shape[1], dtype=table.dtype, device=table.device)
        return gather_concat([table], [rows_idx], [0])

    def owner_plan(self, recv_key, rows_plus):
        """Sorted order of the request slots this rank serves; dead slots carry the key rows_plus - 1."""
        return OccurrencePlan([recv_key], [rows_plus], direct=False)

    def reduce(self, plan, g_rows, rows_local, dim, sparse, any_dead):
        """Per-slot gradient rows [m, D] -> dense [rows_local, D], or the touched rows of the shard."""
        if not sparse:
            (dense,) = plan.reduce_to_dense([GradSource(g_rows, 0, dim, dim, rows_local + 1, 0)])
            return dense[:rows_local]
        holder = types.SimpleNamespace(touched_grad=None)
        plan._reduce_touched([GradSource(g_rows, 0, dim, dim, rows_local + 1, 0, holder)])
        t = holder.touched_grad
        # the dead slots' sentinel row (rows_local) sorts last: drop it from the count if present
        return TouchedRows(t.rows, t.values, t.count - any_dead.to(torch.int64).reshape(1), (rows_local, dim))


class _ShardExchange(torch.autograd.Function):
    """(weight shard [Vs, D], idx [n]) -> (rows [world*n, D] in padded owner-major slots, inv [n]):
    rows[inv[p]] is the embedding row of idx[p]."""

    @staticmethod
    def forward(ctx, weight, idx, module):
        ops, group = module._ops, module.group
        world = dist.get_world_size(group)
        n, D = int(idx.numel()), int(weight.shape[1])
        dev = idx.device
        send_local, inv, counts = ops.route(idx.reshape(-1), module.num_embeddings, module.rows_per_rank, world)
        # slot of every owner-sorted request inside its peer's block of n slots (all on the device)
        offsets = torch.cumsum(counts, 0) - counts
        owner_sorted = torch.repeat_interleave(torch.arange(world, device=dev), counts, output_size=n)
        dst = owner_sorted * n + (torch.arange(n, device=dev) - offsets[owner_sorted])
        send_pad = torch.full((world * n,), -1, dtype=torch.int64, device=dev)
        send_pad[dst] = send_local
        recv_pad = torch.empty_like(send_pad)
        dist.all_to_all_single(recv_pad, send_pad, group=group)                   # equal splits: nothing to ask the host
        live = recv_pad >= 0
        served = ops.gather(weight, recv_pad.clamp(min=0))                        # rows this rank owns (dead slots: row 0, ignored)
        rows = torch.empty(world * n, D, dtype=weight.dtype, device=weight.device)
        dist.all_to_all_single(rows, served, group=group)
        inv_pad = dst[inv]
        ctx.module = module
        if weight.requires_grad:
            rows_local = int(weight.shape[0])
            recv_key = torch.where(live, recv_pad, torch.full_like(recv_pad, rows_local))
            ctx.plan = ops.owner_plan(recv_key, rows_local + 1)
            ctx.any_dead = (~live).any()
        ctx.mark_non_differentiable(inv_pad)
        return rows, inv_pad

    @staticmethod
    def backward(ctx, g_rows, _g_inv):
        module = ctx.module
        D = int(g_rows.shape[1])
        g_rows = g_rows.contiguous()
        g_served = torch.empty_like(g_rows)
        dist.all_to_all_single(g_served, g_rows, group=module.group)
        if module.grad_scale != 1.0:
            g_served = g_served * module.grad_scale
        grad = module._ops.reduce(ctx.plan, g_served, module.rows_local, D, module.sparse_grad, ctx.any_dead)
        if module.sparse_grad:
            if getattr(module.weight, "touched_grad", None) is not None:
                raise RuntimeError("a touched-rows gradient is already attached to the shard: run the optimizer "
                                   "step (or clear weight.touched_grad) between backward passes")
            module.weight.touched_grad = grad
            return None, None, None
        return grad, None, None


class RowShardedEmbedding(nn.Module):
    """`num_embeddings x embedding_dim` table, rows block-partitioned over `group`."""

    def __init__(self, num_embeddings, embedding_dim, group=None, sparse_grad=True, _ops=None):
        super().__init__()
        self.group = group
        self.num_embeddings, self.embedding_dim = int(num_embeddings), int(embedding_dim)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows_per_rank = (self.num_embeddings + world - 1) // world
        lo = min(rank * self.rows_per_rank, self.num_embeddings)
        hi = min(lo + self.rows_per_rank, self.num_embeddings)
        self.row_range = (lo, hi)
        self.rows_local = max(hi - lo, 1)
        self.sparse_grad = sparse_grad
        self.weight = nn.Parameter(torch.empty(self.rows_local, self.embedding_dim))
        nn.init.normal_(self.weight)              # nn.Embedding's initialiser
        self._ops = CudaShardOps() if _ops is None else _ops
        self.weight._rank_local = True            # GradientAllReducer leaves it alone
        # every rank's loss is a mean over its own samples; the owner sums the contributions of all
        # ranks, so 1/world gives the gradient of the mean over the global batch (what the
        # all-reduce average gives the replicated parameters)
        self.grad_scale = 1.0 / world

    @classmethod
    def from_full(cls, full_weight, group=None, sparse_grad=True, _ops=None):
        """Shard of an existing replicated table (each rank keeps its own slice)."""
        self = cls(full_weight.shape[0], full_weight.shape[1], group, sparse_grad, _ops)
        lo, hi = self.row_range
        with torch.no_grad():
            self.weight = nn.Parameter(full_weight[lo:hi].detach().clone().contiguous()
                                       if hi > lo else torch.zeros(1, full_weight.shape[1],
                                                                   device=full_weight.device))
        self.weight._rank_local = True
        return self

    def exchange(self, idx):
        """Rows of idx (any shape) as (rows [world*n, D] in padded owner-major slots, inv with idx's shape):
        `rows[inv]` is the usual embedding output; consumers that index anyway take both."""
        rows, inv = _ShardExchange.apply(self.weight, idx, self)
        return rows, inv.view(idx.shape)

    def forward(self, idx):
        rows, inv = self.exchange(idx)
        return rows[inv]


def shard_bst_feedid_table(model, group=None, sparse_grad=True):
    """Replace `model.embeddings['feedid']` of a BSTModel by its row shard; the model's forward
    then routes the sequence lookups through the all-to-all exchange."""
    full = model.embeddings["feedid"].weight
    model.embeddings["feedid"] = RowShardedEmbedding.from_full(full, group, sparse_grad)
    return model
