"""Row-sharded embedding table with NCCL all-to-all (BASELINE config 5 "scaled").

The reference is single-process; its BST model keeps the whole feedid table on one device.  In the
scaled configuration (1e8 rows x 16 floats = 6.4 GB) the table is block-partitioned by row: rank r
owns rows [r*Vs, (r+1)*Vs), Vs = ceil(V / world).  One lookup of a local batch is

    owner/route kernels -> all_to_all(indices) -> local gather kernel -> all_to_all(rows)

and the consumer (the first BST block kernel) reads the received rows through the inverse
permutation, so no un-permute pass exists.  The backward mirrors it: per-occurrence gradients are
permuted into send order by the gather kernel, all_to_all'ed back to the owners and reduced there
with the sorted segment reduction — into a dense `[Vs, D]` gradient, or (sparse_grad=True, the
only feasible choice at 1e8 rows) into a compact `[unique, D]` block returned as a
`torch.sparse_coo_tensor`, which `torch.optim.SparseAdam` / `SGD` consume.

The per-peer counts are needed on the host to size the all-to-all (one small D2H sync per step).
Device-specific steps sit behind a small ops object so that the exchange logic can be exercised on
CPU with gloo in the tests (the product default is the CUDA implementation; there is no CPU path
in the product).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .sparse import GradSource, OccurrencePlan, gather_concat


class CudaShardOps:
    """The CUDA implementation of the device-side steps (csrc/shard.cu, plan_sort, segment_reduce)."""

    def route(self, idx, rows_total, rows_per_rank, world):
        """idx [n] int64 -> (send_local [n], inv [n], counts [world]) (all int64, on device)."""
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

    def owner_plan(self, recv_local, rows, sparse):
        """Sorted order of the requests this rank serves; for sparse gradients also the compact
        ranks and the unique rows (host-synchronising: the unique count sizes the gradient)."""
        lib = _lib.load()
        plan = OccurrencePlan([recv_local], [rows], direct=False)
        uniq = None
        if sparse:
            plan.join()
            n, dev = int(recv_local.numel()), recv_local.device
            rank_keys = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            uniq_rows = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
            n_uniq = torch.zeros(1, dtype=torch.int64, device=dev)
            rc = lib.rk_plan_compact(plan.sorted_keys.data_ptr(), n, rows, rank_keys.data_ptr(),
                                     uniq_rows.data_ptr(), n_uniq.data_ptr(), _lib.stream_ptr())
            _lib.check(rc, "rk_plan_compact")
            u = int(n_uniq.item())
            plan.sorted_keys, plan.rows, plan.s_rows = rank_keys, [max(u, 1)], [max(u, 1)]
            uniq = uniq_rows[:u]
        return plan, uniq

    def reduce(self, plan, uniq, g_rows, rows, dim):
        """Per-request gradient rows [m, D] (request order) -> dense [rows, D] or sparse COO."""
        if uniq is None:
            (dense,) = plan.reduce_to_dense([GradSource(g_rows, 0, dim, dim, rows, 0)])
            return dense
        u = int(uniq.numel())
        if u == 0:
            return torch.sparse_coo_tensor(torch.empty(1, 0, dtype=torch.int64, device=g_rows.device),
                                           torch.empty(0, dim, device=g_rows.device), (rows, dim))
        (compact,) = plan.reduce_to_dense([GradSource(g_rows, 0, dim, dim, u, 0)])
        return torch.sparse_coo_tensor(uniq.unsqueeze(0), compact, (rows, dim), is_coalesced=True)


def _all_to_all(out, inp, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    return out


class _ShardExchange(torch.autograd.Function):
    """(weight shard [Vs, D], idx [n]) -> (rows [n, D] in owner-sorted request order, inv [n])."""

    @staticmethod
    def forward(ctx, weight, idx, module):
        ops, group = module._ops, module.group
        world = dist.get_world_size(group)
        n, D = int(idx.numel()), int(weight.shape[1])
        send_local, inv, counts = ops.route(idx.reshape(-1), module.num_embeddings, module.rows_per_rank, world)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=group)
        send_splits = counts.cpu().tolist()          # host sync: the all-to-all needs the split sizes
        recv_splits = recv_counts.cpu().tolist()
        m = int(sum(recv_splits))
        recv_local = _all_to_all(torch.empty(m, dtype=torch.int64, device=idx.device), send_local,
                                 recv_splits, send_splits, group)
        served = ops.gather(weight, recv_local)                                   # rows this rank owns
        rows = _all_to_all(torch.empty(n, D, dtype=weight.dtype, device=weight.device), served,
                           send_splits, recv_splits, group)
        ctx.module, ctx.splits, ctx.m = module, (send_splits, recv_splits), m
        if weight.requires_grad:
            ctx.plan, ctx.uniq = ops.owner_plan(recv_local, int(weight.shape[0]), module.sparse_grad)
        ctx.mark_non_differentiable(inv)
        return rows, inv

    @staticmethod
    def backward(ctx, g_rows, _g_inv):
        module = ctx.module
        send_splits, recv_splits = ctx.splits
        D = int(g_rows.shape[1])
        g_rows = g_rows.contiguous()
        g_served = _all_to_all(torch.empty(ctx.m, D, dtype=g_rows.dtype, device=g_rows.device), g_rows,
                               recv_splits, send_splits, module.group)
        if module.grad_scale != 1.0:
            g_served = g_served * module.grad_scale
        grad = module._ops.reduce(ctx.plan, ctx.uniq, g_served, module.rows_local, D)
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
        """Rows of idx (any shape) as (rows [n, D] in owner-sorted order, inv with idx's shape):
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
