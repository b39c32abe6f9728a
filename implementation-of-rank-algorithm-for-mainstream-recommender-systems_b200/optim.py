"""Row-wise Adam for embedding tables with sparse (COO) gradients (opt-in).

`RowwiseAdam` has the arithmetic and state of `torch.optim.SparseAdam` — moments and weights of the
rows a step does not touch are left alone, the bias correction uses the global step — but one CUDA
kernel (csrc/optim.cu) updates weight, exp_avg and exp_avg_sq of the touched rows in place instead
of SparseAdam's dozen sparse-tensor kernels.  It consumes (a) coalesced sparse gradients — what
`RowShardedEmbedding` produces for the row-sharded BST table (sharded.py) and what
`nn.Embedding(sparse=True)` produces — and (b) the `touched_grad` the REPLICATED tables of every
model receive after `rank_b200.sparse.set_table_gradients("touched")`: distinct rows + summed
gradient rows with the count on the device, so no dense `[V, D]` slab is written by the backward,
no dense Adam sweeps the tables, and nothing synchronises with the host.  The reference itself uses dense `optim.Adam`; dense Adam also
decays the moments of untouched rows, so this is a different optimizer, not a drop-in.
"""
from __future__ import annotations

import torch

from . import _lib


class RowwiseAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr <= 0 or eps <= 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("RowwiseAdam: invalid lr / eps / betas")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _state(self, p):
        state = self.state[p]
        if not state:
            state["step"] = 0
            state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return state

    def _step_touched(self, lib, group, p, touched):
        """A `sparse.TouchedRows` gradient (set_table_gradients("touched")): the count stays on the device."""
        if p.dim() != 2 or tuple(touched.shape) != tuple(p.shape):
            raise RuntimeError("RowwiseAdam: touched-rows gradient does not match its table")
        w = _lib.require_cuda(p.data, "table", torch.float32)
        if w.data_ptr() != p.data.data_ptr():
            raise RuntimeError("RowwiseAdam: the table must be contiguous")
        state = self._state(p)
        state["step"] += 1
        cap = int(touched.rows.numel())
        if cap == 0:
            return
        beta1, beta2 = group["betas"]
        rc = lib.rk_rowwise_adam_touched(w.data_ptr(), state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr(),
                                         touched.rows.data_ptr(), touched.values.data_ptr(), touched.count.data_ptr(),
                                         cap, int(p.shape[1]), int(p.shape[0]), group["lr"], beta1, beta2,
                                         group["eps"], state["step"], _lib.err_flag(p.device).data_ptr(),
                                         _lib.stream_ptr())
        _lib.check(rc, "rk_rowwise_adam_touched")

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none)
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "touched_grad", None) is not None:
                    p.touched_grad = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                touched = getattr(p, "touched_grad", None)
                if p.grad is None and touched is None:
                    continue
                if touched is not None:
                    if p.grad is not None:
                        raise RuntimeError("RowwiseAdam: a table carries both .grad and .touched_grad")
                    self._step_touched(lib, group, p, touched)
                    p.touched_grad = None
                    continue
                if not p.grad.is_sparse:
                    raise RuntimeError("RowwiseAdam takes sparse gradients (rows, values); use torch.optim.Adam "
                                       "for dense ones")
                if p.dim() != 2:
                    raise RuntimeError("RowwiseAdam updates 2-D tables")
                w = _lib.require_cuda(p.data, "table", torch.float32)
                if w.data_ptr() != p.data.data_ptr():
                    raise RuntimeError("RowwiseAdam: the table must be contiguous")
                state = self._state(p)
                state["step"] += 1
                grad = p.grad.coalesce()            # one summed row per touched row, as SparseAdam requires
                rows = grad.indices()[0].contiguous()
                vals = _lib.require_cuda(grad.values(), "gradient rows", torch.float32)
                rc = lib.rk_rowwise_adam(w.data_ptr(), state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr(),
                                         rows.data_ptr(), vals.data_ptr(), int(rows.numel()), int(p.shape[1]),
                                         int(p.shape[0]), group["lr"], beta1, beta2, group["eps"], state["step"],
                                         _lib.err_flag(p.device).data_ptr(), _lib.stream_ptr())
                _lib.check(rc, "rk_rowwise_adam")
        return loss
