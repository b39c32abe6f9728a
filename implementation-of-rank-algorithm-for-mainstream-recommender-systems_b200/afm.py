"""AFM with the gathers and the pairwise-interaction attention pooling on the B200 hot path.

Drop-in for the reference's `AFM` and `create_feature_columns` (AFM/afm.py:64-156): same
constructor `(feature_columns, embedding_dim, attention_factor)`, same
`forward(dense_input, category_input) -> (prediction, total_logit)`, same `state_dict` keys.
One kernel (csrc/afm.cu) gathers the F rows, forms the F(F-1)/2 Hadamard pairs, runs the
attention net, the softmax over the pairs and the weighted sum without ever writing the
[B,P,D] / [B,P,A] intermediates to HBM; the backward (one persistent kernel + a fixed-order
reduction of the per-CTA partials) also returns the gradients of the registered attention
weights.  `dense_layer` and `p` stay torch.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from .sparse import GradSource, OccurrencePlan, field_array


def _check_precision(name):
    if name not in ("tensor", "bf16", "fp32"):
        raise ValueError(f"AFM.attention_precision must be 'tensor' or 'fp32', got {name!r}")
    return name


class _AfmPooling(torch.autograd.Function):
    """(w1, b1, w2, b2, idx_0.., table_0..) -> pooled[B, D]."""

    @staticmethod
    def forward(ctx, cfg, w1, b1, w2, b2, *args):
        lib = _lib.load()
        F, precision = cfg
        idx, tables = args[:F], args[F:2 * F]
        D = int(tables[0].shape[1])
        fields, keep = field_array(tables, idx, [f * D for f in range(F)])
        w1 = _lib.require_cuda(w1, "attention.0.weight", torch.float32)
        b1 = _lib.require_cuda(b1, "attention.0.bias", torch.float32)
        w2 = _lib.require_cuda(w2, "attention.2.weight", torch.float32)
        b2 = _lib.require_cuda(b2, "attention.2.bias", torch.float32)
        A = int(w1.shape[0])
        if w1.shape != (A, D) or b1.shape != (A,) or w2.numel() != A or b2.numel() != 1:
            raise ValueError("attention weights do not match (embedding_dim, attention_factor)")
        B = int(idx[0].shape[0])
        dev = w1.device
        out = torch.empty(B, D, dtype=torch.float32, device=dev)
        tiles = None
        if precision in ("tensor", "bf16"):
            # scratch for the weights' operand tiles: made once by the forward's prologue, fetched by bulk (TMA)
            # copy in every CTA of the forward and of the backward
            tiles = torch.empty(lib.rk_afm_tile_bytes(), dtype=torch.uint8, device=dev)
            rc = lib.rk_afm_tc_fwd(fields, F, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), A, B,
                                   out.data_ptr(), tiles.data_ptr(), _lib.err_flag(dev).data_ptr(), _lib.stream_ptr())
        else:
            rc = lib.rk_afm_fwd(fields, F, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), A, B,
                                out.data_ptr(), _lib.err_flag(dev).data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "rk_afm_fwd")
        if _lib.CHECK_EVERY_CALL:
            _lib.check_index_errors(dev)
        ctx.set_materialize_grads(False)
        if any(ctx.needs_input_grad):
            ctx.meta = (F, D, A, B, [int(t.shape[0]) for t in tables])
            ctx.precision = precision
            ctx.fields, ctx.keep, ctx.tiles = fields, keep, tiles
            if any(ctx.needs_input_grad[5 + F:]):
                ctx.plan = OccurrencePlan([keep[2 * f + 1] for f in range(F)], ctx.meta[4])
                ctx.tables = list(tables)
            ctx.save_for_backward(w1, b1, w2, b2)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        F, D, A, B, rows = ctx.meta
        n_in = 5 + 2 * F
        if g_out is None:
            return (None,) * n_in
        w1, b1, w2, b2 = ctx.saved_tensors
        dev = w1.device
        g_out = _lib.require_cuda(g_out, "g_pooled", torch.float32)
        g_rows = torch.empty(B, F * D, dtype=torch.float32, device=dev)
        g_att = torch.empty(A * D + 2 * A + 1, dtype=torch.float32, device=dev)   # w1 | b1 | w2 | b2
        tc = ctx.precision in ("tensor", "bf16")
        n_ctas = (lib.rk_afm_tc_bwd_ctas if tc else lib.rk_afm_bwd_ctas)(B, F)
        partials = torch.empty(n_ctas * g_att.numel(), dtype=torch.float32, device=dev)
        base = g_att.data_ptr()
        common = (ctx.fields, F, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), A, B,
                  g_out.data_ptr(), g_rows.data_ptr(), base, base + 4 * A * D,
                  base + 4 * (A * D + A), base + 4 * (A * D + 2 * A), partials.data_ptr(), n_ctas)
        if tc:
            rc = lib.rk_afm_tc_bwd(*common, _lib.ptr(ctx.tiles), _lib.err_flag(dev).data_ptr(), _lib.stream_ptr())
        else:
            rc = lib.rk_afm_bwd(*common, _lib.err_flag(dev).data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "rk_afm_bwd")
        g_tables = [None] * F
        if any(ctx.needs_input_grad[5 + F:]):
            g_tables = ctx.plan.reduce_to_dense(
                [GradSource(g_rows, f * D, F * D, D, rows[f], f, ctx.tables[f]) for f in range(F)])
        return (None, g_att[:A * D].view(A, D), g_att[A * D:A * D + A], g_att[A * D + A:A * D + 2 * A].view(1, A),
                g_att[A * D + 2 * A:], *([None] * F), *g_tables)


class AFM(nn.Module):
    def __init__(self, feature_columns, embedding_dim, attention_factor):
        super().__init__()
        self.feature_columns = feature_columns
        self.embedding_dim = embedding_dim
        self.attention_factor = attention_factor
        self.dense_features = feature_columns['dense']
        self.num_dense = len(self.dense_features)
        self.dense_layer = nn.Linear(self.num_dense, 1)
        self.category_features = feature_columns['category']
        self.embeddings = nn.ModuleDict()
        for col in self.category_features:
            self.embeddings[col] = nn.Embedding(len(feature_columns['vocab'][col]) + 1, embedding_dim)
        self.num_fields = len(self.category_features)
        self.attention = nn.Sequential(
            nn.Linear(embedding_dim, attention_factor), nn.ReLU(), nn.Linear(attention_factor, 1))
        self.p = nn.Linear(embedding_dim, 1)
        # "tensor" (default): the attention MLP on tcgen05 (csrc/afm_tc.cu) — split-bf16 operands, fp32
        # accumulation, ReLU decisions re-checked in fp32 — inside the 1e-5 bar against the reference;
        # "fp32": the SIMT kernels (csrc/afm.cu).  ("bf16" is accepted as an alias of "tensor".)  Not a parameter.
        self.attention_precision = "tensor"

    def hot_path(self, dense_input, category_input):
        """The part of forward that runs in librank_b200: the attention-pooled pair interactions [B, D]."""
        cols = self.category_features
        return (_AfmPooling.apply(
            (len(cols), _check_precision(self.attention_precision)), self.attention[0].weight, self.attention[0].bias, self.attention[2].weight,
            self.attention[2].bias, *[category_input[c] for c in cols],
            *[self.embeddings[c].weight for c in cols]),)

    def forward(self, dense_input, category_input):
        dense_logit = self.dense_layer(dense_input)
        (pooled,) = self.hot_path(dense_input, category_input)
        total_logit = dense_logit + self.p(pooled)
        prediction = torch.sigmoid(total_logit)
        return prediction, total_logit


DENSE_COLUMNS = [
    "videoplayseconds", "u_read_comment_7d_sum", "u_like_7d_sum", "u_click_avatar_7d_sum",
    "u_forward_7d_sum", "u_comment_7d_sum", "u_follow_7d_sum", "u_favorite_7d_sum",
    "i_read_comment_7d_sum", "i_like_7d_sum", "i_click_avatar_7d_sum", "i_forward_7d_sum",
    "i_comment_7d_sum", "i_follow_7d_sum", "i_favorite_7d_sum", "c_user_author_read_comment_7d_sum",
]


def create_feature_columns(vocabulary_dir):
    """The feature-column description the reference builds for the WeChat data
    (AFM/afm.py:121-156): 16 dense columns, 7 category columns, their vocabularies read from
    `<column>.txt` (manual_tag_list -> manual_tag_id.txt; a missing file is an empty vocabulary)."""
    category = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id", "manual_tag_list"]
    renamed = {"manual_tag_list": "manual_tag_id"}
    feature_columns = {'dense': list(DENSE_COLUMNS), 'category': category, 'sequence': [], 'vocab': {}}
    for col in category:
        path = os.path.join(vocabulary_dir, f"{renamed.get(col, col)}.txt")
        if os.path.exists(path):
            with open(path, 'r') as f:
                feature_columns['vocab'][col] = [line.strip() for line in f if line.strip()]
            print(f"Loaded vocabulary for {col} from {path}")
        else:
            print(f"Warning: Vocabulary file not found for {col} (searched {path})")
            feature_columns['vocab'][col] = []
    return feature_columns, ["read_comment"]
