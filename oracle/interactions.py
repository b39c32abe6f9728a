"""ORACLE — test infrastructure, not product code.

CPU restatement (torch fp32 tensors, autograd for the backward; numpy for the integer/index
work) of the reference's hot path, one function per interaction, each citing the reference
lines it follows (paths relative to /root/reference/algorithm).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this package; the product (`rank_b200`) never does.

Parity pin: the reference ships no golden vectors for this path (its two `__main__` smokes only
print).  The oracle is pinned against outputs of the reference itself, generated in the
authoring container by importing the unmodified reference classes
(`tests/golden/make_golden.py` -> `tests/golden/*.pt`) and checked by
`tests/test_oracle_golden.py`; when /root/reference is present the same test also runs the
reference live next to the oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- index / integer work
def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """nn.Embedding lookup: rows of `table` picked by int64 `idx` (any shape), bit-exact.
    Follows the per-field lookups, e.g. DeepFM/deepfm.py:123-132, DCN/dcn.py:163-167."""
    if idx.dtype != torch.int64:
        raise TypeError("indices are torch.long in the reference")
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= table.shape[0]):
        raise IndexError("index out of range in self")
    return table[idx]


def dense_embedding_grad(idx: np.ndarray, g_rows: np.ndarray, rows: int) -> np.ndarray:
    """embedding_dense_backward: out[v] = sum of g_rows[i] over occurrences i with idx[i] == v,
    added in occurrence order (what autograd produces for the default sparse=False
    nn.Embedding of every reference model, e.g. DIN/din.py:251-260 via loss.backward())."""
    idx = np.asarray(idx).reshape(-1)
    g = np.asarray(g_rows, dtype=np.float32).reshape(idx.shape[0], -1)
    out = np.zeros((rows, g.shape[1]), dtype=np.float32)
    for i in range(idx.shape[0]):          # sequential on purpose: defines the summation order
        out[idx[i]] += g[i]
    return out


def stable_occurrence_order(idx_columns, rows, dead=None):
    """(sorted_keys, perm) of every occurrence by (field, row), ties in occurrence order —
    the order rk_plan_build must produce.  perm = occurrence number inside its field.  Every
    field owns rows+1 keys: the last one is the sentinel of `dead` occurrences (padded history
    positions, whose gradient is identically zero), which therefore sort last in their field."""
    keys, perm, base = [], [], 0
    for f, (col, r) in enumerate(zip(idx_columns, rows)):
        col = np.asarray(col).reshape(-1).astype(np.int64).copy()
        if dead is not None and dead[f] is not None:
            col[np.asarray(dead[f]).reshape(-1)] = r
        keys.append(col + base)
        perm.append(np.arange(col.shape[0], dtype=np.int64))
        base += r + 1
    keys = np.concatenate(keys)
    perm = np.concatenate(perm)
    order = np.argsort(keys, kind="stable")
    return keys[order].astype(np.uint32), perm[order].astype(np.uint32)


# --------------------------------------------------------------------------- DeepFM
def deepfm_fm_part(first_tables, second_tables, idx_columns):
    """DeepFM/deepfm.py:121-142.  Returns (deep_input[B,F*D], first[B,1], second[B,1])."""
    firsts = [gather_rows(t, i) for t, i in zip(first_tables, idx_columns)]        # :123-126
    fm_first = torch.sum(torch.cat(firsts, dim=1), dim=1, keepdim=True)            # :127
    embs = [gather_rows(t, i) for t, i in zip(second_tables, idx_columns)]         # :129-132
    stacked = torch.stack(embs, dim=1)
    sum_then_square = torch.square(torch.sum(stacked, dim=1))                      # :134-135
    square_then_sum = torch.sum(torch.square(stacked), dim=1)                      # :137-138
    fm_second = 0.5 * torch.sum(sum_then_square - square_then_sum, dim=1, keepdim=True)  # :140
    deep_input = torch.cat(embs, dim=1)                                            # :142
    return deep_input, fm_first, fm_second


def fwfm_logit(first_tables, second_tables, idx_columns, field_weight, bias):
    """FwFM/fwfm.py:118-137.  z[B,1] = sum_f w_f + sum_{i<j} r_p <e_i, e_j> + bias, the pairs p
    enumerated i outer / j inner as the reference's double loop (:126-135)."""
    linear_sum = sum(gather_rows(t, i) for t, i in zip(first_tables, idx_columns))      # :122-123
    embs = torch.stack([gather_rows(t, i) for t, i in zip(second_tables, idx_columns)], dim=1)   # [B,F,D]
    F = embs.shape[1]
    left, right = torch.triu_indices(F, F, offset=1)          # row-major upper triangle = loop order
    dots = (embs[:, left, :] * embs[:, right, :]).sum(dim=2)                            # [B,P]  :131
    quadratic = (dots * field_weight).sum(dim=1, keepdim=True)                          # :133
    return linear_sum + quadratic + bias                                                # :137


# --------------------------------------------------------------------------- DCN / DeepCrossing
def concat_features(dense, tables, idx_columns):
    """[dense | e_0 | e_1 ...]: DCN/dcn.py:163-169, DeepCrossing/deepcrossing.py:148-155."""
    return torch.cat([dense] + [gather_rows(t, i) for t, i in zip(tables, idx_columns)], dim=1)


def cross_network(x0, ws, bs):
    """L applications of cross_layer (DCN/dcn.py:25-50,171-173):
    x_{l+1} = x0 * (x_l @ w_l) + b_l^T + x_l, with w_l, b_l of shape [d, 1]."""
    xl = x0
    for w, b in zip(ws, bs):
        xl = x0 * torch.matmul(xl, w) + b.t() + xl
    return xl


def residual_units(x, units):
    """N applications of residual_unit (DeepCrossing/deepcrossing.py:25-42,157-159):
    relu(x + W2 relu(W1 x + b1) + b2); `units` = [(W1[h,d], b1[h], W2[d,h], b2[d]), ...]."""
    for w1, b1, w2, b2 in units:
        h = torch.relu(F.linear(x, w1, b1))
        x = torch.relu(x + F.linear(h, w2, b2))
    return x


# --------------------------------------------------------------------------- AFM
def afm_attention_pooling(embs, w1, b1, w2, b2):
    """AFM/afm.py:101-113: Hadamard product of every field pair (i<j, row-major), attention MLP
    Linear(D,A)->ReLU->Linear(A,1) (:84-88), softmax over the pairs, weighted sum -> [B, D]."""
    pairs = []
    n = len(embs)
    for i in range(n):
        for j in range(i + 1, n):
            pairs.append(embs[i] * embs[j])
    pairs = torch.stack(pairs, dim=1)                                   # [B, P, D]
    scores = F.linear(torch.relu(F.linear(pairs, w1, b1)), w2, b2)      # [B, P, 1]
    weights = torch.softmax(scores, dim=1)
    return torch.sum(pairs * weights, dim=1)


# --------------------------------------------------------------------------- DIN
def din_local_activation(query, keys, keys_length, mlp, use_softmax):
    """din_attention (DIN/din.py:42-84).  mlp = (W1[64,4D], b1, W2[32,64], b2, W3[1,32], b3),
    the weights att_net would have drawn.  query [B,D], keys [B,T,D], keys_length [B] int64."""
    w1, b1, w2, b2, w3, b3 = mlp
    B, T, D = keys.shape
    q = query.unsqueeze(1).expand_as(keys)
    cross = torch.cat([q, keys, q - keys, q * keys], dim=2)                        # :59
    score = F.linear(torch.relu(F.linear(torch.relu(F.linear(cross, w1, b1)), w2, b2)), w3, b3)
    score = score.squeeze(2)                                                       # :69
    mask = torch.arange(T, device=keys.device).expand(B, T) < keys_length.unsqueeze(1)                 # :71
    if use_softmax:
        pad = torch.ones_like(score) * (-2 ** 32 + 1)                              # :74
        score = torch.where(mask, score, pad) / (D ** 0.5)                         # :75-76
        weight = torch.softmax(score, dim=1)                                       # :77
    else:
        weight = score.masked_fill(~mask, 0.0)                                     # :80
    return torch.sum(weight.unsqueeze(2) * keys, dim=1)                            # :82-83


def din_l2_term(l2_lambda, category_rows, target_row, attention_out):
    """Mini-batch-aware regulariser of DIN/din.py:318-322: lambda * mean_b ||[cat|target|att]||_2."""
    v = torch.cat([torch.cat(category_rows, dim=1), target_row, attention_out], dim=1)
    return l2_lambda * torch.norm(v, p=2, dim=1).mean()


# --------------------------------------------------------------------------- BST
def bst_transformer_block(x, pad_mask, p, nhead, dropout_p=0.0, keep=None):
    """BSTTransformer.forward (BST/bst.py:66-91) with queries = keys = values = x [B,T,d].
    `p` maps the block's state_dict names to tensors; pad_mask [B,T] is True on padded keys.
    Dropout (BST/bst.py:57,62,86,90) is deterministic here: `keep` [3,B,T,d] holds the keep masks
    of the three sites (w_o output :86, inside the FFN :62, FFN output :90) and dropout_p the rate;
    nn.Dropout's arithmetic is x * keep / (1 - p).  keep=None requires dropout_p == 0."""
    if dropout_p and keep is None:
        raise ValueError("the oracle is deterministic: pass the keep masks for dropout_p > 0")
    B, T, d = x.shape

    def drop(v, site):
        if keep is None or not dropout_p:
            return v
        return v * keep[site].to(v.dtype) * (1.0 / (1.0 - dropout_p))
    pos = p["position_embedding.weight"][torch.arange(T, device=p["position_embedding.weight"].device)]                          # :68-69
    qk_in = x + pos                                                                # :70-71 (not values)
    q = F.linear(qk_in, p["w_q.weight"], p["w_q.bias"]).view(B, T, nhead, -1).transpose(1, 2)
    k = F.linear(qk_in, p["w_k.weight"], p["w_k.bias"]).view(B, T, nhead, -1).transpose(1, 2)
    v = F.linear(x, p["w_v.weight"], p["w_v.bias"]).view(B, T, nhead, -1).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.size(-1))          # :77
    scores = scores.masked_fill(pad_mask.unsqueeze(1).unsqueeze(2), float("-inf"))  # :79-80
    ctx = torch.matmul(torch.softmax(scores, dim=-1), v)                           # :82-83
    ctx = ctx.transpose(1, 2).contiguous().view(B, T, -1)                          # :84
    o1 = F.layer_norm(qk_in + drop(F.linear(ctx, p["w_o.weight"], p["w_o.bias"]), 0), (d,),
                      p["norm1.weight"], p["norm1.bias"])                          # :86
    hidden = drop(F.leaky_relu(F.linear(o1, p["ffn.0.weight"], p["ffn.0.bias"]), 0.01), 1)   # :59-62
    ffn = F.linear(hidden, p["ffn.3.weight"], p["ffn.3.bias"])                     # :63,88
    return F.layer_norm(o1 + drop(ffn, 2), (d,), p["norm2.weight"], p["norm2.bias"])          # :90


def bst_sequence_feature(seq_rows, seq_length, blocks, nhead, pooling):
    """BSTModel.forward sequence branch (BST/bst.py:224-241): key-padding mask t >= len, the
    transformer blocks, then sum (or sum / len) over ALL T positions."""
    B, T, _ = seq_rows.shape
    mask = torch.arange(T, device=seq_length.device).expand(B, T) >= seq_length.unsqueeze(1)                 # :226-227
    out = seq_rows
    for p in blocks:
        out = bst_transformer_block(out, mask, p, nhead)
    pooled = torch.sum(out, dim=1)
    if pooling != "sum":
        pooled = pooled / seq_length.unsqueeze(1).float()                          # :241
    return pooled
