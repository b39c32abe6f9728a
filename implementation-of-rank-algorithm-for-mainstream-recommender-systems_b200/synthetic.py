"""Synthetic WeChat-challenge-shaped batches for tests and benchmarks.

The real parquet files are not available, so inputs are drawn with the statistics the ETL
produces (dataset/wechat_algo_data1/DataGenerator.py in the reference): Zipf-distributed ids
mapped through a fixed permutation of each vocabulary, `log1p` of count features for the dense
block, history lengths with 30 % empty histories (DIN) or 1..T (BST), zero padding past the
length as `din_collate_fn` does (DIN/din.py:210-211), label rate 0.0356.
All tensors are created on the CPU from a seeded generator; callers move them.
"""
from __future__ import annotations

import torch

from .vocab import WECHAT_VOCAB_LINES

DENSE_NAMES = (
    "videoplayseconds", "u_read_comment_7d_sum", "u_like_7d_sum", "u_click_avatar_7d_sum",
    "u_forward_7d_sum", "u_comment_7d_sum", "u_follow_7d_sum", "u_favorite_7d_sum",
    "i_read_comment_7d_sum", "i_like_7d_sum", "i_click_avatar_7d_sum", "i_forward_7d_sum",
    "i_comment_7d_sum", "i_follow_7d_sum", "i_favorite_7d_sum", "c_user_author_read_comment_7d_sum",
)
SIDE_COLUMNS = ("userid", "device", "authorid", "bgm_song_id", "bgm_singer_id", "manual_tag_list")
DEEPFM_COLUMNS = ("userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id")
LABEL_RATE = 0.0356
SEED = 20261018


def zipf_indices(gen, rows, shape, alpha=1.05, uniform=False):
    """int64 indices in [0, rows): rank ~ Zipf(alpha), rank -> row through a fixed permutation."""
    n = 1
    for s in shape:
        n *= s
    if uniform or rows <= 4:
        return torch.randint(0, rows, shape, generator=gen, dtype=torch.int64)
    # inverse-CDF sampling of a truncated Zipf over ranks 1..rows
    ranks = torch.arange(1, rows + 1, dtype=torch.float64)
    cdf = torch.cumsum(ranks.pow(-alpha), 0)
    cdf /= cdf[-1].clone()
    u = torch.rand(n, generator=gen, dtype=torch.float64)
    rank = torch.searchsorted(cdf, u).clamp_(max=rows - 1)
    perm = torch.randperm(rows, generator=torch.Generator().manual_seed(rows))
    return perm[rank].view(*shape).to(torch.int64)


def table_rows(lines=None):
    lines = WECHAT_VOCAB_LINES if lines is None else lines
    return {c: n + 1 for c, n in lines.items()}


def dense_block(gen, B):
    return torch.log1p(torch.poisson(torch.full((B, len(DENSE_NAMES)), 3.0), generator=gen))


def labels(gen, B):
    return (torch.rand(B, generator=gen) < LABEL_RATE).to(torch.float32)


def deepfm_batch(B, seed=SEED, lines=None, uniform=False):
    gen = torch.Generator().manual_seed(seed)
    rows = table_rows(lines)
    cat = {c: zipf_indices(gen, rows[c], (B,), uniform=uniform) for c in DEEPFM_COLUMNS}
    return dict(category=cat, label=labels(gen, B))


def fwfm_field_dims(lines=None):
    """`field_dims` as FwFM's main() builds them: the vocabulary lengths, no extra row (FwFM/fwfm.py:235-242)."""
    lines = WECHAT_VOCAB_LINES if lines is None else lines
    return [lines[c] for c in DEEPFM_COLUMNS]


def fwfm_batch(B, seed=SEED, lines=None, uniform=False):
    gen = torch.Generator().manual_seed(seed)
    dims = fwfm_field_dims(lines)
    x = {c: zipf_indices(gen, rows, (B,), uniform=uniform) for c, rows in zip(DEEPFM_COLUMNS, dims)}
    return dict(x=x, label=labels(gen, B))


def side_batch(B, seed=SEED, lines=None, uniform=False):
    """dense [B,16] + the six side-information columns (DCN, DeepCrossing)."""
    gen = torch.Generator().manual_seed(seed)
    rows = table_rows(lines)
    cat = {c: zipf_indices(gen, rows[c], (B,), uniform=uniform) for c in SIDE_COLUMNS}
    return dict(dense=dense_block(gen, B), category=cat, label=labels(gen, B))


def afm_feature_columns(n_fields=10, lines=None, extra_vocab=100000):
    """`feature_columns` for AFM with the 7 WeChat vocabularies plus synthetic extra fields."""
    lines = WECHAT_VOCAB_LINES if lines is None else lines
    names = list(DEEPFM_COLUMNS) + ["manual_tag_list"]
    names = names[:n_fields] + [f"extra_{i}" for i in range(max(0, n_fields - len(names)))]
    vocab = {c: range(lines[c]) if c in lines else range(extra_vocab) for c in names}
    return {"dense": list(DENSE_NAMES), "category": names, "sequence": [], "vocab": vocab}


def afm_batch(B, feature_columns, seed=SEED, uniform=False):
    gen = torch.Generator().manual_seed(seed)
    cat = {c: zipf_indices(gen, len(feature_columns["vocab"][c]) + 1, (B,), uniform=uniform)
           for c in feature_columns["category"]}
    return dict(dense=dense_block(gen, B), category=cat, label=labels(gen, B))


def din_batch(B, T=50, seed=SEED, lines=None, uniform=False, empty_rate=0.3):
    gen = torch.Generator().manual_seed(seed)
    rows = table_rows(lines)
    dense = dense_block(gen, B)
    cat = {c: zipf_indices(gen, rows[c], (B,), uniform=uniform) for c in SIDE_COLUMNS}
    length = torch.randint(1, T + 1, (B,), generator=gen, dtype=torch.int64)
    length = length * (torch.rand(B, generator=gen) >= empty_rate)
    seq = zipf_indices(gen, rows["feedid"], (B, T), uniform=uniform)
    seq = seq * (torch.arange(T).expand(B, T) < length.unsqueeze(1))
    return dict(dense={n: dense[:, i].contiguous() for i, n in enumerate(DENSE_NAMES)}, category=cat,
                sequence={"his_read_comment_7d_seq": seq, "his_read_comment_7d_seq_length": length},
                target={"feedid": zipf_indices(gen, rows["feedid"], (B,), uniform=uniform)},
                label=labels(gen, B))


def bst_batch(B, T=20, seed=SEED, lines=None, uniform=False, feed_rows=None):
    gen = torch.Generator().manual_seed(seed)
    rows = table_rows(lines)
    feed_rows = rows["feedid"] if feed_rows is None else feed_rows
    cat = {c: zipf_indices(gen, rows[c], (B,), uniform=uniform) for c in SIDE_COLUMNS}
    length = torch.randint(1, T + 1, (B,), generator=gen, dtype=torch.int64)   # 0 is NaN in the reference
    seq = zipf_indices(gen, feed_rows, (B, T), uniform=uniform or feed_rows > 10_000_000)
    seq = seq * (torch.arange(T).expand(B, T) < length.unsqueeze(1))
    return dict(dense=dense_block(gen, B), category=cat, seq_feedid=seq, seq_length=length,
                label=labels(gen, B))


def to_device(obj, device, non_blocking=False):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, dict):
        return {k: to_device(v, device, non_blocking) for k, v in obj.items()}
    return obj
