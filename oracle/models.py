"""ORACLE — test infrastructure, not product code (see oracle/interactions.py for the rules).

Whole-model CPU restatements: the hot-path functions of `oracle.interactions` wired to plain
torch towers, with the reference's constructor arguments, return tuples, `state_dict` keys and
parameter-creation order, so that (a) a reference `state_dict` loads into them, (b) they are the
checker for the CUDA modules of `rank_b200`, (c) they are the "port" CPU baseline that
`bench.py --impl reference` times on the GPU box (the reference itself cannot travel there).

Per-call random weights (DCN cross layers, DeepCrossing residual units, DIN att_net) are drawn
with the same torch CPU-generator calls in the same order as the reference
(DCN/dcn.py:37-41, DeepCrossing/deepcrossing.py:37-39, DIN/din.py:61-67).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import interactions as X

_VOCAB_FILE = {"userid": "userid.txt", "feedid": "feedid.txt", "device": "device.txt",
               "authorid": "authorid.txt", "bgm_song_id": "bgm_song_id.txt",
               "bgm_singer_id": "bgm_singer_id.txt", "manual_tag_list": "manual_tag_id.txt"}
_SEVEN = tuple(_VOCAB_FILE)
_SIDE = (("userid", 16), ("device", 2), ("authorid", 4), ("bgm_song_id", 4),
         ("bgm_singer_id", 4), ("manual_tag_list", 4))


def _heights(vocab_dir, cols):
    out = {}
    for c in cols:
        path = os.path.join(vocab_dir, _VOCAB_FILE[c])
        n = 0
        if os.path.exists(path):
            with open(path) as f:
                n = len([line.strip() for line in f])
        out[c] = n + 1
    return out


class OracleDeepFM(nn.Module):          # DeepFM/deepfm.py:73-151
    def __init__(self, vocab_dir, embedding_dim=8, hidden_units=None, dropout_rate=0.1, batch_norm=True):
        super().__init__()
        hidden_units = hidden_units or [512, 256, 128]
        self.vocab_sizes = _heights(vocab_dir, [c for c in _SEVEN if c != "manual_tag_list"])
        self.first_order_embeddings = nn.ModuleDict({c: nn.Embedding(v, 1) for c, v in self.vocab_sizes.items()})
        self.second_order_embeddings = nn.ModuleDict(
            {c: nn.Embedding(v, embedding_dim) for c, v in self.vocab_sizes.items()})
        self.deep_layers = nn.ModuleList()
        width = len(self.vocab_sizes) * embedding_dim
        for u in hidden_units:
            self.deep_layers.append(nn.Linear(width, u))
            if batch_norm:
                self.deep_layers.append(nn.BatchNorm1d(u))
            self.deep_layers.append(nn.ReLU())
            if dropout_rate > 0:
                self.deep_layers.append(nn.Dropout(dropout_rate))
            width = u
        self.deep_output_layer = nn.Linear(width, 1)
        self.final_layer = nn.Linear(3, 1)

    def forward(self, category):
        cols = [c for c in self.first_order_embeddings if c in category]
        deep, first, second = X.deepfm_fm_part(
            [self.first_order_embeddings[c].weight for c in cols],
            [self.second_order_embeddings[c].weight for c in cols], [category[c] for c in cols])
        h = deep
        for layer in self.deep_layers:
            h = layer(h)
        deep_logit = self.deep_output_layer(h)
        total = self.final_layer(torch.cat([first, second, deep_logit], dim=1))
        return torch.sigmoid(total), total, first, second, deep_logit


class OracleFwFM(nn.Module):            # FwFM/fwfm.py:87-139
    COLUMNS = ("userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id")

    def __init__(self, field_dims, embed_dim):
        super().__init__()
        self.field_dims, self.num_fields, self.embed_dim = field_dims, len(field_dims), embed_dim
        self.linear = nn.ModuleList([nn.Embedding(v, 1) for v in field_dims])
        self.embedding = nn.ModuleList([nn.Embedding(v, embed_dim) for v in field_dims])
        for table in self.embedding:
            nn.init.xavier_uniform_(table.weight)
        self.num_pairs = self.num_fields * (self.num_fields - 1) // 2
        self.field_weight = nn.Parameter(torch.randn(self.num_pairs), requires_grad=True)
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        idx = [x[c] for c in self.COLUMNS[:self.num_fields]]
        z = X.fwfm_logit([t.weight for t in self.linear], [t.weight for t in self.embedding], idx,
                         self.field_weight, self.bias)
        return torch.sigmoid(z).squeeze(1)


def _side_embeddings(vocab_sizes, extra=()):
    return nn.ModuleDict({c: nn.Embedding(vocab_sizes[c], d) for c, d in _SIDE + tuple(extra)})


class OracleDCN(nn.Module):             # DCN/dcn.py:114-180
    def __init__(self, vocab_dir, hidden_units=[512, 256, 128], num_cross_layer=1):
        super().__init__()
        self.vocab_sizes = _heights(vocab_dir, _SEVEN)
        self.embeddings = _side_embeddings(self.vocab_sizes)
        self.input_dim = 16 + 34
        self.num_cross_layer = num_cross_layer
        layers, width = [], self.input_dim
        for h in hidden_units:
            layers += [nn.Linear(width, h), nn.ReLU()]
            width = h
        self.dnn = nn.Sequential(*layers)
        self.output_layer = nn.Linear(self.input_dim + hidden_units[-1], 1)

    def forward(self, dense, category):
        cols = [c for c in self.embeddings if c in category]
        x0 = X.concat_features(dense, [self.embeddings[c].weight for c in cols], [category[c] for c in cols])
        ws, bs = [], []
        frozen = getattr(self, "frozen_ephemeral", None)           # bench.py's CUDA-graph leg: last draw kept
        for _ in range(self.num_cross_layer if frozen is None else 0):   # draws of DCN/dcn.py:37-41
            w = torch.zeros(x0.shape[-1], 1)
            nn.init.xavier_normal_(w)
            ws.append(w.to(x0.dtype).to(x0.device))                # ".to(x0.device)" as DCN/dcn.py:44-45
            bs.append(torch.zeros(x0.shape[-1], 1, dtype=x0.dtype).to(x0.device))
        if frozen is not None:
            ws, bs = frozen
        self.last_ephemeral = (ws, bs)
        cross = X.cross_network(x0, ws, bs)
        logit = self.output_layer(torch.cat([cross, self.dnn(x0)], dim=1))
        return torch.sigmoid(logit), logit


class OracleDeepCrossing(nn.Module):    # DeepCrossing/deepcrossing.py:106-163
    def __init__(self, vocab_dir, residual_internal_dim=128, residual_network_num=1):
        super().__init__()
        self.vocab_sizes = _heights(vocab_dir, _SEVEN)
        self.embeddings = _side_embeddings(self.vocab_sizes)
        self.input_dim = 16 + 34
        self.residual_internal_dim = residual_internal_dim
        self.residual_network_num = residual_network_num
        self.output_layer = nn.Linear(self.input_dim, 1)

    def forward(self, dense, category):
        cols = [c for c in self.embeddings if c in category]
        x = X.concat_features(dense, [self.embeddings[c].weight for c in cols], [category[c] for c in cols])
        units = []
        frozen = getattr(self, "frozen_ephemeral", None)           # bench.py's CUDA-graph leg: last draw kept
        for _ in range(self.residual_network_num if frozen is None else 0):   # draws of deepcrossing.py:37,39
            l1 = nn.Linear(x.shape[-1], self.residual_internal_dim)
            l2 = nn.Linear(self.residual_internal_dim, x.shape[-1])
            units.append(tuple(t.detach().to(x.dtype).to(x.device) for t in (l1.weight, l1.bias, l2.weight, l2.bias)))
        if frozen is not None:
            units = frozen
        self.last_ephemeral = units
        logit = self.output_layer(X.residual_units(x, units))
        return torch.sigmoid(logit), logit


class OracleAFM(nn.Module):             # AFM/afm.py:64-119
    def __init__(self, feature_columns, embedding_dim, attention_factor):
        super().__init__()
        self.category_features = feature_columns["category"]
        self.dense_layer = nn.Linear(len(feature_columns["dense"]), 1)
        self.embeddings = nn.ModuleDict()
        for c in self.category_features:
            self.embeddings[c] = nn.Embedding(len(feature_columns["vocab"][c]) + 1, embedding_dim)
        self.attention = nn.Sequential(nn.Linear(embedding_dim, attention_factor), nn.ReLU(),
                                       nn.Linear(attention_factor, 1))
        self.p = nn.Linear(embedding_dim, 1)

    def forward(self, dense_input, category_input):
        embs = [X.gather_rows(self.embeddings[c].weight, category_input[c]) for c in self.category_features]
        pooled = X.afm_attention_pooling(embs, self.attention[0].weight, self.attention[0].bias,
                                         self.attention[2].weight, self.attention[2].bias)
        total = self.dense_layer(dense_input) + self.p(pooled)
        return torch.sigmoid(total), total


class OracleDice(nn.Module):            # DIN/din.py:26-36
    def __init__(self, num_features, eps=1e-9):
        super().__init__()
        self.eps = eps
        self.alpha = nn.Parameter(torch.zeros(num_features))
        self.bn = nn.BatchNorm1d(num_features, affine=False)

    def forward(self, x):
        p = torch.sigmoid(self.bn(x))
        return self.alpha * (1.0 - p) * x + p * x


class OracleDIN(nn.Module):             # DIN/din.py:225-323
    def __init__(self, vocab_dir, hidden_units=None, activation="dice", dropout_rate=0.1, batch_norm=True,
                 use_softmax=False, l2_lambda=0.2, mini_batch_aware_regularization=True):
        super().__init__()
        hidden_units = hidden_units or [512, 256, 128]
        self.use_softmax = use_softmax
        self.l2_lambda = l2_lambda
        self.mini_batch_aware_regularization = mini_batch_aware_regularization
        self.vocab_sizes = _heights(vocab_dir, _SEVEN)
        self.embeddings = _side_embeddings(self.vocab_sizes)
        self.embeddings["feedid"] = nn.Embedding(self.vocab_sizes["feedid"], 16)
        self.embeddings["his_read_comment_7d_seq"] = nn.Embedding(self.vocab_sizes["feedid"], 16)
        width = 16 + 34 + 16 + 16
        self.fcn = nn.ModuleList()
        for u in hidden_units:
            self.fcn.append(nn.Linear(width, u))
            self.fcn.append(OracleDice(u) if activation == "dice" else nn.PReLU())
            if batch_norm:
                self.fcn.append(nn.BatchNorm1d(u))
            if dropout_rate > 0:
                self.fcn.append(nn.Dropout(dropout_rate))
            width = u
        self.output_layer = nn.Linear(width, 1)

    def forward(self, dense, category, sequence, target):
        dense_input = torch.cat([dense[c].unsqueeze(1) for c in dense], dim=1)
        cat_rows = [X.gather_rows(e.weight, category[c]) for c, e in self.embeddings.items() if c in category]
        tgt = X.gather_rows(self.embeddings["feedid"].weight, target["feedid"])
        keys = X.gather_rows(self.embeddings["his_read_comment_7d_seq"].weight,
                             sequence["his_read_comment_7d_seq"])
        mlp = getattr(self, "frozen_ephemeral", None)               # bench.py's CUDA-graph leg: last draw kept
        if mlp is None:
            att_net = nn.Sequential(nn.Linear(4 * keys.shape[-1], 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(),
                                    nn.Linear(32, 1))               # draws of DIN/din.py:61-67 (+ .to(device) :68)
            mlp = tuple(t.detach().to(keys.dtype).to(keys.device) for t in (
                att_net[0].weight, att_net[0].bias, att_net[2].weight, att_net[2].bias, att_net[4].weight,
                att_net[4].bias))
        self.last_ephemeral = mlp
        att = X.din_local_activation(tgt, keys, sequence["his_read_comment_7d_seq_length"], mlp,
                                     self.use_softmax)
        net = torch.cat([dense_input] + cat_rows + [tgt, att], dim=1)
        for layer in self.fcn:
            net = layer(net)
        logit = self.output_layer(net)
        l2 = 0.0
        if self.mini_batch_aware_regularization and self.l2_lambda > 0:
            l2 = X.din_l2_term(self.l2_lambda, cat_rows, tgt, att)
        return torch.sigmoid(logit), logit, l2


class OracleBSTBlock(nn.Module):        # parameter container with BSTTransformer's names (BST/bst.py:42-64)
    def __init__(self, d_model, nhead, max_len, dropout=0.1):
        super().__init__()
        self.nhead = nhead
        self.position_embedding = nn.Embedding(max_len, d_model)
        self.w_q = nn.Linear(d_model, d_model)
        self.w_k = nn.Linear(d_model, d_model)
        self.w_v = nn.Linear(d_model, d_model)
        self.w_o = nn.Linear(d_model, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_model), nn.LeakyReLU(negative_slope=0.01),
                                 nn.Dropout(dropout), nn.Linear(d_model, d_model))


class OracleBST(nn.Module):             # BST/bst.py:162-247
    def __init__(self, vocab_dir, hidden_units=[512, 256, 128], dropout_rate=0.1, batch_norm=True, d_model=16,
                 nhead=4, num_transformer_blocks=1, max_seq_length=50, pooling_method="sum"):
        super().__init__()
        if dropout_rate:
            raise ValueError("OracleBST is deterministic: dropout_rate must be 0")
        self.vocab_sizes = _heights(vocab_dir, _SEVEN)
        self.embeddings = _side_embeddings(self.vocab_sizes, extra=(("feedid", 16),))
        self.nhead = nhead
        self.transformer_blocks = nn.ModuleList(
            [OracleBSTBlock(16, nhead, max_seq_length + 1, dropout_rate) for _ in range(num_transformer_blocks)])
        self.pooling_method = pooling_method
        layers, width = [], 16 + 34 + 16
        for h in hidden_units:
            layers.append(nn.Linear(width, h))
            if batch_norm:
                layers.append(nn.BatchNorm1d(h))
            layers.append(nn.LeakyReLU(negative_slope=0.01))
            width = h
        layers.append(nn.Linear(width, 1))
        self.dnn = nn.Sequential(*layers)

    def forward(self, dense, category, seq_feedid, seq_length):
        cols = [c for c in self.embeddings if c in category]
        cat = torch.cat([X.gather_rows(self.embeddings[c].weight, category[c]) for c in cols], dim=1)
        seq_rows = X.gather_rows(self.embeddings["feedid"].weight, seq_feedid)
        blocks = [dict(b.named_parameters()) for b in self.transformer_blocks]
        pooled = X.bst_sequence_feature(seq_rows, seq_length, blocks, self.nhead, self.pooling_method)
        logits = self.dnn(torch.cat([dense, cat, pooled], dim=1))
        return torch.sigmoid(logits), logits
