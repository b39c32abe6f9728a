#!/usr/bin/env python
"""One small eager training step (forward + loss + backward) of every model / precision of the hot path, for
compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_step.py
    compute-sanitizer --tool racecheck python scripts/sanitize_step.py
Batch sizes are odd on purpose (ragged last tiles, partially filled 128-row tiles)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import rank_b200
from rank_b200 import synthetic

DEV = torch.device("cuda")
LINES = {"userid": 300, "feedid": 500, "device": 2, "authorid": 120, "bgm_song_id": 90, "bgm_singer_id": 70,
         "manual_tag_list": 30}


def step(name, model, loss_fn):
    model.to(DEV).train()
    torch.manual_seed(1)
    loss = loss_fn(model)
    loss.backward()
    torch.cuda.synchronize()
    rank_b200.check_index_errors()
    print(f"{name}: loss {float(loss):.5f}", flush=True)


def main():
    only = set(sys.argv[1:])
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(prefix="rk_san_"), LINES) + "/"
    B = 333
    side = synthetic.to_device(synthetic.side_batch(B, lines=LINES), DEV)
    cases = {}
    fm = synthetic.to_device(synthetic.deepfm_batch(B, lines=LINES), DEV)
    cases["deepfm"] = (lambda: rank_b200.DeepFM(vocab, embedding_dim=16, hidden_units=[64, 32], dropout_rate=0.0),
                       lambda m: F.binary_cross_entropy(m(fm["category"])[0].squeeze(1), fm["label"]))
    cases["dcn"] = (lambda: rank_b200.DCNModel(vocab, hidden_units=[64, 32], num_cross_layer=3),
                    lambda m: F.binary_cross_entropy(m(side["dense"], side["category"])[0].squeeze(1), side["label"]))
    cases["deepcrossing"] = (lambda: rank_b200.DeepCrossingModel(vocab, residual_internal_dim=64, residual_network_num=2),
                             lambda m: F.binary_cross_entropy(m(side["dense"], side["category"])[0].squeeze(1), side["label"]))
    fc = synthetic.afm_feature_columns(10, lines=LINES, extra_vocab=400)
    afm = synthetic.to_device(synthetic.afm_batch(B, fc), DEV)
    for prec in ("tensor", "fp32"):
        def make(prec=prec):
            m = rank_b200.AFM(fc, 32, 128)
            m.attention_precision = prec
            return m
        cases[f"afm_{prec}"] = (make, lambda m: F.binary_cross_entropy(m(afm["dense"], afm["category"])[0].squeeze(), afm["label"]))
    din = synthetic.to_device(synthetic.din_batch(B, 50, lines=LINES), DEV)
    for prec in ("fp32", "bf16"):
        for soft in (False, True):
            def make(prec=prec, soft=soft):
                m = rank_b200.DIN(vocab, hidden_units=[64, 32], dropout_rate=0.0, use_softmax=soft)
                m.activation_unit_precision = prec
                return m

            def loss(m):
                p, _, l2 = m(din["dense"], din["category"], din["sequence"], din["target"])
                return F.binary_cross_entropy(p.squeeze(), din["label"]) + l2
            cases[f"din_{prec}_{'softmax' if soft else 'raw'}"] = (make, loss)
    bst = synthetic.to_device(synthetic.bst_batch(B, 20, lines=LINES), DEV)
    for prec in ("fp32", "bf16"):
        for drop in (0.0, 0.1):
            def make(prec=prec, drop=drop):
                m = rank_b200.BSTModel(vocab, hidden_units=[64, 32], dropout_rate=drop, nhead=4, num_transformer_blocks=2,
                                       max_seq_length=20)
                m.block_precision = prec
                return m
            cases[f"bst_{prec}_p{drop}"] = (make, lambda m: F.binary_cross_entropy_with_logits(
                m(bst["dense"], bst["category"], bst["seq_feedid"], bst["seq_length"])[1].squeeze(), bst["label"]))
    for name, (make, loss) in cases.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        step(name, make(), loss)
    print("sanitize_step: done")


if __name__ == "__main__":
    main()
