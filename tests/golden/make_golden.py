"""Generates tests/golden/*.pt by running the UNMODIFIED reference classes on CPU.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py
The reference ships no expected values for its hot path, so these recorded input/output/gradient
sets are the pin for `oracle/` (tests/test_oracle_golden.py) and, through it, for the CUDA path.
Protocol: small vocabularies (so duplicate indices are common), dropout 0, model.train(),
`torch.manual_seed(seed)` immediately before the forward (per-call random weights of DCN /
DeepCrossing / DIN are drawn from the CPU generator inside forward), loss = sum_k <out_k, cot_k>.
"""
import importlib.util
import os
import sys
import tempfile

import torch

REF = os.environ.get("RANK_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True

VOCAB_FILE = {"userid": "userid.txt", "feedid": "feedid.txt", "device": "device.txt",
              "authorid": "authorid.txt", "bgm_song_id": "bgm_song_id.txt",
              "bgm_singer_id": "bgm_singer_id.txt", "manual_tag_list": "manual_tag_id.txt"}
SMALL_LINES = {"userid": 40, "feedid": 60, "device": 2, "authorid": 30, "bgm_song_id": 25,
               "bgm_singer_id": 20, "manual_tag_list": 12}
DENSE_NAMES = ["videoplayseconds", "u_read_comment_7d_sum", "u_like_7d_sum", "u_click_avatar_7d_sum",
               "u_forward_7d_sum", "u_comment_7d_sum", "u_follow_7d_sum", "u_favorite_7d_sum",
               "i_read_comment_7d_sum", "i_like_7d_sum", "i_click_avatar_7d_sum", "i_forward_7d_sum",
               "i_comment_7d_sum", "i_follow_7d_sum", "i_favorite_7d_sum",
               "c_user_author_read_comment_7d_sum"]


def load_reference(rel_path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "algorithm", rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_vocab(path, lines):
    os.makedirs(path, exist_ok=True)
    for col, n in lines.items():
        with open(os.path.join(path, VOCAB_FILE[col]), "w") as f:
            f.write("".join(f"{col}_{i}\n" for i in range(1, n + 1)))
    return path


def rand_idx(gen, rows, shape):
    idx = torch.randint(0, rows, shape, generator=gen, dtype=torch.int64)
    flat = idx.view(-1)
    flat[0] = 0                    # the unknown/padding row is a trained row too
    flat[-1] = rows - 1
    return idx


def record(model, fwd, seed, extra):
    """Run forward+backward; returns the fixture dict."""
    model.train()
    torch.manual_seed(seed)
    outs = fwd()
    tensors = [o for o in outs if torch.is_tensor(o)]
    gen = torch.Generator().manual_seed(seed + 1)
    cots = [torch.randn(o.shape, generator=gen) for o in tensors]
    loss = sum((o * c).sum() for o, c in zip(tensors, cots))
    model.zero_grad()
    loss.backward()
    fx = dict(extra)
    fx.update(seed=seed,
              state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
              outputs=[o.detach().clone() if torch.is_tensor(o) else o for o in outs],
              cotangents=cots,
              grads={k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
    return fx


def save(fx, name):
    """`python make_golden.py fwfm` rewrites only the files whose name starts with `fwfm` (the others
    are still computed, so the generator stream is the same as in a full run)."""
    only = sys.argv[1:]
    if not only or any(name.startswith(o) for o in only):
        torch.save(fx, os.path.join(HERE, name))


def main():
    tmp = tempfile.mkdtemp(prefix="rk_golden_vocab_")
    vocab_dir = write_vocab(tmp, SMALL_LINES) + "/"
    rows = {c: n + 1 for c, n in SMALL_LINES.items()}
    B = 24
    gen = torch.Generator().manual_seed(20261018)
    side = ["userid", "device", "authorid", "bgm_song_id", "bgm_singer_id", "manual_tag_list"]
    dense = torch.log1p(torch.poisson(torch.full((B, 16), 3.0), generator=gen))

    # ---- DeepFM
    ref = load_reference("DeepFM/deepfm.py", "ref_deepfm")
    torch.manual_seed(1)
    for D in (16, 6):
        m = ref.DeepFM(vocab_dir, embedding_dim=D, hidden_units=[32, 16], dropout_rate=0.0, batch_norm=True)
        cat = {c: rand_idx(gen, rows[c], (B,)) for c in
               ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]}
        fx = record(m, lambda: m(cat), 7, dict(
            model="DeepFM", vocab_lines=SMALL_LINES,
            ctor=dict(embedding_dim=D, hidden_units=[32, 16], dropout_rate=0.0, batch_norm=True),
            inputs=dict(category=cat)))
        save(fx, f"deepfm_d{D}.pt")

    # ---- DCN
    ref = load_reference("DCN/dcn.py", "ref_dcn")
    m = ref.DCNModel(vocab_dir, hidden_units=[32, 16], num_cross_layer=3)
    cat = {c: rand_idx(gen, rows[c], (B,)) for c in side}
    fx = record(m, lambda: m(dense, cat), 11, dict(
        model="DCNModel", vocab_lines=SMALL_LINES, ctor=dict(hidden_units=[32, 16], num_cross_layer=3),
        inputs=dict(dense=dense, category=cat)))
    save(fx, "dcn_l3.pt")

    # ---- DeepCrossing
    ref = load_reference("DeepCrossing/deepcrossing.py", "ref_dc")
    m = ref.DeepCrossingModel(vocab_dir, residual_internal_dim=24, residual_network_num=2)
    fx = record(m, lambda: m(dense, cat), 13, dict(
        model="DeepCrossingModel", vocab_lines=SMALL_LINES,
        ctor=dict(residual_internal_dim=24, residual_network_num=2),
        inputs=dict(dense=dense, category=cat)))
    save(fx, "deepcrossing_n2.pt")

    # ---- AFM (10 fields: the 7 shipped + 3 synthetic ones, as BASELINE config 3)
    ref = load_reference("AFM/afm.py", "ref_afm")
    fields = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id", "manual_tag_list",
              "extra_a", "extra_b", "extra_c"]
    vocab = {c: [str(i) for i in range(SMALL_LINES.get(c, 15))] for c in fields}
    fc = {"dense": DENSE_NAMES, "category": fields, "sequence": [], "vocab": vocab}
    torch.manual_seed(3)
    m = ref.AFM(fc, embedding_dim=8, attention_factor=12)
    cat10 = {c: rand_idx(gen, len(vocab[c]) + 1, (B,)) for c in fields}
    fx = record(m, lambda: m(dense, cat10), 17, dict(
        model="AFM", ctor=dict(embedding_dim=8, attention_factor=12),
        feature_columns=fc, inputs=dict(dense=dense, category=cat10)))
    save(fx, "afm_f10.pt")

    # ---- DIN (both attention modes; lengths include 0 and T)
    ref = load_reference("DIN/din.py", "ref_din")
    T = 7
    length = torch.randint(0, T + 1, (B,), generator=gen, dtype=torch.int64)
    length[0], length[1] = 0, T
    seq = rand_idx(gen, rows["feedid"], (B, T))
    seq = seq * (torch.arange(T).expand(B, T) < length.unsqueeze(1))   # din_collate_fn pads with 0
    dense_dict = {n: dense[:, i].clone() for i, n in enumerate(DENSE_NAMES)}
    target = {"feedid": rand_idx(gen, rows["feedid"], (B,))}
    sequence = {"his_read_comment_7d_seq": seq, "his_read_comment_7d_seq_length": length}
    for soft in (False, True):
        torch.manual_seed(5)
        m = ref.DIN(vocab_dir, hidden_units=[32, 16], activation="dice", dropout_rate=0.0, batch_norm=True,
                    use_softmax=soft, l2_lambda=0.2, mini_batch_aware_regularization=True)
        fx = record(m, lambda: m(dense_dict, cat, sequence, target), 19, dict(
            model="DIN", vocab_lines=SMALL_LINES,
            ctor=dict(hidden_units=[32, 16], activation="dice", dropout_rate=0.0, batch_norm=True,
                      use_softmax=soft, l2_lambda=0.2, mini_batch_aware_regularization=True),
            inputs=dict(dense=dense_dict, category=cat, sequence=sequence, target=target)))
        save(fx, f"din_softmax{int(soft)}.pt")

    # the reference's own smoke input for din_attention (DIN/din_attention.py:54-68)
    ref_att = load_reference("DIN/din_attention.py", "ref_din_attention")
    torch.manual_seed(42)
    keys = torch.randn(2, 3, 4)
    query = torch.randn(2, 4)
    klen = torch.tensor([0, 1])
    state = torch.get_rng_state()
    out_raw = ref_att.din_attention(query, keys, klen, is_softmax=False)
    state2 = torch.get_rng_state()
    out_soft = ref_att.din_attention(query, keys, klen, is_softmax=True)
    torch.save(dict(keys=keys, query=query, keys_length=klen, rng_before_raw=state, rng_before_softmax=state2,
                    out_raw=out_raw.detach(), out_softmax=out_soft.detach()),
               os.path.join(HERE, "din_attention_smoke.pt"))

    # ---- BST (lengths >= 1: length 0 is NaN in the reference)
    ref = load_reference("BST/bst.py", "ref_bst")
    Tb = 6
    blen = torch.randint(1, Tb + 1, (B,), generator=gen, dtype=torch.int64)
    blen[0], blen[1] = 1, Tb
    bseq = rand_idx(gen, rows["feedid"], (B, Tb))
    bseq = bseq * (torch.arange(Tb).expand(B, Tb) < blen.unsqueeze(1))
    for nhead, blocks, pool in ((4, 1, "sum"), (2, 2, "mean")):
        torch.manual_seed(9)
        m = ref.BSTModel(vocab_dir, hidden_units=[32, 16], dropout_rate=0.0, batch_norm=True, nhead=nhead,
                         num_transformer_blocks=blocks, max_seq_length=Tb, pooling_method=pool)
        fx = record(m, lambda: m(dense, cat, bseq, blen), 23, dict(
            model="BSTModel", vocab_lines=SMALL_LINES,
            ctor=dict(hidden_units=[32, 16], dropout_rate=0.0, batch_norm=True, nhead=nhead,
                      num_transformer_blocks=blocks, max_seq_length=Tb, pooling_method=pool),
            inputs=dict(dense=dense, category=cat, seq_feedid=bseq, seq_length=blen)))
        save(fx, f"bst_h{nhead}_b{blocks}_{pool}.pt")
    # ---- FwFM: field_dims are vocabulary lengths, forward returns one tensor
    ref = load_reference("FwFM/fwfm.py", "ref_fwfm")
    six = ["userid", "feedid", "device", "authorid", "bgm_song_id", "bgm_singer_id"]
    dims = [SMALL_LINES[c] for c in six]
    for D in (8, 5):
        torch.manual_seed(21)
        m = ref.FwFM(dims, D)
        with torch.no_grad():
            m.bias.fill_(0.25)          # the reference initialises it to 0: make it visible in y
        x = {c: rand_idx(gen, SMALL_LINES[c], (B,)) for c in six}
        fx = record(m, lambda: (m(x),), 29, dict(
            model="FwFM", ctor=dict(field_dims=dims, embed_dim=D), inputs=dict(x=x)))
        save(fx, f"fwfm_d{D}.pt")
    # ---- loader: batches of the reference WechatDataset + DataLoader on the synthetic frame of tests/loader_cases.py
    sys.path.insert(0, os.path.dirname(HERE))
    import loader_cases
    from torch.utils.data import DataLoader
    ldir = tempfile.mkdtemp(prefix="rk_golden_loader_")
    lvocab = loader_cases.write_vocab(os.path.join(ldir, "vocab")) + "/"
    lpath = os.path.join(ldir, "frame.parquet")
    loader_cases.make_frame("string").to_parquet(lpath)
    recorded = {}
    for kind, rel in (("deepfm", "DeepFM/deepfm.py"), ("dcn", "DCN/dcn.py"), ("deepcrossing", "DeepCrossing/deepcrossing.py"),
                      ("din", "DIN/din.py"), ("bst", "BST/bst.py")):
        ref = load_reference(rel, "ref_loader_" + kind)
        ds = ref.WechatDataset(lpath, lvocab, 5) if kind == "bst" else ref.WechatDataset(lpath, lvocab)
        kw = {"collate_fn": ref.din_collate_fn} if kind == "din" else {}
        recorded[kind] = list(DataLoader(ds, batch_size=8, shuffle=False, **kw))
    save(recorded, "loader_batches.pt")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
