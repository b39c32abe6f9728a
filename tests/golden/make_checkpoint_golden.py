"""Generates tests/golden/checkpoint_{dcn,deepcrossing}.pt from the two trained state_dicts the
reference ships (DCN/model_dir/best_model.pth, DeepCrossing/model_dir/best_model.pth), run through the
UNMODIFIED reference classes on CPU with the real WeChat vocabulary sizes.

    python tests/golden/make_checkpoint_golden.py        (authoring container, needs /root/reference)

/root/reference does not exist on the GPU box, so the CUDA modules can only meet the real weights
through a committed fixture.  To keep it small the embedding tables are stored sparsely: only the
rows the recorded batch touches (a forward reads nothing else, and every other row's gradient is
exactly zero, which the script asserts); `expand()` rebuilds full-height tables with zeros elsewhere.
The DNN / output weights are stored whole.  Protocol as make_golden.py: model.train(),
torch.manual_seed(seed) right before the forward, loss = sum_k <out_k, cot_k>.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_reference  # noqa: E402

VOCAB = os.path.join(REF, "dataset/wechat_algo_data1/vocabulary/")
B = 256


def sparsify(sd, touched):
    """Tables -> {"rows": idx, "values": rows, "height": V, "dim": D}; other tensors unchanged."""
    out = {}
    for k, v in sd.items():
        col = k.split(".")[1] if k.startswith("embeddings.") else None
        if col in touched:
            rows = touched[col]
            out[k] = {"rows": rows, "values": v[rows].clone(), "height": v.shape[0], "dim": v.shape[1]}
        else:
            out[k] = v.clone()
    return out


def main():
    gen = torch.Generator().manual_seed(20261018)
    for which, rel, ctor in (("dcn", "DCN/dcn.py", dict(num_cross_layer=3)),
                             ("deepcrossing", "DeepCrossing/deepcrossing.py",
                              dict(residual_internal_dim=128, residual_network_num=2))):
        ref = load_reference(rel, "ref_ckpt_" + which)
        folder = "DCN" if which == "dcn" else "DeepCrossing"
        sd = torch.load(os.path.join(REF, "algorithm", folder, "model_dir/best_model.pth"), map_location="cpu")
        cls = ref.DCNModel if which == "dcn" else ref.DeepCrossingModel
        m = cls(VOCAB, **ctor)
        m.load_state_dict(sd, strict=True)
        m.train()
        dense = torch.log1p(torch.poisson(torch.full((B, 16), 3.0), generator=gen))
        cat = {}
        for c, e in m.embeddings.items():
            idx = torch.randint(0, e.num_embeddings, (B,), generator=gen)
            idx[: B // 4] = idx[0]                 # a hot row: duplicate indices in every table
            idx[-1] = e.num_embeddings - 1
            cat[c] = idx
        seed = 31
        torch.manual_seed(seed)
        outs = m(dense, cat)
        cgen = torch.Generator().manual_seed(seed + 1)
        cots = [torch.randn(o.shape, generator=cgen) for o in outs]
        m.zero_grad()
        sum((o * c).sum() for o, c in zip(outs, cots)).backward()
        touched = {c: torch.unique(i) for c, i in cat.items()}
        grads = {}
        for k, p in m.named_parameters():
            col = k.split(".")[1] if k.startswith("embeddings.") else None
            if col in touched:
                rest = p.grad.clone()
                rest[touched[col]] = 0
                assert float(rest.abs().max()) == 0.0, k      # untouched rows: exactly zero gradient
                grads[k] = {"rows": touched[col], "values": p.grad[touched[col]].clone(),
                            "height": p.shape[0], "dim": p.shape[1]}
            else:
                grads[k] = p.grad.detach().clone()
        vocab_lines = {c: e.num_embeddings - 1 for c, e in m.embeddings.items()}
        fx = dict(model="DCNModel" if which == "dcn" else "DeepCrossingModel", ctor=ctor, seed=seed,
                  vocab_lines=vocab_lines, inputs=dict(dense=dense, category=cat),
                  state_dict=sparsify(m.state_dict(), touched), outputs=[o.detach().clone() for o in outs],
                  cotangents=cots, grads=grads,
                  source=f"algorithm/{folder}/model_dir/best_model.pth through the reference {cls.__name__}")
        torch.save(fx, os.path.join(HERE, f"checkpoint_{which}.pt"))
        print(which, {k: (tuple(v["values"].shape) if isinstance(v, dict) else tuple(v.shape))
                      for k, v in fx["state_dict"].items()})


if __name__ == "__main__":
    main()
