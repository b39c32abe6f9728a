"""Prints per-tensor errors of a product module against a golden fixture (GPU needed)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import rank_b200
import golden_cases
from conftest import load_golden, to_device, rel_err
from oracle import models as oracle_models

for path in sys.argv[1:]:
    fx = load_golden(path)
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(), fx.get("vocab_lines")) + "/" if "vocab_lines" in fx else None
    m = golden_cases.build(fx, rank_b200, vocab, oracle=False)
    m.load_state_dict(fx["state_dict"]); m.to("cuda")
    outs, grads = golden_cases.replay(m, fx, to_device(fx["inputs"], "cuda"), to_device(fx["cotangents"], "cuda"))
    r64 = golden_cases.build(fx, oracle_models, vocab, oracle=True); r64.load_state_dict(fx["state_dict"]); r64.double()
    dbl = lambda o: o.double() if torch.is_tensor(o) and o.is_floating_point() else ({k: dbl(v) for k, v in o.items()} if isinstance(o, dict) else ([dbl(v) for v in o] if isinstance(o, list) else o))
    _, g64 = golden_cases.replay(r64, fx, dbl(fx["inputs"]), dbl(fx["cotangents"]))
    print("==", os.path.basename(path))
    for i, (o, r) in enumerate(zip(outs, fx["outputs"])):
        if torch.is_tensor(r): print(f"  out{i}: {rel_err(o, r):.2e}")
    gmax = max(float(g.abs().max()) for g in fx["grads"].values())
    for k, g in fx["grads"].items():
        own = float(g.abs().max())
        print(f"  {k:45s} own/gmax {own/gmax:8.1e}  vs32 {rel_err(grads[k], g):.2e}  vs64 {rel_err(grads[k], g64[k]):.2e}  ref32vs64 {rel_err(g, g64[k]):.2e}")
