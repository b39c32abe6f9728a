"""Micro-benchmark of the embedding-gradient reductions on one model's fields at batch B:
the one-launch output-partitioned kernel (rk_embgrad_direct_reduce) against rk_plan_build + memset +
rk_embgrad_segment_reduce.  CUDA events, L2 flushed between iterations.

    python scripts/bench_reduce.py [--model dcn|din|afm|deepfm] [--iters 30] [--once]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rank_b200  # noqa: E402
from rank_b200 import synthetic  # noqa: E402
from rank_b200.sparse import GradSource, OccurrencePlan  # noqa: E402

FIELDS = {
    "dcn": [("userid", 16), ("device", 2), ("authorid", 4), ("bgm_song_id", 4), ("bgm_singer_id", 4), ("manual_tag_list", 4)],
    "din": [("userid", 16), ("device", 2), ("authorid", 4), ("bgm_song_id", 4), ("bgm_singer_id", 4), ("manual_tag_list", 4),
            ("feedid", 16)],
    "deepfm": [(c, 16) for c in synthetic.DEEPFM_COLUMNS],
    "afm": [(c, 32) for c in list(synthetic.DEEPFM_COLUMNS) + ["manual_tag_list"]] + [("extra", 32)] * 3,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dcn")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--once", action="store_true", help="one call of each (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(1)
    rows_of = synthetic.table_rows()
    rows_of["extra"] = 100001
    fields = FIELDS[args.model]
    B = args.batch
    idx = [synthetic.zipf_indices(gen, rows_of[c], (B,)).to(dev) for c, _ in fields]
    rows = [rows_of[c] for c, _ in fields]
    width = sum(d for _, d in fields)
    g = torch.randn(B, width, generator=gen).to(dev)
    src, off = [], 0
    for f, (c, d) in enumerate(fields):
        src.append(GradSource(g, off, width, d, rows[f], f))
        off += d
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for name, direct in (("direct", True), ("sorted", False)):
        n = 1 if args.once else args.iters
        times = []
        for i in range(n + (0 if args.once else 3)):
            flush.fill_(i & 0xff)
            g.add_(0.0)                                   # the gradient rows were just written: L2-resident
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rank_b200.sparse.PLAN_ON_SIDE_STREAM = False
            plan = OccurrencePlan(idx, rows, direct=direct)
            grads = plan.reduce_to_dense(src)
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e) * 1e3)
        out[name] = (sum(times[-n:]) / n, grads)
        print(f"{args.model} B={B} {name}: {out[name][0]:.1f} us per step (host launch overhead included)")
    for a, b in zip(out["direct"][1], out["sorted"][1]):
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < 1e-5, err
    print("direct == sorted within 1e-5")


if __name__ == "__main__":
    main()
