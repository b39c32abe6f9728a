#!/usr/bin/env python
"""Device time of the fused tower layers (csrc/tower.cu) per ABI call: CUDA events around graph replays of
one forward / one backward of a single layer, L2 flushed between replays.
    python scripts/bench_tower.py [--batch 8192]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import rank_b200
from rank_b200.tower import run_tower


WARM = False


def timed(fn, flush, n=30):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    evs = []
    for i in range(n):
        if not WARM:
            flush.fill_(i & 0xff)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return 1e3 * sum(s.elapsed_time(e) for s, e in evs) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--warm", action="store_true", help="no L2 flush between replays (x as the preceding GEMM leaves it)")
    args = ap.parse_args()
    global WARM
    WARM = args.warm
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for kind, units in (("dice_bn", 200), ("dice_bn", 80), ("dice", 200), ("bn_lrelu", 512), ("bn_lrelu", 256), ("bn_lrelu", 128)):
        torch.manual_seed(0)
        if kind == "dice_bn":
            layers = [rank_b200.Dice(units), nn.BatchNorm1d(units)]
        elif kind == "dice":
            layers = [rank_b200.Dice(units)]
        else:
            layers = [nn.BatchNorm1d(units), nn.LeakyReLU(0.01)]
        layers = [l.to(dev).train() for l in layers]
        x = torch.randn(args.batch, units, device=dev, requires_grad=True)
        g = torch.randn(args.batch, units, device=dev)
        res = {}
        for fused in (True, False):
            rank_b200.tower.FUSED = fused
            with torch.no_grad():
                fwd_us = timed(lambda: run_tower(layers, x), flush)

            def both():
                x.grad = None
                run_tower(layers, x).backward(g)
            both_us = timed(both, flush)
            res["fused" if fused else "torch"] = {"fwd_us": round(fwd_us, 1), "fwd_bwd_us": round(both_us, 1)}
        rank_b200.tower.FUSED = True
        out[f"{kind}_{units}"] = res
        print(kind, units, res, flush=True)
    print(json.dumps({"batch": args.batch, "layers": out}))


if __name__ == "__main__":
    main()
