#!/usr/bin/env python
"""Which cuBLAS path does torch take for the tower's weight-gradient GEMMs (dW = g^T x, K = batch)?  Times the
autograd form (g.t() @ x), the transposed-copy form (g.t().contiguous() @ x) and x^T g variants at the DIN /
BST / DeepFM tower shapes.   python scripts/bench_mm.py"""
import torch


def timed(fn, n=50):
    for _ in range(5):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / n


def main():
    dev = torch.device("cuda")
    B = 8192
    for k_in, n_out in ((82, 200), (200, 80), (66, 512), (512, 256), (256, 128)):
        x = torch.randn(B, k_in, device=dev)
        g = torch.randn(B, n_out, device=dev)
        w = torch.randn(n_out, k_in, device=dev)
        res = {
            "fwd x@w.t": timed(lambda: torch.addmm(torch.zeros(n_out, device=dev), x, w.t())),
            "dx g@w": timed(lambda: g @ w),
            "dW g.t()@x (autograd)": timed(lambda: g.t() @ x),
            "dW g.t().contiguous()@x": timed(lambda: g.t().contiguous() @ x),
            "dW (x.t()@g).t()": timed(lambda: (x.t() @ g).t()),
            "dW x.t().contiguous()@g": timed(lambda: x.t().contiguous() @ g),
            "db g.sum(0)": timed(lambda: g.sum(0)),
            "db ones@g": timed(lambda: torch.ones(1, B, device=dev) @ g),
        }
        print(f"B={B} in={k_in} out={n_out}: " + ", ".join(f"{k} {v:.1f}us" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
