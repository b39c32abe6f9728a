#!/usr/bin/env python
"""BST with the feedid table row-sharded over the ranks (BASELINE config 5 "scaled").

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_bst.py [--rows 100000000]

1. parity: on a small table, every rank checks the sharded model (local batch, all-to-all) against
   a replicated model fed the GLOBAL batch — logits bit-comparable, shard gradient = the owner's
   slice of the replicated dense gradient, tower gradients equal after the all-reduce average;
2. scale: the table grown to --rows rows (default 1e8 x 16 floats = 6.4 GB over the ranks), sparse
   gradients, a few timed steps; prints one JSON line from rank 0.
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

import rank_b200
from rank_b200 import synthetic
from rank_b200.parallel import GradientAllReducer


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--batch", type=int, default=1024, help="per-rank batch (survey: 1024/GPU when sharded)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--eager", action="store_true", help="launch every step eagerly (default: forward + backward + "
                    "both collectives replayed from one CUDA graph; the exchange has no host sync)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("RANK_B200_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("RANK_B200_TRACE_AFTER", "60")), exit=False)

    def stage(msg):
        if os.environ.get("RANK_B200_TRACE"):
            print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    vocab = rank_b200.write_vocab_dir(tempfile.mkdtemp(prefix="rk_vocab_")) + "/"
    kw = dict(dropout_rate=0.0, nhead=4, num_transformer_blocks=1, max_seq_length=20)
    T, B = 20, args.batch

    # ---- 1. parity on the real-vocabulary table (106445 rows)
    torch.manual_seed(0)
    full = rank_b200.BSTModel(vocab, **kw).to(dev)
    torch.manual_seed(0)
    shard = rank_b200.shard_bst_feedid_table(rank_b200.BSTModel(vocab, **kw).to(dev), sparse_grad=False)
    reducer = GradientAllReducer(shard)
    batches = [synthetic.to_device(synthetic.bst_batch(B, T, seed=100 + r), dev) for r in range(world)]
    mine = batches[rank]

    def loss_of(model, b):
        logit = model(b["dense"], b["category"], b["seq_feedid"], b["seq_length"])[1]
        return F.binary_cross_entropy_with_logits(logit.squeeze(), b["label"]), logit

    shard.train(); full.train()
    stage("models built")
    l_s, logit_s = loss_of(shard, mine)
    torch.cuda.synchronize(); stage("sharded forward done")
    l_s.backward()
    torch.cuda.synchronize(); stage("sharded backward done")
    reducer.allreduce()
    torch.cuda.synchronize(); stage("all-reduce done")
    # the replicated model sees every rank's batch; BatchNorm uses per-rank statistics, so run the
    # ranks' batches one by one and average the losses (= what data parallelism computes)
    total = 0
    for r in range(world):
        l_r, logit_r = loss_of(full, batches[r])
        (l_r / world).backward()
        if r == rank:
            e_logit = rel(logit_s, logit_r)
    torch.cuda.synchronize(); stage("replicated passes done")
    lo, hi = shard.embeddings["feedid"].row_range
    e_shard = rel(shard.embeddings["feedid"].weight.grad[:hi - lo], full.embeddings["feedid"].weight.grad[lo:hi])
    e_rest = max(rel(ps.grad, pf.grad) for (ns, ps), (nf, pf) in zip(shard.named_parameters(), full.named_parameters())
                 if "feedid" not in ns and float(pf.grad.abs().max()) > 1e-6)
    ok = e_logit <= 1e-5 and e_shard <= 1e-5 and e_rest <= 1e-4
    flags = torch.tensor([float(ok), e_logit, e_shard, e_rest], device=dev)
    gathered = [torch.zeros_like(flags) for _ in range(world)]
    dist.all_gather(gathered, flags)
    rank_b200.check_index_errors()
    del full, shard, reducer

    # ---- 2. the scaled table: --rows x 16, sharded, sparse gradients
    torch.manual_seed(0)
    model = rank_b200.BSTModel(vocab, **kw).to(dev)
    big = rank_b200.RowShardedEmbedding(args.rows, 16, sparse_grad=True).to(dev)
    model.embeddings["feedid"] = big
    reducer = GradientAllReducer(model)
    dense_params = [p for p in model.parameters() if not getattr(p, "_rank_local", False)]
    opt_dense = torch.optim.Adam(dense_params, lr=1e-3, fused=True)
    from rank_b200.optim import RowwiseAdam
    opt_sparse = RowwiseAdam([big.weight], lr=1e-3)       # one kernel on the touched rows (SparseAdam's arithmetic)
    data = [synthetic.to_device(synthetic.bst_batch(B, T, seed=500 + 31 * rank + i, feed_rows=args.rows), dev)
            for i in range(4)]
    times = []
    model.train()

    def fwd_bwd(batch):
        loss, _ = loss_of(model, batch)
        loss.backward()
        reducer.allreduce()
        return loss

    graph = None
    if not args.eager:
        # static inputs, 3 eager steps on a side stream, then capture: gather kernels, both all-to-alls, the block,
        # the tower, the backward, the owner-side reduction and the dense all-reduce are one graph
        static = {k: (v.clone() if torch.is_tensor(v) else {kk: vv.clone() for kk, vv in v.items()})
                  for k, v in data[0].items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                opt_dense.zero_grad(set_to_none=True); opt_sparse.zero_grad(set_to_none=True)
                fwd_bwd(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        stage("warm-up done")
        opt_dense.zero_grad(set_to_none=True); opt_sparse.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = fwd_bwd(static)
        captured_touched = big.weight.touched_grad
        torch.cuda.synchronize()
        stage("capture done")

        def copy_in(dst, src):
            for k, v in src.items():
                if torch.is_tensor(v):
                    dst[k].copy_(v, non_blocking=True)
                else:
                    copy_in(dst[k], v)

    for i in range(3 + args.steps):
        dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        stage(f"step {i}")
        if graph is None:
            opt_dense.zero_grad(set_to_none=True); opt_sparse.zero_grad(set_to_none=True)
            loss = fwd_bwd(data[i % 4])
        else:
            copy_in(static, data[i % 4])
            graph.replay()
            big.weight.touched_grad = captured_touched        # refilled by the replay; the optimizer consumed the last one
        opt_dense.step(); opt_sparse.step()
        e.record(); torch.cuda.synchronize()
        if i >= 3:
            times.append(s.elapsed_time(e))
    t = torch.tensor([sum(times) / len(times)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "what": "BST, feedid table row-sharded over the ranks, all-to-all lookup, sparse row-wise update",
            "n_gpus": world, "table_rows": args.rows, "shard_gb": big.weight.numel() * 4 / 1e9,
            "batch_per_gpu": B, "ms_per_step": float(t), "samples_per_s": world * B / (float(t) / 1e3),
            "step": "zero_grad+fwd+loss+bwd+grad allreduce+Adam(dense)+RowwiseAdam(shard)",
            "launch": "eager" if graph is None else "cuda graph (fwd+bwd+all-to-alls+all-reduce) + eager optimizers",
            "parity_vs_replicated": {"ok_all_ranks": all(bool(g[0] > 0.5) for g in gathered),
                                     "logit_rel_err": max(float(g[1]) for g in gathered),
                                     "shard_grad_rel_err": max(float(g[2]) for g in gathered),
                                     "other_grad_rel_err": max(float(g[3]) for g in gathered)},
            "loss": float(loss.detach())}), flush=True)
    # a captured graph holding NCCL work must be gone before the communicator is (the exit otherwise hangs)
    graph = None
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
