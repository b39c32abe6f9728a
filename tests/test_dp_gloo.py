"""World-size-2 data-parallel plumbing on CPU (gloo): the averaged gradient of two half batches
equals the full-batch gradient of a mean-reduced loss."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rank_b200.parallel import GradientAllReducer, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Embedding(11, 4), torch.nn.Flatten(0), )
    emb = torch.nn.Embedding(11, 4)
    lin = torch.nn.Linear(4, 1)
    unused = torch.nn.Linear(3, 3)
    net = torch.nn.ModuleDict(dict(emb=emb, lin=lin, unused=unused))
    gen = torch.Generator().manual_seed(1)
    full = dict(idx=torch.randint(0, 11, (16,), generator=gen), y=torch.randn(16, generator=gen))
    mine = shard_batch(full, rank, world)
    reducer = GradientAllReducer(net)
    loss = ((lin(emb(mine["idx"])).squeeze(1) - mine["y"]) ** 2).mean()
    loss.backward()
    reducer.allreduce()
    flat = reducer.flat_gradients()
    if rank == 0:
        net.zero_grad()
        ref = ((lin(emb(full["idx"])).squeeze(1) - full["y"]) ** 2).mean()
        ref.backward()
        want = torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel())
                          for p in net.parameters()])
        out.put(float((flat - want).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() < 1e-6


def test_shard_batch_nested():
    b = dict(a=torch.arange(8), d=dict(x=torch.arange(16).view(8, 2)))
    s = shard_batch(b, 1, 2)
    assert s["a"].tolist() == [4, 5, 6, 7] and s["d"]["x"].shape == (4, 2)
