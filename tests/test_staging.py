"""PackedBatch (SURVEY 8(f) input staging): layout, bit-exact round trip, one copy per step."""
import pytest
import torch

import rank_b200
from rank_b200 import synthetic
from rank_b200.staging import PackedBatch, _leaves


def _same(a, b):
    la, lb = dict(_leaves(a)), dict(_leaves(b))
    assert la.keys() == lb.keys()
    for k in la:
        assert la[k].dtype == lb[k].dtype and la[k].shape == lb[k].shape, k
        assert torch.equal(la[k].cpu(), lb[k].cpu()), k


@pytest.mark.parametrize("make", [lambda: synthetic.din_batch(64, 50, 3), lambda: synthetic.side_batch(33, 4),
                                  lambda: synthetic.fwfm_batch(7, 5), lambda: synthetic.bst_batch(16, 20, 6)])
def test_layout_and_host_round_trip(make):
    batch = make()
    packed = PackedBatch.like(batch, "cpu", pin=False)
    offs = [off for _, _, _, off in packed.layout]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs)
    assert packed.payload_bytes <= packed.nbytes
    packed.fill(batch)
    _same(packed.host_views, batch)                 # bytes are copied, never converted
    views = packed.to_device(non_blocking=False)    # device "cpu" here: same code path, one copy
    _same(views, batch)
    assert views is packed.device_views             # stable addresses: usable as CUDA-graph inputs


def test_fill_rejects_a_different_shape():
    packed = PackedBatch.like(synthetic.side_batch(8, 1), "cpu", pin=False)
    with pytest.raises(ValueError):
        packed.fill(synthetic.side_batch(9, 1))


@pytest.mark.gpu
def test_one_copy_feeds_the_model_bit_exact(wechat_vocab_dir):
    dev = torch.device("cuda", 0)
    batch = synthetic.din_batch(512, 50, 9)
    packed = PackedBatch.like(batch, dev)
    views = packed.fill(batch).to_device()
    torch.cuda.synchronize()
    _same(views, batch)
    torch.manual_seed(0)
    model = rank_b200.DIN(wechat_vocab_dir, dropout_rate=0.0).to(dev).train()
    outs = []
    for inp in (views, synthetic.to_device(batch, dev)):          # packed views vs per-tensor .to(device)
        torch.manual_seed(5)
        prob, _, l2 = model(inp["dense"], inp["category"], inp["sequence"], inp["target"])
        outs.append((prob.detach().clone(), l2.detach().clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
def test_refill_waits_for_the_previous_copy():
    """fill -> to_device -> fill with the copy still queued behind a busy stream: the second fill must
    not overwrite the pinned buffer before the first copy has read it."""
    from rank_b200 import _lib
    dev = torch.device("cuda", 0)
    a, b = synthetic.side_batch(4096, 1), synthetic.side_batch(4096, 2)
    packed = PackedBatch.like(a, dev)
    lib = _lib.load()
    packed.fill(a)
    _lib.check(lib.rk_debug_spin(50_000, _lib.stream_ptr()), "rk_debug_spin")   # 50 ms ahead of the copy
    views = packed.to_device()
    snap = {k: v.clone() for k, v in dict(_leaves(views)).items()}              # queued behind the copy
    packed.fill(b)                                                              # must wait for the copy
    torch.cuda.synchronize()
    want = dict(_leaves(a))
    for k, v in snap.items():
        assert torch.equal(v.cpu(), want[k]), k


@pytest.mark.gpu
def test_prefetcher_delivers_every_batch_in_order():
    """Double-buffered prefetch: copy k+1 is in flight on the copy stream while step k (a busy main stream)
    runs; every step must see exactly its own batch in the static device views."""
    from rank_b200 import _lib
    from rank_b200.staging import Prefetcher
    dev = torch.device("cuda", 0)
    batches = [synthetic.side_batch(2048, 10 + i) for i in range(5)]
    hosts = [PackedBatch.like(batches[0], dev).fill(b) for b in batches]
    static = PackedBatch.like(batches[0], dev)
    pf = Prefetcher(static)
    lib = _lib.load()
    seen = []
    slot = pf.submit(hosts[0])
    for k in range(5):
        nxt = pf.submit(hosts[k + 1]) if k + 1 < 5 else None
        views = pf.consume(slot)
        _lib.check(lib.rk_debug_spin(2000, _lib.stream_ptr()), "rk_debug_spin")      # the "step": 2 ms of work
        seen.append({key: v.clone() for key, v in dict(_leaves(views)).items()})
        slot = nxt
    torch.cuda.synchronize()
    for k, snap in enumerate(seen):
        want = dict(_leaves(batches[k]))
        for key, v in snap.items():
            assert torch.equal(v.cpu(), want[key]), (k, key)
