"""RowwiseAdam (csrc/optim.cu) against torch.optim.SparseAdam on the same sparse gradients."""
import pytest
import torch
import torch.nn as nn

import rank_b200
from rank_b200.optim import RowwiseAdam
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("V,D", [(5000, 16), (333, 6), (70000, 5), (100, 32)])
def test_matches_sparse_adam_over_several_steps(V, D):
    torch.manual_seed(0)
    a = nn.Embedding(V, D, sparse=True).to(DEV)
    b = nn.Embedding(V, D, sparse=True).to(DEV)
    b.load_state_dict(a.state_dict())
    oa = RowwiseAdam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    ob = torch.optim.SparseAdam(b.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    gen = torch.Generator().manual_seed(1)
    for step in range(5):
        idx = torch.randint(0, V, (4096,), generator=gen).to(DEV)       # duplicates: the gradient is coalesced
        cot = torch.randn(4096, D, generator=gen).to(DEV)
        for emb, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            (emb(idx) * cot).sum().backward()
            opt.step()
        assert rel_err(a.weight, b.weight) <= 2e-6, step
    sa, sb = oa.state[a.weight], ob.state[b.weight]
    assert sa["step"] == 5
    assert rel_err(sa["exp_avg"], sb["exp_avg"]) <= 2e-6 and rel_err(sa["exp_avg_sq"], sb["exp_avg_sq"]) <= 2e-6
    rank_b200.check_index_errors()


def test_rejects_dense_gradients():
    emb = nn.Embedding(10, 4).to(DEV)
    opt = RowwiseAdam(emb.parameters())
    emb(torch.tensor([1, 2], device=DEV)).sum().backward()
    with pytest.raises(RuntimeError):
        opt.step()


# ---- touched-rows gradients of the REPLICATED tables --------------------------
def _model_and_batch(name, vocab_dir):
    from rank_b200 import synthetic
    torch.manual_seed(0)
    if name == "DeepFM":
        return rank_b200.DeepFM(vocab_dir, embedding_dim=16, dropout_rate=0.0), synthetic.deepfm_batch(2048), name
    if name == "DCNModel":
        return rank_b200.DCNModel(vocab_dir, num_cross_layer=3), synthetic.side_batch(8192), name
    if name == "DIN":
        return rank_b200.DIN(vocab_dir, dropout_rate=0.0), synthetic.din_batch(1024, 50), name
    if name == "BSTModel":
        return (rank_b200.BSTModel(vocab_dir, dropout_rate=0.0, max_seq_length=20), synthetic.bst_batch(1024, 20), name)
    raise KeyError(name)


@pytest.fixture
def touched_mode():
    from rank_b200 import sparse
    yield sparse
    sparse.set_table_gradients("dense")


@pytest.mark.parametrize("name", ["DeepFM", "DCNModel", "DIN", "BSTModel"])
def test_touched_rows_equal_the_dense_gradient(wechat_vocab_dir, touched_mode, name):
    """set_table_gradients("touched"): every table gets (distinct rows, summed gradient rows, device count)
    and no dense .grad; expanded, it is the dense gradient the default mode produces."""
    import golden_cases
    from conftest import to_device
    model, batch, fx_model = _model_and_batch(name, wechat_vocab_dir)
    model = model.to(DEV)
    fx = {"model": fx_model, "seed": 3}
    inputs = to_device({k: v for k, v in batch.items() if k != "label"}, DEV)
    with torch.no_grad():
        torch.manual_seed(3)
        n_out = sum(torch.is_tensor(o) for o in golden_cases.call(model, fx, inputs))
    B = batch["label"].shape[0]
    gen = torch.Generator().manual_seed(3)
    cots = [torch.randn(B, 1, generator=gen) / B for _ in range(n_out)]
    if fx_model == "DIN":
        cots[-1] = torch.tensor(1.0)
    cots = to_device(cots, DEV)
    _, dense = golden_cases.replay(model, fx, inputs, cots)
    dense = {k: v.clone() for k, v in dense.items()}
    touched_mode.set_table_gradients("touched")
    _, sparse_run = golden_cases.replay(model, fx, inputs, cots)
    tables = [k for k in dense if dict(model.named_parameters())[k].dim() == 2 and "embedding" in k
              and "position" not in k]
    assert tables
    params = dict(model.named_parameters())
    for k in dense:
        if k in tables:
            tr = params[k].touched_grad
            assert params[k].grad is None and tr is not None, k
            n = int(tr.count)
            rows = tr.rows[:n]
            assert n <= tr.rows.numel() and torch.equal(rows, torch.unique(rows))        # distinct, ascending
            assert rel_err(tr.to_dense(), dense[k]) <= 1e-6, k
            assert float(tr.values[n:].abs().sum()) == 0.0
            touched_rows = (dense[k].abs().sum(dim=1) != 0).nonzero().flatten()
            assert set(touched_rows.tolist()) <= set(rows.tolist())
            params[k].touched_grad = None
        else:
            assert rel_err(sparse_run[k], dense[k]) <= 1e-6, k
    rank_b200.check_index_errors()


def test_rowwise_adam_on_touched_rows_matches_sparse_adam_and_dense_adam_on_those_rows(wechat_vocab_dir, touched_mode):
    """Three training steps of DCN's tables: RowwiseAdam fed by touched_grad (no host sync) against
    torch.optim.SparseAdam fed by the same gradient as a COO tensor; and after ONE step against dense
    optim.Adam (the reference's optimizer, DCN/dcn.py:246) — identical on the touched rows, and the
    untouched rows do not move in either."""
    from rank_b200 import synthetic
    from conftest import to_device
    torch.manual_seed(0)
    ours = rank_b200.DCNModel(wechat_vocab_dir, num_cross_layer=2).to(DEV)
    twin = rank_b200.DCNModel(wechat_vocab_dir, num_cross_layer=2).to(DEV)
    dense_twin = rank_b200.DCNModel(wechat_vocab_dir, num_cross_layer=2).to(DEV)
    twin.load_state_dict(ours.state_dict())
    dense_twin.load_state_dict(ours.state_dict())
    names = [k for k, p in ours.named_parameters() if k.startswith("embeddings.")]
    po, pt, pd = dict(ours.named_parameters()), dict(twin.named_parameters()), dict(dense_twin.named_parameters())
    # eps far below the gradients: SparseAdam (whose arithmetic RowwiseAdam has) adds eps to sqrt(v), dense Adam
    # to sqrt(v / bias_correction2) — the same update only where |g| >> eps, a property of the two torch
    # optimizers, not of this kernel
    opt_o = RowwiseAdam([po[k] for k in names], lr=1e-2, eps=1e-14)
    opt_t = torch.optim.SparseAdam([pt[k] for k in names], lr=1e-2, eps=1e-14)
    opt_d = torch.optim.Adam([pd[k] for k in names], lr=1e-2, eps=1e-14)
    before = {k: po[k].detach().clone() for k in names}

    def loss_of(model, batch, seed):
        torch.manual_seed(seed)
        prob, _ = model(batch["dense"], batch["category"])
        return torch.nn.functional.binary_cross_entropy(prob.squeeze(1), batch["label"])

    for step in range(3):
        batch = to_device(synthetic.side_batch(4096, seed=100 + step), DEV)
        touched_mode.set_table_gradients("touched")
        ours.zero_grad()
        loss_of(ours, batch, step).backward()
        coo = {k: po[k].touched_grad.to_sparse_coo() for k in names}
        touched_rows = {k: po[k].touched_grad.rows[:int(po[k].touched_grad.count)].clone() for k in names}
        opt_o.step()
        assert all(po[k].touched_grad is None for k in names)
        for k in names:
            pt[k].grad = coo[k]
        opt_t.step()
        for k in names:
            assert rel_err(po[k], pt[k]) <= 2e-6, (step, k)
        if step == 0:
            touched_mode.set_table_gradients("dense")
            dense_twin.zero_grad()
            loss_of(dense_twin, batch, step).backward()
            opt_d.step()
            for k in names:
                rows = touched_rows[k]
                assert rel_err(po[k][rows], pd[k][rows]) <= 1e-4, k      # two torch formulations of one update
                mask = torch.ones(po[k].shape[0], dtype=torch.bool, device=DEV)
                mask[rows] = False
                assert torch.equal(po[k][mask], before[k][mask]) and torch.equal(pd[k][mask], before[k][mask])
    rank_b200.check_index_errors()
