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
