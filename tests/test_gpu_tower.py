"""Fused Dice(+BatchNorm1d) tower kernel against the torch modules it replaces (same module objects,
fusion switched off): outputs, every gradient, running statistics and num_batches_tracked."""
import copy

import pytest
import torch
import torch.nn as nn

import rank_b200
from rank_b200 import tower
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


@pytest.fixture(autouse=True)
def _fuse_every_batch_size():
    """The kernels are exercised at every size here; in the models small batches run the torch modules."""
    saved = tower.MIN_BATCH
    tower.MIN_BATCH = 2
    yield
    tower.MIN_BATCH = saved


def _layers(units, with_bn2):
    torch.manual_seed(1)
    dice = rank_b200.Dice(units)
    with torch.no_grad():
        dice.alpha.uniform_(-0.5, 0.5)
    mods = [dice]
    if with_bn2:
        bn = nn.BatchNorm1d(units)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
        mods.append(bn)
    mods.append(nn.Linear(units, 3))
    return nn.ModuleList(mods).to(DEV).train()


@pytest.mark.parametrize("B,units,with_bn2", [(8192, 512, True), (8192, 256, True), (8192, 128, True), (1000, 200, True),
                                              (333, 80, True), (2, 8, True), (4096, 37, False), (16384, 64, True)])
def test_fused_dice_bn_matches_modules(B, units, with_bn2):
    fused, plain = _layers(units, with_bn2), None
    plain = copy.deepcopy(fused)
    gen = torch.Generator().manual_seed(B + units)
    x0 = (torch.randn(B, units, generator=gen) * 1.7 + 0.3).to(DEV)
    cot = torch.randn(B, 3, generator=gen).to(DEV) / B
    results = []
    for mods, flag in ((fused, True), (plain, False)):
        tower.FUSED = flag
        try:
            for step in range(2):                         # two steps: running stats accumulate
                x = x0.clone().requires_grad_()
                out = tower.run_tower(mods, x)
                mods.zero_grad()
                (out * cot).sum().backward()
        finally:
            tower.FUSED = True
        results.append((out.detach(), x.grad.detach(), {k: p.grad.detach() for k, p in mods.named_parameters()},
                        {k: b.detach().clone() for k, b in mods.named_buffers()}))
    (o1, gx1, gp1, buf1), (o2, gx2, gp2, buf2) = results
    TOL = 1e-5 if B >= 8 else 1e-4        # batch norm over 2 samples: the gradient is a difference of near-equal terms
    assert rel_err(o1, o2) <= TOL
    assert rel_err(gx1, gx2) <= TOL
    for k in gp2:
        assert rel_err(gp1[k], gp2[k]) <= TOL, k
    for k in buf2:
        if buf2[k].dtype == torch.int64:
            assert torch.equal(buf1[k], buf2[k]), k       # num_batches_tracked
        else:
            assert rel_err(buf1[k], buf2[k]) <= TOL, k


def test_eval_mode_and_switch_run_the_modules():
    mods = _layers(64, True)
    x = torch.randn(128, 64, device=DEV)
    mods.eval()
    a = tower.run_tower(mods, x)
    b = x
    for m in mods:
        b = m(b)
    assert torch.equal(a, b)


@pytest.mark.parametrize("B,units,act", [(1024, 512, nn.ReLU()), (8192, 256, nn.LeakyReLU(0.01)), (8192, 1024, nn.ReLU()),
                                         (777, 128, nn.LeakyReLU(0.2)), (16384, 96, nn.ReLU())])
def test_fused_bn_activation_matches_modules(B, units, act):
    torch.manual_seed(2)
    bn = nn.BatchNorm1d(units)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.3, 0.3)
    fused = nn.ModuleList([bn, act, nn.Linear(units, 3)]).to(DEV).train()
    plain = copy.deepcopy(fused)
    gen = torch.Generator().manual_seed(B + units)
    x0 = (torch.randn(B, units, generator=gen) * 1.3 - 0.2).to(DEV)
    cot = torch.randn(B, 3, generator=gen).to(DEV) / B
    results = []
    for mods, flag in ((fused, True), (plain, False)):
        tower.FUSED = flag
        try:
            for step in range(2):
                x = x0.clone().requires_grad_()
                out = tower.run_tower(mods, x)
                mods.zero_grad()
                (out * cot).sum().backward()
        finally:
            tower.FUSED = True
        results.append((out.detach(), x.grad.detach(), {k: p.grad.detach() for k, p in mods.named_parameters()},
                        {k: b.detach().clone() for k, b in mods.named_buffers()}))
    (o1, gx1, gp1, buf1), (o2, gx2, gp2, buf2) = results
    assert rel_err(o1, o2) <= TOL and rel_err(gx1, gx2) <= TOL
    for k in gp2:
        assert rel_err(gp1[k], gp2[k]) <= TOL, k
    for k in buf2:
        assert (torch.equal(buf1[k], buf2[k]) if buf2[k].dtype == torch.int64 else rel_err(buf1[k], buf2[k]) <= TOL), k
