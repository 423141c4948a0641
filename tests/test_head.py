"""Classification head (model_discCNN.py:547-566) and the loss terms (train_and_validate.py:136-140) on the
hand-written kernels, against an fp64 torch evaluation of the same modules."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import relerr

import genome_nas_b200.head as HD
from genome_nas_b200.flat import region_of


class _Head(nn.Module):
    _head_wgrad = True

    def __init__(self, T, H, D1, NC, sigmoid=True, p=0.1):
        super().__init__()
        self.decoder, self.classifier = nn.Linear(T * H, D1), nn.Linear(D1, NC)
        self.final = nn.Sigmoid() if sigmoid else nn.Identity()
        self.dropout_lin = nn.Dropout(p)


# (T, B, H, D1, classes, sigmoid, p): ragged tiles, batch above one tile, H % 4 != 0, the reference's 925 / 919 widths
CASES = [(5, 3, 6, 7, 5, True, 0.0), (7, 66, 12, 70, 67, True, 0.3), (3, 4, 8, 925, 919, True, 0.1),
         (42, 5, 16, 33, 9, False, 0.1), (2, 1, 4, 1, 1, True, 0.0)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(str(int(v)) if not isinstance(v, float) else str(v) for v in c))
@pytest.mark.parametrize("wgrad", [True, False])
def test_head_and_losses_match_fp64(backend, case, wgrad):
    T, B, H, D1, NC, sig, p = case
    torch.manual_seed(0)
    m = _Head(T, H, D1, NC, sig, p).train()
    ref = _Head(T, H, D1, NC, sig, p).double()
    ref.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    m = m.to(backend)
    m._head_wgrad = wgrad
    hs = torch.randn(T, B, H)
    y = (torch.rand(B, NC) > 0.5).float()
    region_of(m)
    h1 = hs.clone().to(backend).requires_grad_(True)
    masks = []
    orig = F.dropout

    def recording_dropout(x, pp, training):
        out = orig(x, pp, training)
        masks.append(out.detach().double().cpu())
        return out
    HD.F.dropout = recording_dropout
    try:
        out = HD.head_forward(m, h1)
    finally:
        HD.F.dropout = orig
    loss = (HD.bce_loss(out, y.to(backend)) if sig else out.pow(2).mean()) + HD.tar_loss(h1, 0.1)
    loss.backward()
    h2 = hs.double().requires_grad_(True)
    x = torch.flatten(h2.permute(1, 0, 2), 1)
    if p > 0:
        x = x * masks[0]
    o = ref.final(ref.classifier(torch.relu(ref.decoder(x))))
    l = (F.binary_cross_entropy(o, y.double()) if sig else o.pow(2).mean()) + 0.1 * (h2[1:] - h2[:-1]).pow(2).mean()
    l.backward()
    assert relerr(out, o) < 2e-6
    assert abs(float(loss.detach()) - float(l)) < 2e-6 * abs(float(l))
    assert relerr(h1.grad, h2.grad) < 5e-6
    for name in ("decoder", "classifier"):
        for pn in ("weight", "bias"):
            g = getattr(getattr(m, name), pn).grad
            if wgrad:
                assert relerr(g, getattr(getattr(ref, name), pn).grad) < 5e-6, (name, pn)
            else:                                   # the first-order alpha pass discards them: kernels not launched
                assert g is None or float(g.abs().max()) == 0.0


def test_head_eval_mode_and_accumulation(backend):
    T, B, H, D1, NC = 4, 3, 8, 10, 6
    torch.manual_seed(1)
    m = _Head(T, H, D1, NC, True, 0.5).to(backend).eval()
    hs = torch.randn(T, B, H, device=backend, requires_grad=True)
    out = HD.head_forward(m, hs)                      # eval: no dropout
    x = torch.flatten(hs.detach().permute(1, 0, 2), 1)
    ref = torch.sigmoid(F.linear(torch.relu(F.linear(x, m.decoder.weight, m.decoder.bias)), m.classifier.weight, m.classifier.bias))
    assert relerr(out, ref) < 2e-6
    out.sum().backward()
    g1 = m.decoder.weight.grad.clone()
    out = HD.head_forward(m, hs)
    out.sum().backward()                              # weight gradients accumulate like autograd's
    assert relerr(m.decoder.weight.grad, 2 * g1) < 1e-6


def test_bce_clamps_like_torch(backend):
    p = torch.tensor([[0.0, 1.0, 0.5, 1e-30]], device=backend, requires_grad=True)
    y = torch.tensor([[1.0, 0.0, 1.0, 0.0]], device=backend)
    loss = HD.bce_loss(p, y)
    ref = F.binary_cross_entropy(p.detach().cpu().double(), y.cpu().double())
    assert abs(float(loss.detach()) - float(ref)) < 1e-5 * float(ref)
    loss.backward()
    assert torch.isfinite(p.grad).all()
