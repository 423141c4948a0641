"""DARTSCellSearch (RHN layer) parity against fp64 runs of the unmodified reference (tests/golden/rhn.pt)."""
import os

import pytest
import torch

import genomenas_oracle as O
from conftest import GOLDEN, leaf, relerr

import genome_nas_b200.rnn as R
from genome_nas_b200.rnn import DARTSCellSearch

GOLD = torch.load(os.path.join(GOLDEN, "rhn.pt"))
# the reference's own fp32 RHN backward sits at 7e-5..3.5e-4 of fp64
TOL_OUT, TOL_GRAD = 5e-6, 5e-5


def _build(g, device, monkeypatch):
    m = DARTSCellSearch(g["nhid"], g["nhid"], g["dropouth"], g["dropoutx"], g["switch"])
    P = O.synthetic_params(g["shapes"], 23)
    m.weights = torch.nn.Parameter(P["weights"].clone())
    m.load_state_dict({k[len("rnns.0."):]: v for k, v in P.items() if k.startswith("rnns.0.")})
    m.to(device).train()
    masks = [g["xm"], g["hm"]]   # the masks the reference drew (CPU generator, seed 31)
    monkeypatch.setattr(R, "mask2d", lambda B, D, keep, dev: masks.pop(0).to(dev))
    return m


@pytest.mark.parametrize("tag", list(GOLD))
def test_rhn_matches_reference(backend, tag, monkeypatch):
    g = GOLD[tag]
    m = _build(g, backend, monkeypatch)
    x, h0 = leaf(g["x"], backend), leaf(g["h0"], backend)
    y, last = m(x, h0[None])
    assert tuple(last.shape) == (1, g["B"], g["nhid"]) and torch.equal(last[0], y[-1])
    (y * g["gy"].to(backend)).sum().backward()
    assert relerr(y, g["y64"]) < TOL_OUT
    assert relerr(x.grad, g["dx64"]) < TOL_GRAD
    assert relerr(h0.grad, g["dh064"]) < TOL_GRAD
    assert relerr(m.weights.grad, g["dalpha64"]) < TOL_GRAD
    assert relerr(m._W0.grad, g["dW064"]) < TOL_GRAD
    for w, ref in zip(m._Ws, g["dWs64"]):
        assert relerr(w.grad, ref) < TOL_GRAD
    assert relerr(m.bn.running_mean, g["rm64"]) < 1e-5
    assert relerr(m.bn.running_var, g["rv64"]) < 1e-5
    assert int(m.bn.num_batches_tracked) == 9 * g["T"]


def test_rhn_mask_draw_order_matches_reference(backend):
    """x mask first, then h mask, on the CPU generator (utils.py:354-363)."""
    g = GOLD["full"]
    torch.manual_seed(31)
    xm = R.mask2d(g["B"], g["nhid"], 1.0 - g["dropoutx"], backend)
    hm = R.mask2d(g["B"], g["nhid"], 1.0 - g["dropouth"], backend)
    assert torch.equal(xm.cpu(), g["xm"]) and torch.equal(hm.cpu(), g["hm"])


def test_rhn_eval_mode(backend, monkeypatch):
    g = GOLD["nodrop"]
    m = _build(g, backend, monkeypatch)
    with torch.no_grad():
        m.bn.running_mean.add_(0.03)
        m.bn.running_var.mul_(1.3)
    m.eval()
    with torch.no_grad():
        y, _ = m(g["x"].to(backend), g["h0"].to(backend)[None])
    P = {("rnns.0." + k): v.detach().double().cpu() for k, v in m.state_dict().items() if v.is_floating_point()}
    yo, _ = O.rhn_layer(g["x"].double(), g["h0"].double(), P, "rnns.0", g["switch"], m.weights.detach().double().cpu(),
                        None, None, training=False)
    assert relerr(y, yo) < TOL_OUT
    assert int(m.bn.num_batches_tracked) == 0


@pytest.mark.parametrize("nhid,B,T", [(128, 32, 3), (128, 64, 2), (64, 5, 3)])
def test_rhn_wide_hidden_matches_oracle(backend, monkeypatch, nhid, B, T):
    """Hidden sizes that are multiples of 128 with batches that are multiples of 32 take the tcgen05 weight-gradient
    kernel on the GPU (rhn_wgrad_tc.cuh); (64, 5, 3) is a ragged batch at a hidden size the
    tensor-core forward takes; the fp64 oracle (pinned to the reference by tests/golden/rhn.pt) is the
    checker."""
    if backend.type != "cuda" and B > 32:
        pytest.skip("emulator: one wide case is enough")
    torch.manual_seed(5)
    switch = [[True] * 5 for _ in range(36)]
    m = DARTSCellSearch(nhid, nhid, 0.25, 0.25, switch)
    with torch.no_grad():
        for p in m.parameters():
            p.uniform_(-0.08, 0.08)
    m.weights = torch.nn.Parameter(1e-1 * torch.randn(36, 5))
    m.to(backend).train()
    xm = (torch.rand(B, nhid) < 0.75).float() / 0.75
    hm = (torch.rand(B, nhid) < 0.75).float() / 0.75
    masks = [xm, hm]
    monkeypatch.setattr(R, "mask2d", lambda B_, D, keep, dev: masks.pop(0).to(dev))
    x0, h00, gy = torch.randn(T, B, nhid), 0.1 * torch.randn(B, nhid), torch.randn(T, B, nhid)
    x, h0 = leaf(x0, backend), leaf(h00, backend)
    y, _ = m(x, h0[None])
    (y * gy.to(backend)).sum().backward()
    P = {("rnns.0." + k): v.detach().double().cpu().requires_grad_(True) for k, v in m.state_dict().items()
         if v.is_floating_point() and "running" not in k}
    P.update({"rnns.0.bn.running_mean": torch.zeros(nhid, dtype=torch.float64),
              "rnns.0.bn.running_var": torch.ones(nhid, dtype=torch.float64)})
    al = m.weights.detach().double().cpu().requires_grad_(True)
    xo, ho = x0.double().requires_grad_(True), h00.double().requires_grad_(True)
    yo, _ = O.rhn_layer(xo, ho, P, "rnns.0", switch, al, xm.double(), hm.double(), True, O.Recorder())
    (yo * gy.double()).sum().backward()
    assert relerr(y, yo) < TOL_OUT
    assert relerr(x.grad, xo.grad) < TOL_GRAD
    assert relerr(h0.grad, ho.grad) < TOL_GRAD
    assert relerr(m.weights.grad, al.grad) < TOL_GRAD
    assert relerr(m._W0.grad, P["rnns.0._W0"].grad) < TOL_GRAD
    for i, w in enumerate(m._Ws):
        assert relerr(w.grad, P[f"rnns.0._Ws.{i}"].grad) < TOL_GRAD, i
