"""Mask-based supernet (OSP-NAS / Hyperband / random search) against fp64 runs of the unmodified reference
(randomSearch_and_Hyperband_Tools/model.py, model_search.py; tests/golden/mask_supernet.pt)."""
import os

import pytest
import torch
import torch.nn.functional as F

import genomenas_oracle as O
from conftest import GOLDEN, relerr

import genome_nas_b200.rnn as R
from genome_nas_b200 import mask_supernet as M

G = torch.load(os.path.join(GOLDEN, "mask_supernet.pt"), weights_only=False)


def _build(device):
    g = G
    m = M.RNNModelSearch(g["seq_len"], g["dropouth"], g["dropoutx"], g["C"], g["num_classes"], 6, 4, 4, 3, True, 0.2, None,
                         "TF_bindings", g["sup"])
    assert list(m.state_dict().keys()) == list(g["shapes"].keys())
    P = O.synthetic_params(g["shapes"], 61)
    m.load_state_dict(P)
    m.dropout_lin.p = 0.0
    return m.to(device).train()


def test_mask_supernet_matches_reference(backend, monkeypatch):
    g = G
    m = _build(backend)
    masks = [g["xm"], g["hm"]]
    monkeypatch.setattr(R, "mask2d", lambda B, D, keep, dev: masks.pop(0).to(dev))
    xt, yt = O.synthetic_batch(g["B"], g["seq_len"], g["num_classes"], 0)
    logits, hidden, rnn_hs, _ = m(xt.to(backend), m.init_hidden(g["B"]), g["call"], return_h=True)
    loss = F.binary_cross_entropy(logits, yt.to(backend))
    loss.backward()
    assert relerr(logits, g["logits64"]) < 1e-4
    assert abs(float(loss.detach()) - float(g["loss64"])) < 1e-5 * abs(float(g["loss64"]))
    total = sum(float(v.double().norm()) ** 2 for v in g["grad64"].values() if v is not None) ** 0.5
    worst = 0.0
    for n, p in m.named_parameters():
        ref = g["grad64"][n]
        if ref is None or float(ref.norm()) < 1e-9 * total:      # candidates the call-time mask leaves out
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6 * total, n
            continue
        # tiny model, B = 4: the reference's own fp32 run is 1e-3..1e-2 away from fp64 on these gradients
        lim = max(2e-3, 3.0 * g["ref32_err"][n])
        e = relerr(p.grad, ref)
        worst = max(worst, e / lim)
        assert e < lim, (n, e, lim)
    sd = m.state_dict()
    for k, ref in g["state64"].items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(ref), k                   # only the selected candidates' BatchNorms ran
        else:
            assert relerr(sd[k], ref) < 1e-4 or float((sd[k].double().cpu() - ref).abs().max()) < 1e-6, k


def test_mask_supernet_rejects_masks_outside_the_supernet(backend):
    m = _build(backend)
    call = [G["call"][0].copy(), G["call"][1], G["call"][2]]
    e, k = [(e, k) for e in range(14) for k in range(1, 9) if not G["sup"][0][e][k]][0]
    call[0][e, k] = 1
    xt, _ = O.synthetic_batch(G["B"], G["seq_len"], G["num_classes"], 0)
    with pytest.raises(RuntimeError):
        m(xt.to(backend), m.init_hidden(G["B"]), call)
