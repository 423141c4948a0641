"""Edge cases of the hot path (SURVEY.md section 4 iv): ragged batch sizes, odd lengths (125 -> 42 with stride 3),
single-candidate edges, skip-connect dropout (P-DARTS), against the CPU oracle in fp64."""
import pytest
import torch

import genomenas_oracle as O
from conftest import leaf, relerr

import genome_nas_b200 as G


def _shapes(module, prefix):
    kinds = {"running_mean": "rm", "running_var": "rv", "num_batches_tracked": "nbt"}
    return {prefix + k: (tuple(v.shape), next((t for s, t in kinds.items() if k.endswith(s)), "w"))
            for k, v in module.state_dict().items()}


def _oracle_mixed(m, x, w, switch, stride, P, skip_mask=None, p=0.0):
    Q = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xo, wo = x.double().clone().requires_grad_(True), w.double().clone().requires_grad_(True)
    y = O.mixed_op(xo, wo, switch, stride, Q, "cell_ops.0", True, O.Recorder(), p,
                   {"cell_ops.0": skip_mask.double()} if skip_mask is not None else None)
    return y, xo, wo, Q


@pytest.mark.parametrize("B,C,L,stride,switch", [
    (1, 8, 17, 1, [True] * 9),                                   # single row, L smaller than any tile
    (5, 8, 125, 3, [True] * 9),                                  # 125 -> 42 (cell 5), ragged batch
    (9, 16, 43, 2, [True] * 9),                                  # odd length, batch not a multiple of 8
    (3, 8, 30, 1, [False, False, False, False, False, False, False, False, True]),   # conv_15 only
    (3, 8, 30, 2, [False, False, False, True, False, False, False, False, False]),   # strided skip only
    (4, 8, 64, 1, [True, False, False, True, False, False, False, False, False]),    # none + identity
    (2, 32, 250, 1, [False, True, True, False, True, False, False, True, False]),
    # conv_15 at the channel counts the tcgen05 kernel takes (C = 32 / 64 / 128, stride-1 convolutions):
    (3, 32, 250, 1, [False] * 8 + [True]),                       # two position tiles, ragged tail
    (2, 32, 300, 1, [True] + [False] * 7 + [True]),              # three tiles
    (2, 64, 125, 1, [False] * 8 + [True]),                       # one partial tile (cell 4 geometry)
    (3, 128, 42, 1, [False] * 8 + [True]),                       # cell 5 geometry
    (2, 64, 125, 2, [False] * 8 + [True]),                       # strided first conv (SIMT) + stride-1 second conv
    # rows longer than one depthwise / pool segment (256 positions at stride 1, 252 at stride 2 / 3):
    (2, 8, 1000, 1, [True] * 9),                                 # cell 0 geometry: 4 segments
    (2, 8, 530, 2, [True] * 9),                                  # 3 segments, ragged last one, 530 -> 265
    (9, 8, 520, 3, [False, True, True, True, True, True, True, True, False]),    # 3 segments at stride 3, 520 -> 174
])
def test_mixedop_shapes(backend, B, C, L, stride, switch):
    torch.manual_seed(B * 100 + L)
    m = G.MixedOp(C, stride, switch, 0.0)
    P = O.synthetic_params(_shapes(m, "cell_ops.0."), 5)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    x0 = torch.randn(B, C, L)
    w0 = torch.softmax(torch.randn(sum(switch)), 0)
    x, w = leaf(x0, backend), leaf(w0, backend)
    y = m(x, w)
    gy = torch.randn(y.shape)
    (y * gy.to(backend)).sum().backward()
    yo, xo, wo, Q = _oracle_mixed(m, x0, w0, switch, stride, P)
    (yo * gy.double()).sum().backward()
    assert y.shape == yo.shape
    assert relerr(y, yo) < 1e-5
    assert relerr(x.grad, xo.grad) < 5e-5
    assert relerr(w.grad, wo.grad) < 5e-5
    for n, p in m.named_parameters():
        assert relerr(p.grad, Q["cell_ops.0." + n].grad) < 5e-5, n


@pytest.mark.parametrize("stride", [1, 2, 3])
def test_fused_edge_kernels_every_candidate_subset(backend, stride):
    """k_edge_fwd / k_edge_bwd run all enabled depthwise candidates of an edge in one pass: every non-empty subset of
    {sep_conv_15, sep_conv_9, dil_conv_15, dil_conv_9} (P-DARTS / CWP pruning leaves arbitrary ones) against the oracle."""
    B, C, L = 3, 8, 70
    torch.manual_seed(11 + stride)
    x0 = torch.randn(B, C, L)
    for sub in range(1, 16):
        switch = [False] * 9
        for v in range(4):
            switch[4 + v] = bool(sub >> v & 1)
        m = G.MixedOp(C, stride, switch, 0.0)
        P = O.synthetic_params(_shapes(m, "cell_ops.0."), 100 + sub)
        m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
        m.to(backend).train()
        w0 = torch.softmax(torch.randn(sum(switch)), 0)
        x, w = leaf(x0, backend), leaf(w0, backend)
        y = m(x, w)
        gy = torch.randn(y.shape)
        (y * gy.to(backend)).sum().backward()
        yo, xo, wo, Q = _oracle_mixed(m, x0, w0, switch, stride, P)
        (yo * gy.double()).sum().backward()
        assert relerr(y, yo) < 1e-5, sub
        assert relerr(x.grad, xo.grad) < 5e-5, sub
        assert relerr(w.grad, wo.grad) < 5e-5, sub
        for n, p in m.named_parameters():
            assert relerr(p.grad, Q["cell_ops.0." + n].grad) < 5e-5, (sub, n)


def test_skip_connect_dropout_pdarts(backend):
    """P-DARTS: Identity -> Dropout(p) (model_discCNN.py:60-61).  The mask is drawn by the module; the kernels
    must apply exactly that mask in forward, dX and dmix."""
    switch = [True, True, False, True, False, True, False, False, False]
    m = G.MixedOp(8, 1, switch, 0.3)
    assert isinstance(m.m_ops.skip_connect, torch.nn.Sequential) and m.m_ops.skip_connect[1].p == 0.3
    P = O.synthetic_params(_shapes(m, "cell_ops.0."), 6)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    x0, w0 = torch.randn(4, 8, 50), torch.softmax(torch.randn(4), 0)
    x, w = leaf(x0, backend), leaf(w0, backend)
    captured = {}
    orig = m._draw_skip_masks

    def spy(desc, like):
        masks, arr = orig(desc, like)
        captured["mask"] = masks[0].detach().cpu().clone()
        return masks, arr

    m._draw_skip_masks = spy
    y = m(x, w)
    gy = torch.randn(y.shape)
    (y * gy.to(backend)).sum().backward()
    mask = captured["mask"]
    assert all(v == 0.0 or abs(v - 1 / 0.7) < 1e-6 for v in torch.unique(mask).tolist())
    assert 0.1 < float((mask == 0).float().mean()) < 0.5
    yo, xo, wo, _ = _oracle_mixed(m, x0, w0, switch, 1, P, mask, 0.3)
    (yo * gy.double()).sum().backward()
    assert relerr(y, yo) < 1e-5 and relerr(x.grad, xo.grad) < 5e-5 and relerr(w.grad, wo.grad) < 5e-5
    m.p = 0.1
    m.update_p()
    assert m.m_ops.skip_connect[1].p == 0.1


def test_gradient_accumulation_semantics(backend):
    """Two backward passes without zeroing accumulate into p.grad like autograd does."""
    sw = [True] * 9
    m = G.MixedOp(8, 2, sw, 0.0)
    P = O.synthetic_params(_shapes(m, "cell_ops.0."), 7)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    x, w = torch.randn(2, 8, 40, device=backend), torch.softmax(torch.randn(9, device=backend), 0)
    m(x, w).sum().backward()
    once = {n: p.grad.clone() for n, p in m.named_parameters()}
    m(x, w).sum().backward()
    for n, p in m.named_parameters():
        assert relerr(p.grad, 2 * once[n]) < 1e-5, n
    m.zero_grad(set_to_none=True)
    m(x, w).sum().backward()
    for n, p in m.named_parameters():
        assert relerr(p.grad, once[n]) < 1e-5, n


def test_stem_full_width_alignment(backend):
    """Stem at the BASELINE width (24 output channels): the weight-gradient kernel splits positions over
    10 thread groups, which once produced rows that were not 16-byte aligned."""
    from genome_nas_b200.model import _StemFunction
    torch.manual_seed(0)
    stem = torch.nn.Sequential(torch.nn.Conv1d(4, 24, 13, padding=6, bias=False), torch.nn.BatchNorm1d(24)).to(backend)
    x = torch.zeros(3, 4, 200)
    x.scatter_(1, torch.randint(0, 4, (3, 1, 200)), 1.0)
    y = _StemFunction.apply(stem, x.to(backend), stem[0].weight)
    gy = torch.randn(y.shape)
    (y * gy.to(backend)).sum().backward()
    ref = torch.nn.Sequential(torch.nn.Conv1d(4, 24, 13, padding=6, bias=False), torch.nn.BatchNorm1d(24)).double()
    ref.load_state_dict({k: v.detach().cpu() for k, v in stem.state_dict().items()})
    with torch.no_grad():
        ref[1].running_mean.zero_(); ref[1].running_var.fill_(1.0)
    yr = ref(x.double())
    (yr * gy.double()).sum().backward()
    assert relerr(y, yr) < 1e-5
    for (n, p), (_, q) in zip(stem.named_parameters(), ref.named_parameters()):
        assert relerr(p.grad, q.grad) < 5e-5, n


# ---------------------------------------------------------------- shapes that take the tcgen05 kernels on the GPU
def _shapes_cell(module, prefix):
    return _shapes(module, prefix)


@pytest.mark.parametrize("C,L,stride", [(32, 60, 1), (64, 45, 2)])
def test_wide_mixedop_eval_mode(backend, C, L, stride):
    """eval(): running statistics, no batch statistics accumulated (acc = null in the conv epilogues)."""
    switch = [True] * 9
    m = G.MixedOp(C, stride, switch, 0.0)
    P = O.synthetic_params(_shapes(m, "cell_ops.0."), 8)
    for k in list(P):
        if k.endswith("running_var"):
            P[k] = P[k] * 1.3
        if k.endswith("running_mean"):
            P[k] = P[k] + 0.02
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).eval()
    torch.manual_seed(C + L)
    x0, w0 = torch.randn(3, C, L), torch.softmax(torch.randn(9), 0)
    with torch.no_grad():
        y = m(x0.to(backend), w0.to(backend))
        Q = {k: v.double() if v.is_floating_point() else v for k, v in P.items()}
        yo = O.mixed_op(x0.double(), w0.double(), switch, stride, Q, "cell_ops.0", training=False)
    assert relerr(y, yo) < 1e-5
    assert all(int(v) == 0 for k, v in m.state_dict().items() if k.endswith("num_batches_tracked"))


def test_wide_mixedop_alpha_only_pass(backend):
    """need_wgrad = 0 (Architect step): dX and d(mix weights) only, no weight-gradient kernel runs."""
    switch = [True] * 9
    m = G.MixedOp(32, 1, switch, 0.0)
    P = O.synthetic_params(_shapes(m, "cell_ops.0."), 9)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    m._wgrad_enabled = False
    torch.manual_seed(77)
    x0, w0 = torch.randn(5, 32, 42), torch.softmax(torch.randn(9), 0)      # 42 positions: batch rows packed per tile
    x, w = leaf(x0, backend), leaf(w0, backend)
    y = m(x, w)
    gy = torch.randn(y.shape)
    (y * gy.to(backend)).sum().backward()
    yo, xo, wo, Q = _oracle_mixed(m, x0, w0, switch, 1, P)
    (yo * gy.double()).sum().backward()
    assert relerr(y, yo) < 1e-5
    assert relerr(x.grad, xo.grad) < 5e-5
    assert relerr(w.grad, wo.grad) < 5e-5
    assert all(p.grad is None for p in m.parameters())


def test_wide_search_cell_matches_oracle(backend):
    """A reduction cell at C = 32 (strided SIMT first convs + tcgen05 second convs / pointwise / weight gradients)."""
    if backend.type != "cuda":
        pytest.skip("emulator: covered by the small golden cells")
    torch.manual_seed(11)
    sw = [[True] * 9 for _ in range(14)]
    m = G.CNN_Cell_search(4, 4, 48, 64, 32, True, False, sw, 0.0, False)
    shapes = _shapes(m, "cells.0.")
    P = O.synthetic_params(shapes, 12)
    m.load_state_dict({k[len("cells.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    B, L = 3, 90
    s00, s10 = torch.randn(B, 48, L), torch.randn(B, 64, L)
    al0 = 0.5 * torch.randn(14, 9)
    s0, s1, al = leaf(s00, backend), leaf(s10, backend), leaf(al0, backend)
    y = m(s0, s1, torch.softmax(al, -1))
    gy = torch.randn(y.shape)
    (y * gy.to(backend)).sum().backward()
    Q = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    s0o, s1o, alo = s00.double().requires_grad_(True), s10.double().requires_grad_(True), al0.double().requires_grad_(True)
    geom = O.CellGeom(index=0, C_pp=48, C_p=64, C=32, reduction=True, reduction_prev=False, stride=2)
    yo = O.cell_forward(s0o, s1o, torch.softmax(alo, -1), geom, sw, Q, "cells.0", training=True, rec=O.Recorder())
    (yo * gy.double()).sum().backward()
    assert relerr(y, yo) < 1e-5
    assert relerr(s0.grad, s0o.grad) < 5e-5
    assert relerr(s1.grad, s1o.grad) < 5e-5
    assert relerr(al.grad, alo.grad) < 5e-5
    for n, p in m.named_parameters():
        assert relerr(p.grad, Q["cells.0." + n].grad) < 5e-5, n


def test_compact_base_codes_input(backend):
    """uint8 base codes [B, L] (0..3, 255 = N) expanded on the device give the same one-hot tensor the reference's loader
    builds on the host, and the supernet takes them as input."""
    torch.manual_seed(2)
    B, L = 3, 37
    codes = torch.randint(0, 4, (B, L))
    codes[0, 5] = 255
    x = torch.zeros(B, 4, L)
    for b in range(B):
        for l in range(L):
            if codes[b, l] < 4:
                x[b, codes[b, l], l] = 1.0
    enc = G.encode_onehot(x)
    assert torch.equal(enc, codes.to(torch.uint8))
    assert torch.equal(G.expand_onehot(enc.to(backend)).cpu(), x)
