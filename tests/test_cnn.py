"""MixedOp / CNN_Cell_search parity: kernels (through the C ABI and the reference-facing module API)
against fp64 runs of the unmodified reference stored in tests/golden (oracle/make_golden.py)."""
import os

import pytest
import torch

import genomenas_oracle as O
from conftest import GOLDEN, leaf, relerr

import genome_nas_b200 as G

# fp32 kernels vs the fp64 reference; the reference's own fp32 run sits at 1e-7..1e-6 here
TOL_OUT, TOL_GRAD = 5e-6, 2e-5

MIXED = torch.load(os.path.join(GOLDEN, "mixedop.pt"))
CELLS = torch.load(os.path.join(GOLDEN, "cell.pt"))


@pytest.mark.parametrize("tag", list(MIXED))
def test_mixedop_matches_reference(backend, tag):
    g = MIXED[tag]
    m = G.MixedOp(g["C"], g["stride"], g["switch"], 0.0)
    P = O.synthetic_params(g["shapes"], 21)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    x = leaf(g["x"], backend)
    w = leaf(g["w"], backend)
    y = m(x, w)
    (y * g["gy"].to(backend)).sum().backward()
    assert relerr(y, g["y64"]) < TOL_OUT
    assert relerr(x.grad, g["dx64"]) < TOL_GRAD
    assert relerr(w.grad, g["dw64"]) < TOL_GRAD
    for n, p in m.named_parameters():
        assert relerr(p.grad, g["dW64"]["cell_ops.0." + n]) < TOL_GRAD, n
    for k, v in m.state_dict().items():
        if "running" in k:
            assert relerr(v, g["buf64"]["cell_ops.0." + k]) < 1e-5, k
        if k.endswith("num_batches_tracked"):
            assert int(v) == 1


@pytest.mark.parametrize("tag", list(CELLS))
def test_cell_matches_reference(backend, tag):
    g = CELLS[tag]
    m = G.CNN_Cell_search(4, 4, g["Cpp"], g["Cp"], g["C"], g["reduction"], g["reduction_prev"], g["switches"], 0.0,
                          g["high"])
    assert list(m.state_dict().keys()) == [k[len("cells.0."):] for k in g["shapes"]]
    P = O.synthetic_params(g["shapes"], 22)
    m.load_state_dict({k[len("cells.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    s0 = leaf(g["s0"], backend)
    s1 = leaf(g["s1"], backend)
    al = leaf(g["alpha"], backend)
    y = m(s0, s1, torch.softmax(al, -1))
    (y * g["gy"].to(backend)).sum().backward()
    assert relerr(y, g["y64"]) < TOL_OUT
    assert relerr(s0.grad, g["ds064"]) < TOL_GRAD
    assert relerr(s1.grad, g["ds164"]) < TOL_GRAD
    assert relerr(al.grad, g["dalpha64"]) < TOL_GRAD
    for n, p in m.named_parameters():
        assert relerr(p.grad, g["dW64"]["cells.0." + n]) < TOL_GRAD, n


def test_alpha_pass_skips_weight_gradients(backend):
    g = CELLS["reduce"]
    m = G.CNN_Cell_search(4, 4, g["Cpp"], g["Cp"], g["C"], g["reduction"], g["reduction_prev"], g["switches"], 0.0,
                          g["high"])
    P = O.synthetic_params(g["shapes"], 22)
    m.load_state_dict({k[len("cells.0."):]: v for k, v in P.items()})
    m.to(backend).train()
    m._wgrad_enabled = False
    s0 = leaf(g["s0"], backend)
    s1 = leaf(g["s1"], backend)
    al = leaf(g["alpha"], backend)
    y = m(s0, s1, torch.softmax(al, -1))
    (y * g["gy"].to(backend)).sum().backward()
    assert relerr(al.grad, g["dalpha64"]) < TOL_GRAD
    assert relerr(s1.grad, g["ds164"]) < TOL_GRAD
    assert all(p.grad is None for p in m.parameters())


def test_eval_mode_uses_running_statistics(backend):
    g = MIXED["s2"]
    m = G.MixedOp(g["C"], g["stride"], g["switch"], 0.0)
    P = O.synthetic_params(g["shapes"], 21)
    for k in list(P):
        if k.endswith("running_var"):
            P[k] = P[k] * 1.7
        if k.endswith("running_mean"):
            P[k] = P[k] + 0.05
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to(backend).eval()
    with torch.no_grad():
        y = m(g["x"].to(backend), g["w"].to(backend))
        Q = {k: v.double() if v.is_floating_point() else v for k, v in P.items()}
        yo = O.mixed_op(g["x"].double(), g["w"].double(), g["switch"], g["stride"], Q, "cell_ops.0", training=False)
    assert relerr(y, yo) < TOL_OUT
    assert all(int(v) == 0 for k, v in m.state_dict().items() if k.endswith("num_batches_tracked"))


def test_bad_arguments_raise(backend):
    g = MIXED["s1"]
    m = G.MixedOp(g["C"], g["stride"], g["switch"], 0.0).to(backend)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, g["C"] + 4, 10, device=backend), g["w"].to(backend))
    with pytest.raises(RuntimeError):
        m(g["x"].to(backend), torch.ones(3, device=backend))
