"""The oracle (CPU restatement) against the golden vectors produced by the unmodified reference."""
import os

import pytest
import torch

import genomenas_oracle as O
from conftest import GOLDEN, relerr


def _leaf(P, dt):
    return {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}


@pytest.mark.parametrize("dt,tol", [(torch.float32, 1e-4), (torch.float64, 1e-11)])
def test_oracle_mixedop(dt, tol):
    gold = torch.load(os.path.join(GOLDEN, "mixedop.pt"))
    sfx = "32" if dt == torch.float32 else "64"
    for tag, g in gold.items():
        Q = _leaf(O.synthetic_params(g["shapes"], 21), dt)
        x = g["x"].to(dt).clone().requires_grad_(True)
        w = g["w"].to(dt).clone().requires_grad_(True)
        y = O.mixed_op(x, w, g["switch"], g["stride"], Q, "cell_ops.0", True, O.Recorder())
        (y * g["gy"].to(dt)).sum().backward()
        assert relerr(y, g["y" + sfx]) < tol, tag
        assert relerr(x.grad, g["dx" + sfx]) < 10 * tol, tag
        assert relerr(w.grad, g["dw" + sfx]) < 10 * tol, tag
        for n, ref in g["dW" + sfx].items():
            assert relerr(Q[n].grad, ref) < 20 * tol, (tag, n)


def test_oracle_cell_fp64():
    gold = torch.load(os.path.join(GOLDEN, "cell.pt"))
    for tag, g in gold.items():
        Q = _leaf(O.synthetic_params(g["shapes"], 22), torch.float64)
        s0 = g["s0"].double().requires_grad_(True)
        s1 = g["s1"].double().requires_grad_(True)
        al = g["alpha"].double().requires_grad_(True)
        geom = O.CellGeom(0, g["Cpp"], g["Cp"], g["C"], g["reduction"], g["reduction_prev"],
                          3 if g["high"] else (2 if g["reduction"] else 1))
        y = O.cell_forward(s0, s1, torch.softmax(al, -1), geom, g["switches"], Q, "cells.0")
        (y * g["gy"].double()).sum().backward()
        assert relerr(y, g["y64"]) < 1e-11, tag
        assert relerr(s0.grad, g["ds064"]) < 1e-10 and relerr(s1.grad, g["ds164"]) < 1e-10, tag
        assert relerr(al.grad, g["dalpha64"]) < 1e-10, tag


def test_oracle_rhn_fp64():
    gold = torch.load(os.path.join(GOLDEN, "rhn.pt"))
    for tag, g in gold.items():
        Q = _leaf(O.synthetic_params(g["shapes"], 23), torch.float64)
        x = g["x"].double().requires_grad_(True)
        h0 = g["h0"].double().requires_grad_(True)
        rec = O.Recorder()
        y, _ = O.rhn_layer(x, h0, Q, "rnns.0", g["switch"], Q["weights"], g["xm"].double(), g["hm"].double(), True, rec)
        (y * g["gy"].double()).sum().backward()
        assert relerr(y, g["y64"]) < 1e-11, tag
        assert relerr(x.grad, g["dx64"]) < 1e-9 and relerr(h0.grad, g["dh064"]) < 1e-9, tag
        assert relerr(Q["weights"].grad, g["dalpha64"]) < 1e-9, tag
        assert relerr(Q["rnns.0._W0"].grad, g["dW064"]) < 1e-9, tag
        P = {k: v.detach() for k, v in Q.items()}
        rec.apply(P)
        assert relerr(P["rnns.0.bn.running_var"], g["rv64"]) < 1e-9, tag


@pytest.mark.parametrize("tag", ["tiny", "tiny_pruned"])
def test_oracle_supernet_fp64(tag):
    g = torch.load(os.path.join(GOLDEN, f"model_{tag}.pt"), weights_only=False)
    cfg = g["cfg"]
    Q = _leaf(O.synthetic_params(g["shapes"], g["seed_params"]), torch.float64)
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0, torch.float64)
    logits, hid = O.supernet_forward(Q, cfg, xt, g["xm"].double(), g["hm"].double(), True, O.Recorder())
    loss, _ = O.train_loss(logits, hid, yt, O.Hyper().beta)
    assert relerr(logits, g["logits64"]) < 1e-10
    assert abs(float(loss) - float(g["loss64"])) < 1e-10
    names = O.param_names(Q)
    grads = O.grads_of(loss, Q, names)
    for i, n in enumerate(names):
        ref = g["grad_norm64"][n]
        assert abs(float(grads[n].norm()) - ref) <= 1e-8 * max(ref, 1e-12), n
    for n, ref in g["grad64"].items():
        assert relerr(grads[n], ref) < 1e-6, n   # stored as fp32


def test_oracle_search_step_matches_reference_train_loop():
    g = torch.load(os.path.join(GOLDEN, "model_tiny.pt"), weights_only=False)
    cfg = g["cfg"]
    P = O.synthetic_params(g["shapes"], g["seed_params"])
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
    xv, yv = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 1)
    mv, mt = g["step_masks"]
    Pn, obs = O.search_step(P, cfg, O.Hyper(), xt, yt, xv, yv, {}, mv, mt)
    for n, ref in g["step_alphas"].items():
        assert relerr(Pn[n], ref) < 1e-5, n
    for n, ref in g["step_sample"].items():
        assert relerr(Pn[n], ref) < 2e-4, n
    for n, ref in g["step_state_norm"].items():
        assert abs(float(Pn[n].double().norm()) - ref) <= 2e-4 * max(ref, 1e-6), n
