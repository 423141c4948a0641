"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from
/root/reference, CPU) and, in the same run, check that the oracle restatement
(oracle/genomenas_oracle.py) reproduces it in fp32 and fp64.

Run in the authoring container only:  python oracle/make_golden.py [--full]
The GPU box has no /root/reference; tests replay the committed vectors.
The only shim is an empty `matplotlib` module (generalNAS_tools/utils.py:38
imports it; it is not installed and never used on this path).
"""
import argparse
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
for _m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.path.insert(0, os.environ.get("GNAS_REFERENCE", "/root/reference"))

import genomenas_oracle as O  # noqa: E402
import model_search as ref_ms  # noqa: E402
from darts_tools import model_discCNN as ref_md  # noqa: E402
from darts_tools.architect import Architect as RefArchitect  # noqa: E402
from generalNAS_tools import train_and_validate as ref_tv  # noqa: E402


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def leaf(t, dt):
    return t.detach().to(dt).clone().requires_grad_(True)


def probe_vec(i, n):
    return torch.randn(n, generator=torch.Generator().manual_seed(1000 + i), dtype=torch.float64)


def kind_of(name):
    if name.endswith("running_mean"):
        return "rm"
    if name.endswith("running_var"):
        return "rv"
    if name.endswith("num_batches_tracked"):
        return "nbt"
    if name.startswith("alphas_") or name in ("weights", "rnns.0.weights"):
        return "alpha"
    if name == "stem.1.weight":
        return "gamma"
    if name.endswith(".bias"):
        return "bias"
    return "w"


def shapes_of(module):
    return {k: (tuple(v.shape), kind_of(k)) for k, v in module.state_dict().items()}


def load_into(module, P):
    sd = module.state_dict()
    with torch.no_grad():
        for k in sd:
            sd[k].copy_(P[k].to(sd[k].dtype))


def pruned_switches(n_edges, n_prims, keep, seed):
    """Random switch matrix with `keep` enabled primitives per edge (what
    discard_operations.py produces after pruning: same count per edge)."""
    rng = np.random.RandomState(seed)
    sw = []
    for _ in range(n_edges):
        row = [False] * n_prims
        for k in rng.choice(n_prims, keep, replace=False):
            row[int(k)] = True
        sw.append(row)
    return sw


# ------------------------------------------------------------------ MixedOp
def golden_mixedop():
    out = {}
    cases = [("s1", 8, 50, 3, 1, [True] * 9),
             ("s2", 16, 37, 2, 2, [True] * 9),
             ("s3", 8, 41, 2, 3, [True] * 9),
             ("s1_pruned", 8, 33, 2, 1, [False, True, False, True, False, True, False, False, True]),
             ("s2_pruned", 8, 30, 2, 2, [True, False, True, True, True, False, True, False, False])]
    for tag, C, L, B, stride, sw in cases:
        torch.manual_seed(11)
        m = ref_md.MixedOp(C, stride, sw, 0.0)
        shp = {"cell_ops.0." + k: v for k, v in shapes_of(m).items()}
        P = O.synthetic_params(shp, seed=21)
        load_into(m, {k[len("cell_ops.0."):]: v for k, v in P.items()})
        gen = torch.Generator().manual_seed(5)
        x = torch.randn(B, C, L, generator=gen)
        w = torch.softmax(torch.randn(sum(sw), generator=gen), 0)
        res = {}
        for dt in (torch.float32, torch.float64):
            mm = ref_md.MixedOp(C, stride, sw, 0.0).to(dt)
            load_into(mm, {k[len("cell_ops.0."):]: v for k, v in P.items()})
            mm.train()
            xx = leaf(x, dt)
            ww = leaf(w, dt)
            y = mm(xx, ww)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6)).to(dt)
            (y * gy).sum().backward()
            # oracle
            Q = {k: (leaf(v, dt) if v.is_floating_point() else v) for k, v in P.items()}
            xo = leaf(x, dt)
            wo = leaf(w, dt)
            rec = O.Recorder()
            yo = O.mixed_op(xo, wo, sw, stride, Q, "cell_ops.0", True, rec)
            (yo * gy).sum().backward()
            assert relerr(yo, y) < (1e-5 if dt == torch.float32 else 1e-12), (tag, relerr(yo, y))
            assert relerr(xo.grad, xx.grad) < (1e-4 if dt == torch.float32 else 1e-11)
            assert relerr(wo.grad, ww.grad) < (1e-4 if dt == torch.float32 else 1e-11)
            for n, p in mm.named_parameters():
                assert relerr(Q["cell_ops.0." + n].grad, p.grad) < (1e-3 if dt == torch.float32 else 1e-10), n
            sd = mm.state_dict()
            Pn = {k: v.detach() for k, v in Q.items()}
            rec.apply(Pn)
            for k in sd:
                if k.endswith("running_var") or k.endswith("running_mean"):
                    assert relerr(Pn["cell_ops.0." + k], sd[k]) < 1e-5, k
            sfx = "32" if dt == torch.float32 else "64"
            res["y" + sfx] = y.detach()
            res["dx" + sfx] = xx.grad
            res["dw" + sfx] = ww.grad
            res["dW" + sfx] = {"cell_ops.0." + n: p.grad for n, p in mm.named_parameters()}
            res["buf" + sfx] = {"cell_ops.0." + k: v.clone() for k, v in sd.items() if "running" in k}
        out[tag] = dict(C=C, L=L, B=B, stride=stride, switch=sw, shapes=shp, x=x, w=w, gy=gy.float(), **res)
        print("mixedop", tag, "ok")
    torch.save(out, os.path.join(GOLD, "mixedop.pt"))


# --------------------------------------------------------------------- cell
def golden_cell():
    out = {}
    full = [[True] * 9 for _ in range(14)]
    cases = [("normal", 24, 24, 8, False, False, False, 40, 40, 2, full),
             ("reduce", 24, 32, 16, True, False, False, 40, 40, 2, full),
             ("after_reduce", 32, 64, 8, False, True, False, 40, 20, 2, full),
             ("reduce_high", 32, 32, 16, True, False, True, 31, 31, 2, full),
             ("pruned", 16, 16, 8, True, False, False, 26, 26, 2, pruned_switches(14, 9, 3, 7))]
    for tag, Cpp, Cp, C, red, redp, high, L0, L1, B, sw in cases:
        torch.manual_seed(12)
        m = ref_md.CNN_Cell_search(4, 4, Cpp, Cp, C, red, redp, sw, 0.0, high)
        shp = {"cells.0." + k: v for k, v in shapes_of(m).items()}
        P = O.synthetic_params(shp, seed=22)
        gen = torch.Generator().manual_seed(7)
        s0 = torch.randn(B, Cpp, L0, generator=gen)
        s1 = torch.randn(B, Cp, L1, generator=gen)
        n_ops = sum(sw[0])
        alpha = torch.randn(14, n_ops, generator=gen)
        geom = O.CellGeom(0, Cpp, Cp, C, red, redp, 3 if high else (2 if red else 1))
        res = {}
        for dt in (torch.float32, torch.float64):
            mm = ref_md.CNN_Cell_search(4, 4, Cpp, Cp, C, red, redp, sw, 0.0, high).to(dt)
            load_into(mm, {k[len("cells.0."):]: v for k, v in P.items()})
            mm.train()
            a0 = leaf(s0, dt)
            a1 = leaf(s1, dt)
            al = leaf(alpha, dt)
            y = mm(a0, a1, torch.softmax(al, -1))
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(8)).to(dt)
            (y * gy).sum().backward()
            Q = {k: (leaf(v, dt) if v.is_floating_point() else v) for k, v in P.items()}
            b0 = leaf(s0, dt)
            b1 = leaf(s1, dt)
            bl = leaf(alpha, dt)
            yo = O.cell_forward(b0, b1, torch.softmax(bl, -1), geom, sw, Q, "cells.0")
            (yo * gy).sum().backward()
            tol = 2e-5 if dt == torch.float32 else 1e-11
            assert relerr(yo, y) < tol, (tag, relerr(yo, y))
            assert relerr(b0.grad, a0.grad) < 50 * tol and relerr(b1.grad, a1.grad) < 50 * tol
            assert relerr(bl.grad, al.grad) < 50 * tol
            sfx = "32" if dt == torch.float32 else "64"
            res["y" + sfx] = y.detach()
            res["ds0" + sfx], res["ds1" + sfx], res["dalpha" + sfx] = a0.grad, a1.grad, al.grad
            res["dWnorm" + sfx] = {"cells.0." + n: float(p.grad.double().norm()) for n, p in mm.named_parameters()}
            if dt == torch.float64:
                res["dW64"] = {"cells.0." + n: p.grad.float() for n, p in mm.named_parameters()}
        out[tag] = dict(Cpp=Cpp, Cp=Cp, C=C, reduction=red, reduction_prev=redp, high=high, B=B,
                        switches=sw, shapes=shp, s0=s0, s1=s1, alpha=alpha, gy=gy.float(), **res)
        print("cell", tag, "ok")
    torch.save(out, os.path.join(GOLD, "cell.pt"))


# ---------------------------------------------------------------------- RHN
def golden_rhn():
    out = {}
    cases = [("full", 32, 5, 4, [[True] * 5 for _ in range(36)], 0.25, 0.25),
             ("pruned", 32, 4, 6, pruned_switches(36, 5, 2, 9), 0.0, 0.1),
             ("nodrop", 64, 3, 8, [[True] * 5 for _ in range(36)], 0.0, 0.0)]
    for tag, nhid, T, B, sw, dh, dx in cases:
        m = ref_ms.DARTSCellSearch(nhid, nhid, dh, dx, sw)
        shp = {"rnns.0." + k: v for k, v in shapes_of(m).items()}
        n_ops = sum(sw[0])
        shp["weights"] = ((36, n_ops), "alpha")
        P = O.synthetic_params(shp, seed=23)
        gen = torch.Generator().manual_seed(9)
        x = torch.randn(T, B, nhid, generator=gen)
        h0 = 0.3 * torch.randn(B, nhid, generator=gen)
        res = {}
        for dt in (torch.float32, torch.float64):
            mm = ref_ms.DARTSCellSearch(nhid, nhid, dh, dx, sw).to(dt)
            load_into(mm, {k[len("rnns.0."):]: v for k, v in P.items() if k.startswith("rnns.0.") and k != "rnns.0.weights"})
            al = leaf(P["weights"], dt)
            mm.weights = al
            mm.train()
            xx = leaf(x, dt)
            hh = leaf(h0, dt)
            torch.manual_seed(31)
            # mask2d draws fp32 on the CPU generator; in fp64 the reference multiplies fp64 by fp32 masks
            y, last = mm(xx, hh[None])
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(10)).to(dt)
            (y * gy).sum().backward()
            torch.manual_seed(31)
            xm, hm = O.draw_rhn_masks(B, nhid, dx, dh, dt)
            Q = {k: (leaf(v, dt) if v.is_floating_point() else v) for k, v in P.items()}
            xo = leaf(x, dt)
            ho = leaf(h0, dt)
            rec = O.Recorder()
            yo, _ = O.rhn_layer(xo, ho, Q, "rnns.0", sw, Q["weights"], xm, hm, True, rec)
            (yo * gy).sum().backward()
            tol = 2e-5 if dt == torch.float32 else 1e-11
            assert relerr(yo, y) < tol, (tag, relerr(yo, y))
            assert relerr(xo.grad, xx.grad) < 100 * tol, relerr(xo.grad, xx.grad)
            assert relerr(ho.grad, hh.grad) < 100 * tol
            assert relerr(Q["weights"].grad, al.grad) < 100 * tol
            assert relerr(Q["rnns.0._W0"].grad, mm._W0.grad) < 100 * tol
            Pn = {k: v.detach() for k, v in Q.items()}
            rec.apply(Pn)
            assert relerr(Pn["rnns.0.bn.running_var"], mm.bn.running_var) < 1e-5
            sfx = "32" if dt == torch.float32 else "64"
            res["y" + sfx] = y.detach()
            res["dx" + sfx], res["dh0" + sfx], res["dalpha" + sfx] = xx.grad, hh.grad, al.grad
            res["dW0" + sfx] = mm._W0.grad
            res["dWs" + sfx] = [w.grad for w in mm._Ws]
            res["rm" + sfx], res["rv" + sfx] = mm.bn.running_mean.clone(), mm.bn.running_var.clone()
        out[tag] = dict(nhid=nhid, T=T, B=B, switch=sw, dropouth=dh, dropoutx=dx, shapes=shp, x=x, h0=h0,
                        xm=xm.float(), hm=hm.float(), gy=gy.float(), **res)
        print("rhn", tag, "ok")
    torch.save(out, os.path.join(GOLD, "rhn.pt"))


# --------------------------------------------------------------- full model
def build_ref_model(cfg: O.SupernetConfig, P=None, dtype=torch.float32):
    n_cnn = sum(cfg.switches_normal[0])
    n_rnn = sum(cfg.switches_rnn[0])
    a_n = torch.nn.Parameter(torch.zeros(14, n_cnn))
    a_r = torch.nn.Parameter(torch.zeros(14, n_cnn))
    a_w = torch.nn.Parameter(torch.zeros(36, n_rnn))
    m = ref_ms.RNNModelSearch(cfg.seq_len, cfg.dropouth, cfg.dropoutx, cfg.C, cfg.num_classes, cfg.layers,
                              cfg.steps, cfg.multiplier, cfg.stem_multiplier, True, 0.2, None, cfg.task,
                              cfg.switches_normal, cfg.switches_reduce, cfg.switches_rnn, cfg.p_skip,
                              a_n, a_r, a_w)
    if P is not None:
        load_into(m, P)
    if dtype == torch.float64:
        m = m.double()
        # .double() re-creates the alpha Parameters: re-wire the shared references
        m._arch_parameters = [m.alphas_normal, m.alphas_reduce, m.weights]
        for r in m.rnns:
            r.weights = m.weights
    return m


class _Args:
    wdecay = 5e-7
    clip = 0.25
    arch_lr = 3e-3
    arch_wdecay = 1e-3


def ref_search_step(m, hyper, xt, yt, xv, yv, seed, pdarts=False):
    """The reference's own train() for exactly one step, with its own Architect
    and torch.optim.SGD set up as train_genomicDARTS.py:258-292 does."""
    conv, rhn = [], []
    for name, p in m.named_parameters():
        (rhn if "rnns" in name else conv).append(p)
    opt = torch.optim.SGD([{"params": conv}, {"params": rhn}], lr=hyper.cnn_lr, weight_decay=hyper.cnn_wd)
    opt.param_groups[0]["momentum"] = hyper.momentum
    opt.param_groups[1]["lr"] = hyper.rhn_lr
    opt.param_groups[1]["weight_decay"] = hyper.rhn_wd
    arch = RefArchitect(m, torch.nn.BCELoss(), _Args)
    # P-DARTS / CWP-DARTS: the driver owns the alpha optimiser (train_genomicPDARTS.py:300-301)
    opt_a = torch.optim.Adam(m.arch_parameters(), lr=hyper.arch_lr, betas=hyper.adam_betas,
                             weight_decay=hyper.arch_wd) if pdarts else None
    torch.manual_seed(seed)
    ref_tv.train([(xt, yt)], [(xv, yv)], m, rhn, conv, torch.nn.BCELoss(), opt, opt_a, arch, False,
                 hyper.cnn_lr, 0, 0, [5, 0.25, hyper.clip], 1000, hyper.beta, True, True, pdarts, cfg_task)
    return opt, arch


cfg_task = "TF_bindings"


def golden_model(tag, cfg, B, seed_params, full=False, pdarts=False):
    m = build_ref_model(cfg)
    shp = shapes_of(m)
    P = O.synthetic_params(shp, seed_params)
    xt, yt = O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 0)
    xv, yv = O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 1)
    res = dict(cfg=cfg, B=B, seed_params=seed_params, shapes=shp, pdarts=pdarts)
    hyper = O.pdarts_hyper() if pdarts else O.Hyper()
    # ---- forward/backward with explicit masks (dropout_lin off), fp32 + fp64
    torch.manual_seed(77)
    geoms, neurons, nhid = O.cell_geometry(cfg)
    xm, hm = O.draw_rhn_masks(B, nhid, cfg.dropoutx, cfg.dropouth)
    res["xm"], res["hm"] = xm, hm
    grads = {}
    for dt in (torch.float32, torch.float64):
        mm = build_ref_model(cfg, P, dt)
        mm.dropout_lin.p = 0.0
        mm.train()
        torch.manual_seed(77)
        t0 = time.time()
        logits, hidden, rnn_hs, _ = mm(xt.to(dt), mm.init_hidden(B), return_h=True)
        raw = torch.nn.BCELoss()(logits, yt.to(dt))
        loss = raw + hyper.beta * (rnn_hs[-1][1:] - rnn_hs[-1][:-1]).pow(2).mean()
        loss.backward()
        print(f"  ref {tag} {dt} fwd+bwd {time.time() - t0:.1f}s loss {float(loss):.8f}")
        sfx = "32" if dt == torch.float32 else "64"
        res["logits" + sfx], res["loss" + sfx], res["raw" + sfx] = logits.detach(), loss.detach(), raw.detach()
        grads[sfx] = {n: p.grad.detach() for n, p in mm.named_parameters()}
        res["hid" + sfx + "_norm"] = float(rnn_hs[-1].double().norm())
        # oracle cross-check
        Q = {k: (leaf(v, dt) if v.is_floating_point() else v) for k, v in P.items()}
        lo, hid = O.supernet_forward(Q, cfg, xt.to(dt), xm.to(dt), hm.to(dt), True, O.Recorder())
        lso, rawo = O.train_loss(lo, hid, yt.to(dt), hyper.beta)
        names = O.param_names(Q)
        go = O.grads_of(lso, Q, names)
        tol = 5e-5 if dt == torch.float32 else 1e-10
        assert relerr(lo, logits) < tol, relerr(lo, logits)
        assert abs(float(lso) - float(loss)) < tol * abs(float(loss))
        assert names == [n for n, _ in mm.named_parameters()]
        if dt == torch.float64:
            worst = max(relerr(go[n], grads["64"][n]) for n in names if float(grads["64"][n].norm()) > 0)
            print("  oracle64 vs ref64 worst grad rel-L2", worst)
            assert worst < 1e-8
    res["grad_norm64"] = {n: float(g.double().norm()) for n, g in grads["64"].items()}
    res["ref32_err"] = {n: relerr(grads["32"][n], grads["64"][n]) for n in grads["64"]}
    keep = ["alphas_normal", "alphas_reduce", "weights", "stem.0.weight", "stem.1.weight", "stem.1.bias",
            "rnns.0._Ws.0", "classifier.weight", "classifier.bias", "decoder.bias",
            "cells.0.preprocess0.op.1.weight", "cells.5.preprocess1.op.1.weight"]
    for e in (0, 5, 13):
        for c in (0, 1, 5):
            keep += [n for n in grads["64"] if n.startswith(f"cells.{c}.cell_ops.{e}.")]
    if full:
        keep = [n for n in keep if n not in ("classifier.weight", "rnns.0._Ws.0")
                and not (n.startswith("cells.5") and "conv_15" in n)]
    keep = [n for n in dict.fromkeys(keep) if n in grads["64"]]
    res["grad64"] = {n: grads["64"][n].float() for n in keep}
    # every tensor: <g, r_i> with r_i ~ N(0,1) from generator seed 1000+i (direction check without storing g)
    res["grad_dot64"] = {n: float((g.double().flatten() * probe_vec(i, g.numel())).sum())
                         for i, (n, g) in enumerate(grads["64"].items())}
    if full:   # every tensor: a sketch that estimates the actual rel-L2 error (O.sketch)
        res["grad_sketch64"] = {n: O.sketch(g, i) for i, (n, g) in enumerate(grads["64"].items())}
    # a projection of the big decoder gradient
    gen = torch.Generator().manual_seed(3)
    v = torch.randn(grads["64"]["decoder.weight"].shape[1], generator=gen, dtype=torch.float64)
    res["decoder_grad_proj64"] = (grads["64"]["decoder.weight"] @ v).float()
    res["decoder_proj_vec"] = v.float()
    # ---- one full search step through the reference's own train()/Architect (fp32), dropouts as configured
    m2 = build_ref_model(cfg, P)
    m2.dropout_lin.p = 0.0
    t0 = time.time()
    opt, arch = ref_search_step(m2, hyper, xt, yt, xv, yv, seed=88, pdarts=pdarts)
    print(f"  ref {tag} search step {time.time() - t0:.1f}s")
    sd = m2.state_dict()
    # oracle search step with the same RNG stream for the masks (valid pass first, then train pass)
    torch.manual_seed(88)
    mv = O.draw_rhn_masks(B, nhid, cfg.dropoutx, cfg.dropouth)
    mt = O.draw_rhn_masks(B, nhid, cfg.dropoutx, cfg.dropouth)
    Pn, obs = O.search_step({k: v.clone() for k, v in P.items()}, cfg, hyper, xt, yt, xv, yv, {}, mv, mt)
    worst, worst_k = 0.0, None
    for k in sd:
        if sd[k].is_floating_point():
            e = relerr(Pn[k], sd[k])
            if e > worst:
                worst, worst_k = e, k
        else:
            assert int(Pn[k]) == int(sd[k]), k
    print("  oracle search step vs reference train(): worst state rel-L2", worst, worst_k)
    # two fp32 evaluations of an ill-conditioned gradient: fp32 noise floor, not a defect
    assert worst < (5e-3 if full else 2e-4), worst
    res["step_oracle_vs_ref_worst"] = (worst, worst_k)
    res["step_masks"] = (mv, mt)
    res["step_alphas"] = {n: sd[n].clone() for n in ("alphas_normal", "alphas_reduce", "weights")}
    res["step_state_norm"] = {k: float(v.double().norm()) for k, v in sd.items() if v.is_floating_point()}
    res["step_delta_norm"] = {k: float((v.double() - P[k].double()).norm()) for k, v in sd.items() if v.is_floating_point()}
    res["step_sample"] = {k: sd[k].clone() for k in ("stem.0.weight", "stem.1.running_var", "rnns.0.bn.running_mean",
                                                     "rnns.0.bn.running_var", "classifier.bias",
                                                     "cells.2.cell_ops.4.m_ops.sep_conv_9.op.3.running_var",
                                                     "cells.2.cell_ops.4.m_ops.sep_conv_9.op.1.weight")
                          if k in sd}
    res["step_loss_valid"] = obs["loss_valid"]
    res["step_loss_train"] = obs["loss_train"]
    torch.save(res, os.path.join(GOLD, f"model_{tag}.pt"))
    print("model", tag, "ok")


def golden_unrolled():
    """Second-order (unrolled) Architect.step of the reference (architect.py:55-129) on the tiny supernet, dropouts
    off (the four passes of the step would each draw their own RHN masks), fp32 and fp64."""
    cfg = O.SupernetConfig(seq_len=96, C=4, num_classes=10, dropouth=0.0, dropoutx=0.0)
    B = 4
    m0 = build_ref_model(cfg)
    shp = shapes_of(m0)
    P = O.synthetic_params(shp, 71)
    xt, yt = O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 0)
    xv, yv = O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 1)
    res = dict(cfg=cfg, B=B, shapes=shp, seed_params=71, eta=0.025)
    for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
        mm = build_ref_model(cfg, P, dt)
        mm.dropout_lin.p = 0.0
        mm.train()
        # the reference's new() builds a fresh fp32 model with the head dropout on (p = 0.1): the copy gets the dtype of
        # the run and p = 0 (CPU and CUDA draw different masks; the test patches new() of its model the same way)
        orig_new = mm.new

        def new_nodrop(orig_new=orig_new, dt=dt):
            m2 = orig_new().to(dt)
            m2._arch_parameters = [m2.alphas_normal, m2.alphas_reduce, m2.weights]
            for r in m2.rnns:
                r.weights = m2.weights
            m2.dropout_lin.p = 0.0
            return m2
        mm.new = new_nodrop
        opt = torch.optim.SGD(mm.parameters(), lr=res["eta"])
        arch = RefArchitect(mm, torch.nn.BCELoss(), _Args)
        torch.manual_seed(5)
        arch.step(mm.init_hidden(B), xt.to(dt), yt.to(dt), mm.init_hidden(B), xv.to(dt), yv.to(dt), opt, True)
        res["dalpha" + sfx] = [p.grad.detach().clone() for p in mm.arch_parameters()]
        res["alpha" + sfx] = [p.detach().clone() for p in mm.arch_parameters()]
    torch.save(res, os.path.join(GOLD, "unrolled.pt"))
    print("unrolled ok: fp32 vs fp64 dalpha rel err", [relerr(a, b) for a, b in zip(res["dalpha32"], res["dalpha64"])])


def golden_mask_supernet():
    """OSP-NAS / Hyperband mask supernet (randomSearch_and_Hyperband_Tools/model.py, model_search.py): forward +
    backward of the reference's own classes with a supernet mask and a call-time sub-mask, fp32 and fp64."""
    from randomSearch_and_Hyperband_Tools import model_search as ref_hb
    rng = np.random.RandomState(7)
    sup_n = (rng.rand(14, 9) < 0.6).astype(np.int64); sup_n[:, 0] = 0
    sup_r = (rng.rand(14, 9) < 0.6).astype(np.int64); sup_r[:, 0] = 0
    for m in (sup_n, sup_r):
        for e in range(14):
            if m[e].sum() == 0:
                m[e, 1 + rng.randint(8)] = 1
    sup_w = (rng.rand(36, 5) < 0.6).astype(np.int64); sup_w[:, 0] = 0
    call_n = sup_n * (rng.rand(14, 9) < 0.7); call_r = sup_r * (rng.rand(14, 9) < 0.7)
    call_n[3] = 0                                   # an edge with nothing selected contributes zero
    call_w = sup_w * (rng.rand(36, 5) < 0.8)
    for r in range(36):                             # every RHN step keeps at least one selected activation
        if call_w[r].sum() == 0:
            call_w[r, 1 + rng.randint(4)] = 1
    seq_len, C, ncls, B = 96, 4, 10, 4
    res = dict(sup=[sup_n, sup_r, sup_w], call=[call_n.astype(np.int64), call_r.astype(np.int64), call_w.astype(np.int64)],
               seq_len=seq_len, C=C, num_classes=ncls, B=B, dropouth=0.05, dropoutx=0.1)
    args = (seq_len, 0.05, 0.1, C, ncls, 6, 4, 4, 3, True, 0.2, None, "TF_bindings", [sup_n, sup_r, sup_w])
    m0 = ref_hb.RNNModelSearch(*args)
    shp = shapes_of(m0)
    P = O.synthetic_params(shp, 61)
    res["shapes"] = shp
    xt, yt = O.synthetic_batch(B, seq_len, ncls, 0)
    for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
        mm = ref_hb.RNNModelSearch(*args)
        load_into(mm, P)
        if dt == torch.float64:
            mm = mm.double()
        mm.dropout_lin.p = 0.0
        mm.train()
        torch.manual_seed(77)
        logits, hidden, rnn_hs, _ = mm(xt.to(dt), mm.init_hidden(B), res["call"], return_h=True)
        loss = torch.nn.functional.binary_cross_entropy(logits, yt.to(dt))
        loss.backward()
        res["logits" + sfx], res["loss" + sfx] = logits.detach(), loss.detach()
        res["grad" + sfx] = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in mm.named_parameters()}
        res["state" + sfx] = {k: v.detach().clone() for k, v in mm.state_dict().items() if "running" in k or "tracked" in k}
    torch.manual_seed(77)
    res["xm"], res["hm"] = O.draw_rhn_masks(B, 4 * C * 16, 0.1, 0.05)
    # compact fixture: the fp64 truth stored in fp32 (1e-7, far below the test bars) and, instead of the reference's
    # fp32 gradients, their per-tensor error against fp64 (what the test grades against)
    res["ref32_err"] = {k: (relerr(res["grad32"][k], v) if v is not None and res["grad32"][k] is not None else None)
                        for k, v in res["grad64"].items()}
    res["grad64"] = {k: (v.float() if v is not None else None) for k, v in res["grad64"].items()}
    res["state64"] = {k: (v.float() if v.is_floating_point() else v) for k, v in res["state64"].items()}
    res.pop("grad32", None)
    res.pop("state32", None)
    torch.save(res, os.path.join(GOLD, "mask_supernet.pt"))
    print("mask supernet ok: loss32", float(res["loss32"]), "loss64", float(res["loss64"]),
          "params with grad", sum(g is not None for g in res["grad64"].values()), "of", len(res["grad64"]))


def golden_genotype():
    """Genotypes the reference derives (model_search.py:164-243) from recorded alphas."""
    import json
    cfg = O.SupernetConfig(seq_len=96, C=4, num_classes=10)
    m = build_ref_model(cfg)
    cases = []
    for seed in range(8):
        gen = torch.Generator().manual_seed(500 + seed)
        scale = 1e-3 if seed < 2 else 1.0          # search starts from 1e-3 * randn alphas
        a = [scale * torch.randn(s, generator=gen) for s in ((14, 9), (14, 9), (36, 5))]
        with torch.no_grad():
            m.alphas_normal.copy_(a[0]); m.alphas_reduce.copy_(a[1]); m.weights.copy_(a[2])
        geno = m.genotype()
        cases.append({"seed": seed, "alphas_normal": a[0].tolist(), "alphas_reduce": a[1].tolist(),
                      "weights": a[2].tolist(),
                      "normal": [[op, int(j)] for op, j in geno.normal], "reduce": [[op, int(j)] for op, j in geno.reduce],
                      "rnn": [[op, int(j)] for op, j in geno.rnn], "normal_concat": list(geno.normal_concat),
                      "reduce_concat": list(geno.reduce_concat), "rnn_concat": list(geno.rnn_concat)})
    with open(os.path.join(GOLD, "genotype.json"), "w") as fh:
        json.dump({"cases": cases, "how": "oracle/make_golden.py:golden_genotype -- reference RNNModelSearch.genotype()"}, fh)
    print("genotype ok")


def pdarts3_switches():
    """BASELINE configs[3]: stage-3 switches of P-DARTS / CWP-DARTS, produced by the reference's OWN
    discard_cnn_ops / discard_rhn_ops (darts_tools/discard_operations.py:25,117) run twice (9 -> 6 -> 3 CNN,
    5 -> 3 -> 2 RNN candidates per edge) from fixed alphas, exactly as the stage loop of
    train_genomicPDARTS.py:198-203,390-391 does.  Stored as tests/golden/pdarts3_switches.json."""
    import copy
    import json
    from types import SimpleNamespace
    from darts_tools.discard_operations import discard_cnn_ops, discard_rhn_ops
    num_to_keep, num_to_drop = [6, 3, 1], [3, 3, 2]
    num_to_keep_rnn, num_to_drop_rnn = [3, 2, 1], [2, 1, 1]
    sn = [[True] * 9 for _ in range(14)]
    sr = copy.deepcopy(sn)
    sw = [[True] * 5 for _ in range(36)]
    rng = np.random.RandomState(2024)
    for sp in range(2):
        nc, nr = sum(sn[0]), sum(sw[0])
        model = SimpleNamespace(alphas_normal=torch.FloatTensor(rng.randn(14, nc)),
                                alphas_reduce=torch.FloatTensor(rng.randn(14, nc)),
                                weights=torch.FloatTensor(rng.randn(36, nr)))
        _, _, sn, sr = discard_cnn_ops(model, sn, sr, num_to_keep, num_to_drop, sp, new_alpha_values=False)
        _, sw = discard_rhn_ops(model, sw, num_to_keep_rnn, num_to_drop_rnn, sp, new_alpha_values=False)
    assert all(sum(r) == 3 for r in sn + sr) and all(sum(r) == 2 for r in sw)
    out = {"normal": [[bool(x) for x in r] for r in sn], "reduce": [[bool(x) for x in r] for r in sr],
           "rnn": [[bool(x) for x in r] for r in sw],
           "how": "oracle/make_golden.py:pdarts3_switches -- reference discard_cnn_ops/discard_rhn_ops, stages 0 and 1, "
                  "alphas ~ N(0,1) from numpy RandomState(2024)"}
    with open(os.path.join(GOLD, "pdarts3_switches.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also the B=64, 4x1000, 919-label configuration")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    todo = args.only.split(",") if args.only else ["mixedop", "cell", "rhn", "tiny", "tiny_pruned"]
    if "mixedop" in todo:
        golden_mixedop()
    if "cell" in todo:
        golden_cell()
    if "rhn" in todo:
        golden_rhn()
    if "tiny" in todo:
        golden_model("tiny", O.SupernetConfig(seq_len=96, C=4, num_classes=10, dropouth=0.05, dropoutx=0.1), 4, 41)
    if "tiny_pruned" in todo:
        cfg = O.SupernetConfig(seq_len=96, C=4, num_classes=10, dropouth=0.05, dropoutx=0.1,
                               switches_normal=pruned_switches(14, 9, 3, 1),
                               switches_reduce=pruned_switches(14, 9, 3, 2),
                               switches_rnn=pruned_switches(36, 5, 2, 3))
        golden_model("tiny_pruned", cfg, 4, 42)
    if "unrolled" in todo:
        golden_unrolled()
    if "mask" in todo:
        golden_mask_supernet()
    if "genotype" in todo:
        golden_genotype()
    if "pdarts3" in todo:
        sw = pdarts3_switches()
        cfg = O.SupernetConfig(dropouth=0.05, dropoutx=0.1, switches_normal=sw["normal"], switches_reduce=sw["reduce"],
                               switches_rnn=sw["rnn"])
        golden_model("pdarts3", cfg, 64, 44, full=True, pdarts=True)
    if args.full or "full" in todo:
        golden_model("full", O.SupernetConfig(dropouth=0.05, dropoutx=0.1), 64, 43, full=True)


if __name__ == "__main__":
    main()
