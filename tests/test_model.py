"""Whole-supernet parity (tiny configuration on both backends; the B=64 / 4x1000 / 919-label configuration
on the GPU) against runs of the unmodified reference: logits, loss, every gradient tensor, and one full
search step through the reference's own train()/Architect/SGD/Adam (tests/golden/model_*.pt)."""
import math
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

import genomenas_oracle as O
from conftest import GOLDEN, relerr

import genome_nas_b200 as G
import genome_nas_b200.rnn as R


def load(tag):
    return torch.load(os.path.join(GOLDEN, f"model_{tag}.pt"), weights_only=False)


def build(g, device):
    cfg = g["cfg"]
    P = O.synthetic_params(g["shapes"], g["seed_params"])
    mk = lambda k: torch.nn.Parameter(P[k].clone())
    m = G.RNNModelSearch(cfg.seq_len, cfg.dropouth, cfg.dropoutx, cfg.C, cfg.num_classes, cfg.layers, cfg.steps,
                         cfg.multiplier, cfg.stem_multiplier, True, 0.2, None, cfg.task, cfg.switches_normal,
                         cfg.switches_reduce, cfg.switches_rnn, cfg.p_skip, mk("alphas_normal"), mk("alphas_reduce"),
                         mk("weights"))
    assert list(m.state_dict().keys()) == list(g["shapes"].keys())
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(s) for s, _ in g["shapes"].values()]
    m.load_state_dict(P)
    m.dropout_lin.p = 0.0
    return m.to(device).train(), P


def probe_vec(i, n):
    return torch.randn(n, generator=torch.Generator().manual_seed(1000 + i), dtype=torch.float64)


def check_forward_backward(g, device, monkeypatch):
    cfg = g["cfg"]
    m, _ = build(g, device)
    masks = [g["xm"], g["hm"]]
    monkeypatch.setattr(R, "mask2d", lambda B, D, keep, dev: masks.pop(0).to(dev))
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
    xt, yt = xt.to(device), yt.to(device)
    logits, hidden, rnn_hs, _ = m(xt, m.init_hidden(g["B"]), return_h=True)
    raw = F.binary_cross_entropy(logits, yt)
    loss = raw + O.Hyper().beta * (rnn_hs[-1][1:] - rnn_hs[-1][:-1]).pow(2).mean()
    # north star: logits and losses within 1e-3 relative of the reference
    assert relerr(logits, g["logits64"]) < 1e-4
    assert abs(float(loss) - float(g["loss64"])) < 1e-5 * abs(float(g["loss64"]))
    assert abs(float(raw) - float(g["raw64"])) < 1e-5 * abs(float(g["raw64"]))
    loss.backward()
    gn, r32 = g["grad_norm64"], g["ref32_err"]
    total = math.sqrt(sum(v * v for v in gn.values()))
    names = [n for n, _ in m.named_parameters()]
    assert names == list(gn.keys())
    score, ratio = {}, {}  # per tensor: error / limit, error / the reference's own fp32 error
    rb = g.get("ref32b_err")
    for i, (n, p) in enumerate(m.named_parameters()):
        ref = gn[n]
        if ref < 1e-12 * total:
            continue
        # per-tensor rel-L2 against fp64, graded against the reference's own fp32 runs: the north
        # star's 1e-3, or 1.5x the error the reference itself makes on that tensor.  BN statistics are summed in
        # double from the per-warp partials on, so the forward pass (and with it every ReLU / max-pool decision)
        # is reproducible run to run.  The full-size goldens hold the reference's fp32 error for BOTH of its CPU
        # convolution backends (oracle/ref_backend_spread.py: oneDNN on / off); the bar uses the larger of the two.
        own = max(rb["onednn"][n], rb["native"][n]) if rb else r32[n]
        lim = max(1e-3, 1.5 * own)
        if n in g["grad64"]:
            err = relerr(p.grad, g["grad64"][n])
        elif "grad_sketch64" in g:
            # full-size goldens: S random projections of the fp64 gradient per tensor; sqrt(mean_k <g - g_ref, r_k>^2)
            # is an unbiased estimate of ||g - g_ref|| (O.sketch: sigma 6 % at S = 128)
            d = O.sketch(p.grad, i) - g["grad_sketch64"][n]
            err = float(d.pow(2).mean().sqrt()) / ref
        else:
            gd = p.grad.detach().double().cpu()
            dot = float((gd.flatten() * probe_vec(i, gd.numel())).sum())
            # <g, r> with r ~ N(0, I): |error| <~ ||g - g_ref|| up to a chi-distributed factor
            err = max(abs(float(gd.norm()) - ref) / ref, abs(dot - g["grad_dot64"][n]) / (5 * ref))
        score[n] = err / lim
        ratio[n] = err / max(own, 1e-12)
    over = {n: v for n, v in score.items() if v >= 1.0}
    if not rb:
        assert not over, over          # op-level and tiny configurations: every tensor, no allowance
    else:
        # Full-size configurations (38 M parameters, 1,274 tensors).  A ReLU / max-pool decision that sits within
        # fp32 rounding of its threshold falls on either side in any two correct fp32 implementations, and one such
        # element moves the gradient of the tensors upstream of it by ~1/sqrt(N) = 2-3e-3.  The reference's own two
        # backends show it: they differ from each other by a median factor of 2.5 per tensor (p90 6, max 22), one
        # fails the 1.5x bar of the other on 81 % of the tensors.  So the gate is the reference's own distribution:
        #   (1) the median over all tensors of err / (reference's fp32 error on that tensor) is <= 1,
        #   (2) at most 2 % of the tensors exceed their own 1.5x bar (observed 2 of 1,274 / 1 of 450), and
        #   (3) none of them is worse than the reference's own worst tensor.
        import statistics
        worst_ref = max(max(rb["onednn"][n], rb["native"][n]) for n in score)
        assert statistics.median(ratio.values()) <= 1.0, statistics.median(ratio.values())
        assert len(over) <= max(1, len(score) // 50), over
        lims = {n: max(1e-3, 1.5 * max(rb["onednn"][n], rb["native"][n])) for n in over}
        assert all(v * lims[n] <= worst_ref for n, v in over.items()), (over, worst_ref)
    dg = dict(m.named_parameters())["decoder.weight"].grad.detach().double().cpu()
    assert relerr(dg @ g["decoder_proj_vec"].double(), g["decoder_grad_proj64"]) < 2e-3
    return m


@pytest.mark.parametrize("tag", ["tiny", "tiny_pruned"])
def test_supernet_forward_backward(backend, tag, monkeypatch):
    check_forward_backward(load(tag), backend, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["full", "pdarts3"])
def test_supernet_full_size_forward_backward(tag, monkeypatch):
    """BASELINE configs[1]: batch 64, one-hot 4x1000, 919 labels, full supernet (38.0 M parameters); configs[3]: the
    same with the P-DARTS stage-3 switches (3 CNN / 2 RNN candidates per edge)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from genome_nas_b200 import _lib
    _lib._use_library_for_tests(None, "cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    check_forward_backward(load(tag), torch.device("cuda:0"), monkeypatch)


def _args():
    return SimpleNamespace(wdecay=5e-7, clip=0.25, arch_lr=3e-3, arch_wdecay=1e-3)


def _check_step_state(m, g, tol_scale=1.0):
    sd = m.state_dict()
    for n, ref in g["step_alphas"].items():
        assert relerr(sd[n], ref) < 2e-5 * tol_scale, n
    for n, ref in g["step_sample"].items():
        assert relerr(sd[n], ref) < 5e-4 * tol_scale, n
    for n, ref in g["step_state_norm"].items():
        assert abs(float(sd[n].double().norm()) - ref) <= 5e-4 * tol_scale * max(ref, 1e-6), n
    T = m.decoder.in_features // m.nhid
    for n, v in sd.items():
        if n.endswith("num_batches_tracked"):      # two training forwards (valid batch + train batch)
            assert int(v) == (2 * 9 * T if n == "rnns.0.bn.num_batches_tracked" else 2), n


@pytest.mark.parametrize("tag", ["tiny", "tiny_pruned"])
def test_search_step_reference_loop(backend, tag, monkeypatch):
    """The reference's loop body (train() + Architect + torch.optim.SGD/Adam) running on the fused modules."""
    g = load(tag)
    cfg, hyper = g["cfg"], O.Hyper()
    m, _ = build(g, backend)
    mv, mt = g["step_masks"]
    masks = [mv[0], mv[1], mt[0], mt[1]]
    monkeypatch.setattr(R, "mask2d", lambda B, D, keep, dev: masks.pop(0).to(dev))
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
    xv, yv = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 1)
    conv, rhn = [], []
    for name, p in m.named_parameters():
        (rhn if "rnns" in name else conv).append(p)
    opt = torch.optim.SGD([{"params": conv}, {"params": rhn}], lr=hyper.cnn_lr, weight_decay=hyper.cnn_wd)
    opt.param_groups[0]["momentum"] = hyper.momentum
    opt.param_groups[1]["lr"] = hyper.rhn_lr
    opt.param_groups[1]["weight_decay"] = hyper.rhn_wd
    arch = G.Architect(m, torch.nn.BCELoss(), _args())
    labels, preds, avg = G.train([(xt, yt)], [(xv, yv)], m, rhn, conv, torch.nn.BCELoss(), opt, None, arch, False,
                                 hyper.cnn_lr, 0, 0, [5, 0.25, hyper.clip], 1000, hyper.beta, True, True, False,
                                 cfg.task)
    assert abs(float(avg) - float(g["step_loss_train"])) < 1e-4 * abs(float(g["step_loss_train"]))
    assert preds[0].shape == (g["B"], cfg.num_classes)
    _check_step_state(m, g)


@pytest.mark.parametrize("tag", ["tiny"])
def test_search_step_fused_optimisers(backend, tag, monkeypatch):
    """SearchStep: same step with gnas_adam_step / gnas_grad_sumsq / gnas_sgd_step on the flat buffers."""
    g = load(tag)
    cfg = g["cfg"]
    m, _ = build(g, backend)
    mv, mt = g["step_masks"]
    masks = [mv[0], mv[1], mt[0], mt[1]]
    monkeypatch.setattr(R, "mask2d", lambda B, D, keep, dev: masks.pop(0).to(dev))
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
    xv, yv = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 1)
    stepper = G.SearchStep(m)
    la, lw, raw, logits = stepper.step(xt.to(backend), yt.to(backend), xv.to(backend), yv.to(backend))
    assert abs(float(la) - float(g["step_loss_valid"])) < 1e-4 * abs(float(g["step_loss_valid"]))
    assert abs(float(lw) - float(g["step_loss_train"])) < 1e-4 * abs(float(g["step_loss_train"]))
    _check_step_state(m, g)


def _hyper_of(g):
    h = G.Hyper()
    if g.get("pdarts"):      # P-DARTS / CWP-DARTS inline alpha step (train_and_validate.py:87-101)
        h.arch_lr, h.adam_betas, h.arch_clip = 6e-4, (0.5, 0.999), h.clip
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["eager", "graph"])
@pytest.mark.parametrize("tag", ["full", "pdarts3"])
def test_full_size_search_step(tag, mode, monkeypatch):
    """BASELINE configs[1] (full supernet) and configs[3] (P-DARTS stage-3 switches from the reference's own
    discard_* functions, inline clipped Adam): one B=64 search step through SearchStep, launched kernel by
    kernel and replayed from the captured CUDA graph, against the state the reference's own train() left
    behind (tests/golden/model_<tag>.pt: step_alphas, step_sample, step_state_norm)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from genome_nas_b200 import _lib
    _lib._use_library_for_tests(None, "cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = load(tag)
    cfg = g["cfg"]
    m, _ = build(g, dev)
    stepper = G.SearchStep(m, hyper=_hyper_of(g))
    mv, mt = g["step_masks"]
    seq = [mv[0], mv[1], mt[0], mt[1]]
    xt, yt = [t.to(dev) for t in O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)]
    xv, yv = [t.to(dev) for t in O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 1)]
    if mode == "eager":
        masks = list(seq)
        monkeypatch.setattr(R, "mask2d", lambda B, D, keep, d: masks.pop(0).to(d))
        la, lw, raw, logits = stepper.step(xt, yt, xv, yv)
    else:
        # capture wants the lazily initialised state (kernel attributes, allocator pools) of one earlier eager
        # step: run it, put the golden initial state back, then capture and replay ONCE with the golden masks
        snap = {k: v.clone() for k, v in m.state_dict().items()}
        stepper.step(xt, yt, xv, yv)
        m.load_state_dict(snap)
        for t in (stepper.mom, stepper.adam_m, stepper.adam_v, stepper.t_adam):
            t.zero_()
        stepper.mask_source = lambda i, B, H, keep: seq[i]
        stepper.capture(xt, yt, xv, yv, warmup=0)
        la, lw, raw, logits = stepper.step(xt, yt, xv, yv)
    torch.cuda.synchronize()
    assert abs(float(la) - float(g["step_loss_valid"])) < 1e-4 * abs(float(g["step_loss_valid"]))
    assert abs(float(lw) - float(g["step_loss_train"])) < 1e-4 * abs(float(g["step_loss_train"]))
    # full size: the reference's own fp32 step and its fp64-exact restatement differ by up to 8e-4 on the
    # near-zero running means (make_golden.py: step_oracle_vs_ref_worst) -> 4x the tiny-model tolerances
    _check_step_state(m, g, tol_scale=4.0)


def test_seeded_construction_matches_reference_initialisation():
    """Same constructor calls in the same order as the reference => identical initial weights for a seed.
    Known answer from SURVEY.md 8(c): sum |params| = 800971.928 for seed 3 on the full supernet."""
    import numpy as np
    torch.manual_seed(3)
    np.random.seed(3)
    a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s))) for s in ((14, 9), (14, 9), (36, 5))]
    sw = lambda n, k: [[True] * k for _ in range(n)]
    m = G.RNNModelSearch(1000, 0.0, 0.0, 8, 919, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sw(14, 9), sw(14, 9),
                         sw(36, 5), 0.0, *a)
    assert sum(p.numel() for p in m.parameters()) == 38035471
    assert len(m.state_dict()) == 3933
    total = sum(float(p.detach().double().abs().sum()) for p in m.parameters())
    assert abs(total - 800971.928) < 0.05
    assert torch.allclose(m.alphas_normal[0, :3].detach(), torch.tensor([1.78863e-3, 4.36510e-4, 9.64975e-5]), rtol=1e-4)


def test_genotype_matches_reference():
    """genotype() against the reference's own derivation (model_search.py:164-243) on recorded alphas
    (tests/golden/genotype.json, written by oracle/make_golden.py:golden_genotype)."""
    import json
    gold = json.load(open(os.path.join(GOLDEN, "genotype.json")))
    g = load("tiny")
    m, _ = build(g, torch.device("cpu"))
    assert len(gold["cases"]) >= 6
    for case in gold["cases"]:
        with torch.no_grad():
            m.alphas_normal.copy_(torch.tensor(case["alphas_normal"]))
            m.alphas_reduce.copy_(torch.tensor(case["alphas_reduce"]))
            m.weights.copy_(torch.tensor(case["weights"]))
        geno = m.genotype()
        for field in ("normal", "reduce", "rnn"):
            assert [[op, int(j)] for op, j in getattr(geno, field)] == case[field], (case["seed"], field)
        for field in ("normal_concat", "reduce_concat", "rnn_concat"):
            assert list(getattr(geno, field)) == case[field], (case["seed"], field)


def test_genotype_and_api_surface():
    g = load("tiny")
    m, _ = build(g, torch.device("cpu"))
    geno = m.genotype()
    assert len(geno.normal) == 8 and len(geno.reduce) == 8 and len(geno.rnn) == 8
    assert all(op != "none" for op, _ in geno.normal + geno.reduce + geno.rnn)
    assert list(geno.normal_concat) == [2, 3, 4, 5] and list(geno.rnn_concat) == list(range(1, 9))
    assert [p.data_ptr() for p in m.arch_parameters()] == [m.alphas_normal.data_ptr(), m.alphas_reduce.data_ptr(),
                                                           m.weights.data_ptr()]
    assert m.rnns[0].weights is m.weights
    new = m.new()
    assert all(torch.equal(a, b) for a, b in zip(new.arch_parameters(), m.arch_parameters()))
    h = m.init_hidden(5)
    assert len(h) == 1 and tuple(h[0].shape) == (1, 5, m.nhid)


def test_unrolled_architect_matches_reference(backend):
    """Second-order Architect.step (architect.py:55-129: virtual SGD step, dalpha at w', finite-difference
    Hessian-vector product) against an fp64 run of the reference (tests/golden/unrolled.pt)."""
    g = torch.load(os.path.join(GOLDEN, "unrolled.pt"), weights_only=False)
    cfg, B = g["cfg"], g["B"]
    m, _ = build(g, backend)
    orig_new = m.new

    def new_nodrop():          # as in the golden run: the virtual-step copy without the head dropout (device-specific masks)
        m2 = orig_new()
        m2.dropout_lin.p = 0.0
        return m2
    m.new = new_nodrop
    xt, yt = [t.to(backend) for t in O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 0)]
    xv, yv = [t.to(backend) for t in O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 1)]
    opt = torch.optim.SGD(m.parameters(), lr=g["eta"])
    arch = G.Architect(m, torch.nn.BCELoss(), _args())
    before = [p.detach().clone() for p in m.arch_parameters()]
    arch.step(m.init_hidden(B), xt, yt, m.init_hidden(B), xv, yv, opt, True)
    for p, b, d64, d32, a64 in zip(m.arch_parameters(), before, g["dalpha64"], g["dalpha32"], g["alpha64"]):
        # the finite-difference term loses digits in fp32: grade against the reference's own fp32 error
        lim = max(2e-3, 3.0 * relerr(d32, d64))
        assert relerr(p.grad, d64) < lim
        assert relerr(p.detach() - b, a64 - b.double().cpu()) < max(5e-2, 10 * lim)      # Adam: first step ~ lr * sign(g)
    # the weights themselves are untouched by the architecture step (perturbed by +R, -2R, +R)
    P = O.synthetic_params(g["shapes"], g["seed_params"])
    assert relerr(dict(m.named_parameters())["stem.0.weight"], P["stem.0.weight"]) < 1e-6


def test_infer_matches_oracle_eval_mode(backend):
    """infer() (train_and_validate.py:182-232): eval-mode forward (running statistics, no masks) and the
    sample-weighted mean BCE over the queue, against the CPU oracle in fp64."""
    g = load("tiny")
    cfg = g["cfg"]
    m, P = build(g, backend)
    batches = [O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, s) for s in (3, 4)]
    labels, preds, avg = G.infer(batches, m, torch.nn.BCELoss(), g["B"], 10, 1000, cfg.task)
    assert not m.training and len(preds) == 2
    Q = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
    tot = 0.0
    for (x, y), p in zip(batches, preds):
        lo, _ = O.supernet_forward(Q, cfg, x.double(), None, None, False, O.Recorder())
        assert relerr(torch.from_numpy(p), lo) < 1e-5
        tot += float(F.binary_cross_entropy(lo, y.double()))
    assert abs(float(avg) - tot / 2) < 1e-5 * (tot / 2)


@pytest.mark.gpu
def test_side_streams_do_not_change_results(monkeypatch):
    """Independent kernel chains of a pass overlap on the library's side streams (fork / join by events inside every C
    call).  A missing dependency would show as run-to-run garbage: at full size (where the kernels are long enough to
    overlap for real) two overlapped passes and one serialised pass must agree -- logits to 1e-6 (BN statistics are
    summed in double), gradients up to the order of the float atomics."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from genome_nas_b200 import _lib
    _lib._use_library_for_tests(None, "cuda")
    g = load("full")
    cfg = g["cfg"]
    dev = torch.device("cuda:0")
    m, _ = build(g, dev)
    xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
    xt, yt = xt.to(dev), yt.to(dev)
    runs = []
    for side in (True, True, False):
        prev = _lib.side_streams(side)
        try:
            masks = [g["xm"], g["hm"]]
            monkeypatch.setattr(R, "mask2d", lambda B, D, keep, d: masks.pop(0).to(d))
            for p in m.parameters():
                p.grad = None
            logits, hidden, rnn_hs, _ = m(xt, m.init_hidden(g["B"]), return_h=True)
            loss = G.bce_loss(logits, yt) + G.tar_loss(rnn_hs[-1], O.Hyper().beta)
            loss.backward()
            torch.cuda.synchronize()
            runs.append((logits.detach().clone(), [p.grad.detach().clone() for p in m.parameters()]))
        finally:
            _lib.side_streams(prev)
        # undo the running-statistics side effect so that every run starts from the same state (not needed for the
        # training-mode arithmetic, kept for clarity)
    base_logits, base_grads = runs[-1]
    for logits, grads in runs[:-1]:
        assert relerr(logits, base_logits) < 1e-6
        worst = max(relerr(a, b) for a, b in zip(grads, base_grads) if float(b.abs().max()) > 0)
        assert worst < 1e-4, worst
