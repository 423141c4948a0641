"""CPU oracle for the genomeDARTS supernet search step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may execute it, and there only as
the checker / CPU baseline, never as the thing measured or shipped.

What it is: a functional restatement (plain ``torch.nn.functional`` calls on a
flat ``{name: tensor}`` dictionary that uses the reference's ``state_dict`` key
names) of the arithmetic of the reference hot path.  It runs in whatever dtype
the dictionary holds, so the same code is the fp32 reference and the fp64
ground truth.  Every function cites the reference lines it follows
(paths relative to the reference checkout).

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself, imported unmodified in the authoring container by
``oracle/make_golden.py``; the resulting vectors live in ``tests/golden/`` and
``tests/test_oracle_golden.py`` replays them without the reference present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# generalNAS_tools/genotypes.py:29-48
PRIMS_CNN = ("none", "max_pool_5", "avg_pool_5", "skip_connect", "sep_conv_15",
             "sep_conv_9", "dil_conv_15", "dil_conv_9", "conv_15")
PRIMS_RNN = ("none", "tanh", "identity", "sigmoid", "relu")
RNN_STEPS = 8
BN_EPS = 1e-5
BN_MOM = 0.1


@dataclass
class SupernetConfig:
    """Constructor arguments of RNNModelSearch (model_search.py:113, model_discCNN.py:309-313)."""
    seq_len: int = 1000
    C: int = 8
    num_classes: int = 919
    layers: int = 6
    steps: int = 4
    multiplier: int = 4
    stem_multiplier: int = 3
    dropouth: float = 0.0
    dropoutx: float = 0.0
    p_skip: float = 0.0
    task: str = "TF_bindings"
    switches_normal: List[List[bool]] = field(default_factory=lambda: [[True] * 9 for _ in range(14)])
    switches_reduce: List[List[bool]] = field(default_factory=lambda: [[True] * 9 for _ in range(14)])
    switches_rnn: List[List[bool]] = field(default_factory=lambda: [[True] * 5 for _ in range(36)])


@dataclass
class CellGeom:
    index: int
    C_pp: int
    C_p: int
    C: int
    reduction: bool
    reduction_prev: bool
    stride: int  # stride of the edges reading s0/s1 (1, 2 or 3)


def cell_geometry(cfg: SupernetConfig) -> Tuple[List[CellGeom], int, int]:
    """Channel/stride plan of the cell stack.  model_discCNN.py:349-409."""
    C_curr = cfg.stem_multiplier * cfg.C
    C_pp, C_p, C_curr = C_curr, C_curr, cfg.C
    normal = list(range(cfg.layers))[::3]
    neurons = cfg.seq_len
    red_prev = False
    out: List[CellGeom] = []
    for i in range(cfg.layers):
        if i not in normal:
            C_curr *= 2
            red = True
            if i == 5:
                stride = 3
                neurons = round(neurons / 3)
            else:
                stride = 2
                neurons = round(neurons / 2)
        else:
            red, stride = False, 1
        out.append(CellGeom(i, C_pp, C_p, C_curr, red, red_prev, stride))
        red_prev = red
        C_pp, C_p = C_p, cfg.multiplier * C_curr
    return out, neurons, C_curr * cfg.steps


class Recorder:
    """Collects batch statistics of every BatchNorm use, in call order, so the
    running_mean / running_var side effect (torch BatchNorm1d defaults) can be
    replayed: rm <- 0.9 rm + 0.1 mean ; rv <- 0.9 rv + 0.1 var_unbiased."""

    def __init__(self):
        self.events: List[Tuple[str, torch.Tensor, torch.Tensor]] = []

    def add(self, key, mean, var_unbiased):
        self.events.append((key, mean.detach(), var_unbiased.detach()))

    def apply(self, P: Dict[str, torch.Tensor]):
        for key, m, v in self.events:
            P[key + ".running_mean"] = (1 - BN_MOM) * P[key + ".running_mean"] + BN_MOM * m
            P[key + ".running_var"] = (1 - BN_MOM) * P[key + ".running_var"] + BN_MOM * v
            P[key + ".num_batches_tracked"] = P[key + ".num_batches_tracked"] + 1


def batchnorm(z, P, key, training, rec: Optional[Recorder], gamma=None, beta=None):
    """BatchNorm1d over (B, L) per channel (or over B for 2-D input)."""
    dims = [0, 2] if z.dim() == 3 else [0]
    shape = [1, -1, 1] if z.dim() == 3 else [1, -1]
    if training:
        mean = z.mean(dims)
        var = z.var(dims, unbiased=False)
        n = z.numel() // z.shape[1]
        if rec is not None:
            rec.add(key, mean, var * (n / max(n - 1, 1)))
    else:
        mean = P[key + ".running_mean"].to(z.dtype)
        var = P[key + ".running_var"].to(z.dtype)
    y = (z - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS)
    if gamma is not None:
        y = y * gamma.view(shape) + beta.view(shape)
    return y


# ---------------------------------------------------------------- candidates
def op_pool(x, kind, stride, P, key, training, rec):
    """max_pool_5 / avg_pool_5 wrapped in BN(affine=False); no ReLU.
    operations_14_9.py:21-22, model_discCNN.py:58-59."""
    if kind == "max":
        z = F.max_pool1d(x, 5, stride=stride, padding=2)
    else:
        z = F.avg_pool1d(x, 5, stride=stride, padding=2, count_include_pad=False)
    return batchnorm(z, P, key + ".1", training, rec)


def op_factorized_reduce(x, stride, P, key, training, rec):
    """ReLU -> two strided 1x1 convs on the SAME positions -> cat -> BN.
    operations_14_9.py:280-300 (quirk Q4)."""
    r = F.relu(x)
    z = torch.cat([F.conv1d(r, P[key + ".conv_1.weight"], stride=stride),
                   F.conv1d(r, P[key + ".conv_2.weight"], stride=stride)], dim=1)
    return batchnorm(z, P, key + ".bn", training, rec)


def op_relu_conv_bn(x, P, key, training, rec):
    """ReLU -> 1x1 conv -> BN(affine=False).  operations_14_9.py:124-135."""
    z = F.conv1d(F.relu(x), P[key + ".op.1.weight"])
    return batchnorm(z, P, key + ".op.2", training, rec)


def op_sep_conv(x, k, stride, P, key, training, rec):
    """operations_14_9.py:190-206."""
    C = x.shape[1]
    pad = (k - 1) // 2
    z = F.conv1d(F.relu(x), P[key + ".op.1.weight"], stride=stride, padding=pad, groups=C)
    z = F.conv1d(z, P[key + ".op.2.weight"])
    a = F.relu(batchnorm(z, P, key + ".op.3", training, rec))
    z = F.conv1d(a, P[key + ".op.5.weight"], stride=1, padding=pad, groups=C)
    z = F.conv1d(z, P[key + ".op.6.weight"])
    return batchnorm(z, P, key + ".op.7", training, rec)


def op_dil_conv(x, k, stride, P, key, training, rec):
    """operations_14_9.py:164-176 (dilation 2, padding k-1)."""
    C = x.shape[1]
    z = F.conv1d(F.relu(x), P[key + ".op.1.weight"], stride=stride, padding=k - 1, dilation=2, groups=C)
    z = F.conv1d(z, P[key + ".op.2.weight"])
    return batchnorm(z, P, key + ".op.3", training, rec)


def op_conv15(x, stride, P, key, training, rec):
    """operations_14_9.py:67-81."""
    z = F.conv1d(F.relu(x), P[key + ".op.1.weight"], stride=stride, padding=7)
    a = F.relu(batchnorm(z, P, key + ".op.2", training, rec))
    z = F.conv1d(a, P[key + ".op.4.weight"], stride=1, padding=7)
    return batchnorm(z, P, key + ".op.5", training, rec)


def mixed_op(x, w_row, switch, stride, P, key, training=True, rec=None, p_skip=0.0, skip_mask=None):
    """sum_k w_k * op_k(x) over enabled primitives in PRIMS_CNN order.
    model_discCNN.py:46-93."""
    out = 0
    q = 0
    for k, name in enumerate(PRIMS_CNN):
        if not switch[k]:
            continue
        ok = f"{key}.m_ops.{name}"
        if name == "none":
            y = (x if stride == 1 else x[:, :, ::stride]) * 0.0
        elif name == "max_pool_5":
            y = op_pool(x, "max", stride, P, ok, training, rec)
        elif name == "avg_pool_5":
            y = op_pool(x, "avg", stride, P, ok, training, rec)
        elif name == "skip_connect":
            if stride == 1:
                y = x
                if p_skip > 0 and training:
                    if skip_mask is not None:
                        y = y * skip_mask[key]
                    else:
                        y = F.dropout(y, p_skip, True)
            else:
                y = op_factorized_reduce(x, stride, P, ok, training, rec)
        elif name.startswith("sep_conv"):
            y = op_sep_conv(x, int(name.split("_")[-1]), stride, P, ok, training, rec)
        elif name.startswith("dil_conv"):
            y = op_dil_conv(x, int(name.split("_")[-1]), stride, P, ok, training, rec)
        else:
            y = op_conv15(x, stride, P, ok, training, rec)
        out = out + w_row[q] * y
        q += 1
    return out


def cell_forward(s0, s1, weights, g: CellGeom, switches, P, key, steps=4, multiplier=4,
                 training=True, rec=None, p_skip=0.0, skip_mask=None):
    """CNN_Cell_search.forward, model_discCNN.py:96-201 (all edges present: DARTS/P-DARTS/CWP)."""
    if g.reduction_prev:
        s0 = op_factorized_reduce(s0, 2, P, key + ".preprocess0", training, rec)
    else:
        s0 = op_relu_conv_bn(s0, P, key + ".preprocess0", training, rec)
    s1 = op_relu_conv_bn(s1, P, key + ".preprocess1", training, rec)
    states = [s0, s1]
    e = 0
    for i in range(steps):
        node = 0
        for j in range(2 + i):
            stride = g.stride if (g.reduction and j < 2) else 1
            node = node + mixed_op(states[j], weights[e], switches[e], stride, P,
                                   f"{key}.cell_ops.{e}", training, rec, p_skip, skip_mask)
            e += 1
        states.append(node)
    return torch.cat(states[-multiplier:], dim=1)


# ------------------------------------------------------------------- RHN layer
def rnn_edge_table(switch_rnn):
    """For RHN step i: list of (j, k, prob_row, prob_col) for every enabled
    (edge, activation).  Restates comp_aux.py:127-233."""
    table = []
    r = 0
    disc = 0
    for i in range(RNN_STEPS):
        rows = []
        for j in range(i + 1):
            sw = switch_rnn[r]
            if not any(sw):
                disc += 1
            for k in range(1, 5):
                if sw[k]:
                    rows.append((j, k, r - disc, sum(1 for t in sw[:k] if t)))
            r += 1
        table.append(rows)
    return table


def _act(k, h):
    return (torch.tanh(h), h, torch.sigmoid(h), torch.relu(h))[k - 1]


def rhn_layer(x, h0, P, key, switch_rnn, weights, x_mask=None, h_mask=None, training=True, rec=None):
    """DARTSCell.forward + DARTSCellSearch.cell.
    model_discCNN.py:226-268, model_search.py:54-97.  x [T,B,ninp], h0 [B,nhid]."""
    nhid = P[key + "._Ws.0"].shape[0]
    W0 = P[key + "._W0"]
    Ws = [P[f"{key}._Ws.{i}"] for i in range(RNN_STEPS)]
    probs = F.softmax(weights, dim=-1)
    table = rnn_edge_table(switch_rnn)
    h = h0
    outs = []
    for t in range(x.shape[0]):
        xt = x[t]
        if training and x_mask is not None:
            a0 = torch.cat([xt * x_mask, h * h_mask], dim=-1)
        else:
            a0 = torch.cat([xt, h], dim=-1)
        ch = a0 @ W0
        c0, hh0 = ch[:, :nhid].sigmoid(), ch[:, nhid:].tanh()
        s = batchnorm(h + c0 * (hh0 - h), P, key + ".bn", training, rec)
        states = [s]
        for i in range(RNN_STEPS):
            acc = torch.zeros_like(s)
            cs, hs = {}, {}
            for (j, k, pr, pc) in table[i]:
                if j not in cs:
                    sj = states[j] * h_mask if (training and h_mask is not None) else states[j]
                    chj = sj @ Ws[i]
                    cs[j], hs[j] = chj[:, :nhid].sigmoid(), chj[:, nhid:]
                acc = acc + probs[pr, pc] * (states[j] + cs[j] * (_act(k, hs[j]) - states[j]))
            states.append(batchnorm(acc, P, key + ".bn", training, rec))
        h = torch.stack(states[1:], 0).mean(0)
        outs.append(h)
    return torch.stack(outs, 0), h


def draw_rhn_masks(B, nhid, dropoutx, dropouth, dtype=torch.float32):
    """utils.py:354-363 -- CPU generator, x mask first, then h mask (quirk Q10)."""
    xm = torch.floor(torch.rand(B, nhid) + (1.0 - dropoutx)) / (1.0 - dropoutx)
    hm = torch.floor(torch.rand(B, nhid) + (1.0 - dropouth)) / (1.0 - dropouth)
    return xm.to(dtype), hm.to(dtype)


# ------------------------------------------------------------------ full model
def softmax_alpha(alpha):
    """model_discCNN.py:497-507 (dim 0 when one column is left, quirk Q7)."""
    return F.softmax(alpha, dim=0 if alpha.shape[1] == 1 else -1)


def supernet_forward(P, cfg: SupernetConfig, x, x_mask=None, h_mask=None, training=True,
                     rec=None, lin_mask=None, skip_mask=None, h0=None):
    """RNNModel.forward, model_discCNN.py:491-570.  Returns (logits, hiddens[T,B,nhid])."""
    geoms, neurons, nhid = cell_geometry(cfg)
    z = F.conv1d(x, P["stem.0.weight"], padding=6)
    s0 = s1 = batchnorm(z, P, "stem.1", training, rec, P["stem.1.weight"], P["stem.1.bias"])
    for g in geoms:
        if g.reduction:
            w, sw = softmax_alpha(P["alphas_reduce"]), cfg.switches_reduce
        else:
            w, sw = softmax_alpha(P["alphas_normal"]), cfg.switches_normal
        s0, s1 = s1, cell_forward(s0, s1, w, g, sw, P, f"cells.{g.index}", cfg.steps, cfg.multiplier,
                                  training, rec, cfg.p_skip, skip_mask)
    seq = s1.permute(2, 0, 1)
    B = x.shape[0]
    if h0 is None:
        h0 = torch.zeros(B, nhid, dtype=x.dtype)
    hid, _ = rhn_layer(seq, h0, P, "rnns.0", cfg.switches_rnn, P["weights"], x_mask, h_mask, training, rec)
    flat = hid.permute(1, 0, 2).flatten(1)
    if lin_mask is not None and training:
        flat = flat * lin_mask
    y = F.relu(F.linear(flat, P["decoder.weight"], P["decoder.bias"]))
    y = F.linear(y, P["classifier.weight"], P["classifier.bias"])
    if cfg.task == "TF_bindings":
        y = torch.sigmoid(y)
    return y, hid


def bce(logits, target):
    """nn.BCELoss (mean), log clamped at -100 as torch does."""
    lp = torch.clamp(torch.log(logits), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - logits), min=-100.0)
    return -(target * lp + (1.0 - target) * l1p).mean()


def train_loss(logits, hid, target, beta):
    """train_and_validate.py:136-140: BCE + beta * TAR."""
    raw = bce(logits, target)
    return raw + beta * (hid[1:] - hid[:-1]).pow(2).mean(), raw


def param_names(P) -> List[str]:
    """Names in the reference's named_parameters() order are the insertion order
    of the dictionary; buffers and the alias rnns.0.weights are skipped."""
    skip = ("running_mean", "running_var", "num_batches_tracked")
    return [k for k in P if not k.endswith(skip) and k != "rnns.0.weights"]


def grads_of(loss, P, names):
    leaves = [P[n] for n in names]
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    return {n: (g if g is not None else torch.zeros_like(P[n])) for n, g in zip(names, gs)}


@dataclass
class Hyper:
    """README.md:72 + train_genomicDARTS.py:53-140 defaults."""
    cnn_lr: float = 0.025
    cnn_wd: float = 3e-4
    momentum: float = 0.9
    rhn_lr: float = 8.0
    rhn_wd: float = 5e-7
    clip: float = 0.25
    beta: float = 1e-3
    arch_lr: float = 3e-3
    arch_wd: float = 1e-3
    # P-DARTS / CWP-DARTS inline alpha step (train_and_validate.py:87-101, train_genomicPDARTS.py:300-301):
    # Adam(6e-4, betas (0.5, 0.999), wd 1e-3) after clip_grad_norm_(arch_parameters, clip); None = Architect
    adam_betas: Tuple[float, float] = (0.9, 0.999)
    arch_clip: Optional[float] = None


def pdarts_hyper() -> "Hyper":
    return Hyper(arch_lr=6e-4, adam_betas=(0.5, 0.999), arch_clip=0.25)


def adam_update(p, g, state, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used by Architect (architect.py:42)."""
    state["t"] = state.get("t", 0) + 1
    g = g + wd * p
    state["m"] = b1 * state.get("m", torch.zeros_like(p)) + (1 - b1) * g
    state["v"] = b2 * state.get("v", torch.zeros_like(p)) + (1 - b2) * g * g
    t = state["t"]
    step = lr / (1 - b1 ** t)
    denom = state["v"].sqrt() / math.sqrt(1 - b2 ** t) + eps
    return p - step * state["m"] / denom


def search_step(P, cfg, hyper: Hyper, x_train, y_train, x_valid, y_valid, opt_state,
                masks_valid=(None, None), masks_train=(None, None)):
    """One iteration of train() with train_arch=True, unrolled=False; pdarts=False (Architect) or, with
    hyper.arch_clip set, pdarts=True (inline clipped Adam, train_and_validate.py:87-101).
    train_and_validate.py:55-152 + architect.py:55-72.  P holds detached tensors;
    returns the updated dictionary and a dict of observables."""
    obs = {}
    arch = ("alphas_normal", "alphas_reduce", "weights")
    # ---- alpha step: fwd+bwd on the validation batch, Adam on the three alphas
    Q = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in P.items()}
    rec = Recorder()
    logits, hid = supernet_forward(Q, cfg, x_valid, masks_valid[0], masks_valid[1], True, rec)
    loss_a = bce(logits, y_valid)
    g = grads_of(loss_a, Q, list(arch))
    P = {k: v.detach() for k, v in Q.items()}
    rec.apply(P)
    if hyper.arch_clip is not None:      # nn.utils.clip_grad_norm_(model.arch_parameters(), clip)  (:97)
        tot_a = torch.sqrt(sum((g[n].double() ** 2).sum() for n in arch)).to(loss_a.dtype)
        coef_a = torch.clamp(hyper.arch_clip / (tot_a + 1e-6), max=1.0)
        g = {n: g[n] * coef_a for n in arch}
    for n in arch:
        P[n] = adam_update(P[n], g[n], opt_state.setdefault("adam_" + n, {}), hyper.arch_lr, hyper.arch_wd,
                           hyper.adam_betas[0], hyper.adam_betas[1])
    obs["loss_valid"] = loss_a.detach()
    obs["alpha_grads"] = {n: g[n].detach() for n in arch}
    # ---- weight step: fwd+bwd with TAR on the training batch, global clip, SGD
    Q = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in P.items()}
    rec = Recorder()
    logits, hid = supernet_forward(Q, cfg, x_train, masks_train[0], masks_train[1], True, rec)
    loss, raw = train_loss(logits, hid, y_train, hyper.beta)
    names = param_names(Q)
    g = grads_of(loss, Q, names)
    P = {k: v.detach() for k, v in Q.items()}
    rec.apply(P)
    total = torch.sqrt(sum((v.double() ** 2).sum() for v in g.values())).to(loss.dtype)
    coef = torch.clamp(hyper.clip / (total + 1e-6), max=1.0)
    for n in names:
        gn = g[n] * coef
        if n.startswith("rnns."):
            P[n] = P[n] - hyper.rhn_lr * (gn + hyper.rhn_wd * P[n])
        else:
            d = gn + hyper.cnn_wd * P[n]
            buf = opt_state.get("mom_" + n)
            buf = d.clone() if buf is None else hyper.momentum * buf + d
            opt_state["mom_" + n] = buf
            P[n] = P[n] - hyper.cnn_lr * buf
    P["rnns.0.weights"] = P["weights"]
    obs.update(loss_train=loss.detach(), raw_train=raw.detach(), logits=logits.detach(),
               grad_norm=total.detach(), grads={n: v.detach() for n, v in g.items()})
    return P, obs


# -------------------------------------------------------------- gradient sketches (golden fixtures)
def sketch_rows(numel: int) -> int:
    """random projections kept per gradient tensor in the full-size goldens"""
    return 128 if numel <= 262144 else 8


def sketch(g: torch.Tensor, index: int) -> torch.Tensor:
    """S random projections <g, r_k>, r_k ~ N(0, I) from a generator seeded by the tensor's position.  For an error
    e = g - g_ref, mean_k <e, r_k>^2 is an unbiased estimate of ||e||^2 (relative std sqrt(2/S): 12.5 % on ||e|| at
    S = 128), so every tensor of the 38 M-parameter model is graded by its actual error without storing it."""
    g = g.detach().double().flatten().cpu()
    S = sketch_rows(g.numel())
    gen = torch.Generator().manual_seed(7000 + index)
    out = torch.empty(S, dtype=torch.float64)
    for k in range(S):
        out[k] = torch.dot(torch.randn(g.numel(), generator=gen, dtype=torch.float64), g)
    return out


# -------------------------------------------------------------- synthetic data
def synthetic_batch(B, seq_len, num_classes, seed, dtype=torch.float32):
    """DeepSEA-shape synthetic batch."""
    gen = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, 4, (B, seq_len), generator=gen)
    x = torch.zeros(B, 4, seq_len).scatter_(1, idx[:, None, :], 1.0)
    y = (torch.rand(B, num_classes, generator=gen) < 0.05).float()
    return x.to(dtype), y.to(dtype)


def synthetic_params(shapes: Dict[str, Tuple[Tuple[int, ...], str]], seed: int):
    """Deterministic parameters for golden vectors that need no reference to
    rebuild: every float tensor is drawn from its own generator seeded by
    (seed, position in the dictionary).  shapes: name -> (shape, kind) with
    kind in {'w', 'alpha', 'rm', 'rv', 'nbt', 'gamma', 'beta', 'bias'}."""
    P = {}
    for i, (name, (shape, kind)) in enumerate(shapes.items()):
        gen = torch.Generator().manual_seed(seed * 100003 + i)
        if kind == "nbt":
            P[name] = torch.zeros((), dtype=torch.long)
        elif kind == "rm":
            P[name] = torch.zeros(shape)
        elif kind == "rv":
            P[name] = torch.ones(shape)
        elif kind == "gamma":
            P[name] = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif kind == "alpha":
            P[name] = 0.5 * torch.randn(shape, generator=gen)
        else:
            fan = 1
            for s in shape[1:]:
                fan *= s
            scale = 1.0 / math.sqrt(max(fan, 1)) if kind == "w" else 0.05
            P[name] = scale * torch.randn(shape, generator=gen)
    if "weights" in P:
        P["rnns.0.weights"] = P["weights"]
    return P
