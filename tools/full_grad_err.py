"""Full-size goldens (B=64, 4x1000, 919 labels): per-tensor gradient error vs the fp64 reference (exact where the
golden stores the tensor, else estimated from the 128-projection sketch), graded like tests/test_model.py:
score = err / max(1e-3, 1.5 * max over the reference's two CPU conv backends of its own fp32 error).  Usage: python tools/full_grad_err.py [full|pdarts3]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import genomenas_oracle as O
import genome_nas_b200 as G
from genome_nas_b200 import rnn as R
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
tag = sys.argv[1] if len(sys.argv) > 1 else "full"
g = torch.load(os.path.join(ROOT, "tests", "golden", f"model_{tag}.pt"), weights_only=False)
cfg = g["cfg"]
P = O.synthetic_params(g["shapes"], g["seed_params"])
mk = lambda k: torch.nn.Parameter(P[k].clone())
m = G.RNNModelSearch(cfg.seq_len, cfg.dropouth, cfg.dropoutx, cfg.C, cfg.num_classes, cfg.layers, cfg.steps,
                     cfg.multiplier, cfg.stem_multiplier, True, 0.2, None, cfg.task, cfg.switches_normal,
                     cfg.switches_reduce, cfg.switches_rnn, cfg.p_skip, mk("alphas_normal"), mk("alphas_reduce"), mk("weights"))
m.load_state_dict(P); m.dropout_lin.p = 0.0
dev = torch.device("cuda:0")
m = m.to(dev).train()
masks = [g["xm"], g["hm"]]
R.mask2d = lambda B, D, keep, d: masks.pop(0).to(d)
xt, yt = O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0)
xt, yt = xt.to(dev), yt.to(dev)
logits, hidden, rnn_hs, _ = m(xt, m.init_hidden(g["B"]), return_h=True)
raw = F.binary_cross_entropy(logits, yt)
loss = raw + O.Hyper().beta * (rnn_hs[-1][1:] - rnn_hs[-1][:-1]).pow(2).mean()
loss.backward()
rel = lambda a, b: float((a.detach().double().cpu() - b.double()).norm() / (b.double().norm() + 1e-30))
print(f"[{tag}] TC={os.environ.get('GNAS_TC', '1')} RHN_TC={os.environ.get('GNAS_RHN_TC', '1')}  logits {rel(logits, g['logits64']):.2e}  "
      f"loss {abs(float(loss.detach()) - float(g['loss64'])) / abs(float(g['loss64'])):.2e}  hid {abs(float(rnn_hs[-1].double().norm()) - g['hid64_norm']) / g['hid64_norm']:.2e}")
gn, r32 = g["grad_norm64"], g["ref32_err"]
total = math.sqrt(sum(v * v for v in gn.values()))
rows = []
for i, (n, p) in enumerate(m.named_parameters()):
    if gn[n] < 1e-12 * total:
        continue
    if n in g["grad64"]:
        e, how = rel(p.grad, g["grad64"][n]), "exact"
    else:
        d = O.sketch(p.grad, i) - g["grad_sketch64"][n]
        e, how = float(d.pow(2).mean().sqrt()) / gn[n], "sketch"
    rb = g.get("ref32b_err")
    spread = max(rb["onednn"][n], rb["native"][n]) if rb else r32[n]
    lim = max(1e-3, 1.5 * spread)          # the bar of tests/test_model.py
    rows.append((e / lim, e, r32[n], how + f" native {rb['native'][n]:.2e}" if rb else how, n))
rows.sort(reverse=True)
for sc, e, r, how, n in rows[:int(os.environ.get("TOP", "14"))]:
    print(f"  score {sc:5.2f}  err {e:.2e}  ref32 {r:.2e}  {how:6s} {n}")
ratios = sorted(e / max(r, 1e-12) for _, e, r, _, _ in rows)
print(f"  tensors {len(rows)}  over the bar: {sum(1 for r in rows if r[0] >= 1.0)}  median err/ref32 {ratios[len(ratios) // 2]:.2f}  "
      f"p90 {ratios[int(0.9 * len(ratios))]:.2f}  max {ratios[-1]:.2f}")
for n in ("alphas_normal", "alphas_reduce", "weights", "rnns.0._W0", "decoder.weight"):
    for sc, e, r, how, nn in rows:
        if nn == n:
            print(f"  {n:16s} err {e:.2e} ref32 {r:.2e} score {sc:.2f}")
