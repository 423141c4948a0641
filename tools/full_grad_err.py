"""Full-size golden (B=64, 4x1000, 919 labels): per-tensor gradient error vs the fp64 reference, next to the
reference's own fp32 error (ref32_err).  GNAS_TC=0/1 selects the conv_15 path."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import genomenas_oracle as O
import genome_nas_b200 as G
from genome_nas_b200 import rnn as R
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
g = torch.load(os.path.join(ROOT, "tests", "golden", "model_full.pt"), weights_only=False)
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
print("logits", f"{rel(logits, g['logits64']):.2e}", "loss", f"{abs(float(loss) - float(g['loss64'])) / abs(float(g['loss64'])):.2e}")
gn, r32 = g["grad_norm64"], g["ref32_err"]
worst = []
for i, (n, p) in enumerate(m.named_parameters()):
    if n in g["grad64"]:
        e = rel(p.grad, g["grad64"][n])
        worst.append((e / max(r32[n], 1e-12), e, r32[n], n))
worst.sort(reverse=True)
for ratio, e, r, n in worst[:12]:
    print(f"{n:60s} err {e:.2e}  ref32 {r:.2e}  ratio {ratio:.1f}")
for ratio, e, r, n in worst:
    if n in ("alphas_normal", "alphas_reduce", "weights"):
        print(f"{n:60s} err {e:.2e}  ref32 {r:.2e}  ratio {ratio:.1f}")
print("in grad64:", [n for n in ("alphas_normal", "alphas_reduce", "weights") if n in g["grad64"]], "params:", [n for n, _ in m.named_parameters()][:4])
print("median ratio", sorted(w[0] for w in worst)[len(worst) // 2], "n", len(worst))
