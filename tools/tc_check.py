"""conv_15 MixedOp on the GPU vs the fp64 oracle: prints relative errors per tensor (debug aid for the tcgen05 path)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import genomenas_oracle as O
import genome_nas_b200 as G


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def shapes(module, prefix):
    kinds = {"running_mean": "rm", "running_var": "rv", "num_batches_tracked": "nbt"}
    return {prefix + k: (tuple(v.shape), next((t for s, t in kinds.items() if k.endswith(s)), "w"))
            for k, v in module.state_dict().items()}


cases = [(2, 32, 125, 1), (3, 32, 250, 1), (2, 64, 125, 1), (5, 64, 42, 1), (2, 128, 200, 1), (1, 128, 200, 1), (3, 128, 42, 1), (4, 128, 42, 1), (7, 32, 20, 1), (4, 64, 300, 1)]
ALL = [True] * 9
PW = [False, False, False, False, True, True, True, True, False]
cases = [(c + ([False] * 8 + [True],)) for c in cases] + [(2, 32, 250, 1, PW), (2, 64, 125, 1, PW), (3, 128, 42, 1, PW), (5, 128, 42, 3, ALL),
                                                        (2, 64, 125, 2, ALL), (3, 32, 250, 1, ALL), (2, 128, 130, 1, PW),
                                                        (3, 8, 1000, 1, ALL), (2, 16, 500, 1, ALL), (2, 8, 77, 1, ALL), (3, 16, 40, 2, ALL), (2, 16, 333, 1, [False] * 8 + [True]), (3, 16, 500, 1, PW), (5, 16, 50, 1, ALL),
                                                        (2, 32, 250, 2, ALL), (3, 128, 125, 3, ALL), (2, 64, 77, 2, [False] * 8 + [True]), (5, 128, 41, 3, [False] * 8 + [True]),
                                                        (1, 32, 1, 1, ALL), (70, 32, 3, 1, ALL), (2, 64, 1000, 1, ALL), (3, 128, 2, 3, ALL), (2, 64, 5, 2, ALL), (1, 16, 129, 1, ALL)]
for B, C, L, S, sw in cases:
    torch.manual_seed(1)
    m = G.MixedOp(C, S, sw, 0.0)
    P = O.synthetic_params(shapes(m, "cell_ops.0."), 5)
    m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
    m.to("cuda").train()
    x0, w0 = torch.randn(B, C, L), torch.softmax(torch.randn(sum(sw)), 0)
    x = x0.clone().cuda().requires_grad_(True); w = w0.clone().cuda().requires_grad_(True)
    y = m(x, w)
    gy = torch.randn(y.shape)
    (y * gy.cuda()).sum().backward()
    Q = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xo, wo = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    yo = O.mixed_op(xo, wo, sw, S, Q, "cell_ops.0", True, O.Recorder(), 0.0, None)
    (yo * gy.double()).sum().backward()
    line = [f"B{B} C{C} L{L} S{S} ops{sum(sw)}: y {relerr(y, yo):.2e} dx {relerr(x.grad, xo.grad):.2e} dw {relerr(w.grad, wo.grad):.2e}"]
    errs = {n: relerr(p.grad, Q['cell_ops.0.' + n].grad) for n, p in m.named_parameters()}
    worst = max(errs, key=errs.get)
    line.append(f"worst dW {errs[worst]:.2e} ({worst})")
    print(" | ".join(line), flush=True)
    if os.environ.get("TC_KNIFE"):
        import torch.nn.functional as F
        W1 = Q["cell_ops.0.m_ops.conv_15.op.1.weight"].detach()
        z1 = F.conv1d(torch.relu(x0.double()), W1, padding=7)
        mu, sd = z1.mean(dim=(0, 2), keepdim=True), z1.std(dim=(0, 2), keepdim=True)
        r = ((z1 - mu) / sd).abs()
        k = torch.topk(-r.flatten(), 3)
        print("   closest |z1-mean|/sd:", [(f"{-v:.2e}", tuple(int(t) for t in torch.unravel_index(i, r.shape))) for v, i in zip(k.values, k.indices)])
        d = (x.grad.cpu().double() - xo.grad).abs()
        i = d.flatten().argmax()
        print("   max dx err at (b, ch, l):", tuple(int(t) for t in torch.unravel_index(i, d.shape)), f"{float(d.flatten()[i]):.2e}")
    if os.environ.get("TC_VERBOSE"):
        d = (x.grad.cpu().double() - xo.grad).abs()
        print("   dx abs err by channel (max):", [f"{v:.1e}" for v in d.amax(dim=(0, 2))[:16].tolist()])
        bad = (d.amax(dim=1) > 20 * d.median()).nonzero().tolist()
        rows = {}
        for b_, l_ in bad:
            rows.setdefault(b_, []).append(l_)
        print("   dx bad positions per batch row:", {k: (min(v), max(v), len(v)) for k, v in rows.items()})
        for n, p in m.named_parameters():
            dd = (p.grad.cpu().double() - Q['cell_ops.0.' + n].grad).abs()
            if dd.dim() == 3:
                print("   ", n, "err by tap:", [f"{v:.1e}" for v in dd.amax(dim=(0, 1)).tolist()])
