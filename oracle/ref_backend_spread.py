"""How far do two fp32 runs of the UNMODIFIED reference differ from the fp64 truth -- per gradient tensor --
when only the CPU convolution backend changes (oneDNN on / off: torch.backends.mkldnn)?

The whole-model gradients are ill-conditioned: individual ReLU / max-pool decisions sit within
fp32 rounding of their thresholds, so two correct fp32 implementations land on different sides and whole groups
of weight gradients move by 1e-3..3e-3.  This script measures that spread on the full-size goldens with the
reference's own code and adds it to the fixture as `ref32b_err` (authoring container only):

    python oracle/ref_backend_spread.py full pdarts3
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG          # noqa: E402  (imports the reference from /root/reference)
import genomenas_oracle as O      # noqa: E402


def run(tag):
    path = os.path.join(MG.GOLD, f"model_{tag}.pt")
    g = torch.load(path, weights_only=False)
    cfg, B = g["cfg"], g["B"]
    P = O.synthetic_params(g["shapes"], g["seed_params"])
    xt, yt = O.synthetic_batch(B, cfg.seq_len, cfg.num_classes, 0)
    hyper = O.Hyper()
    errs = {}
    for backend, enabled in (("onednn", True), ("native", False)):
        with torch.backends.mkldnn.flags(enabled=enabled):
            mm = MG.build_ref_model(cfg, P)
            mm.dropout_lin.p = 0.0
            mm.train()
            torch.manual_seed(77)      # the seed the golden's masks were drawn with (make_golden.py:golden_model)
            logits, hidden, rnn_hs, _ = mm(xt, mm.init_hidden(B), return_h=True)
            raw = torch.nn.functional.binary_cross_entropy(logits, yt)
            loss = raw + hyper.beta * (rnn_hs[-1][1:] - rnn_hs[-1][:-1]).pow(2).mean()
            loss.backward()
        e = {}
        for i, (n, p) in enumerate(mm.named_parameters()):
            ref = g["grad_norm64"][n]
            if n in g["grad64"]:
                e[n] = MG.relerr(p.grad, g["grad64"][n])
            else:
                d = O.sketch(p.grad, i) - g["grad_sketch64"][n]
                e[n] = float(d.pow(2).mean().sqrt()) / max(ref, 1e-300)
        errs[backend] = e
        print(f"[{tag}] reference fp32, conv backend {backend}: loss {float(loss):.8f}")
    a, b = errs["onednn"], errs["native"]
    stored = g["ref32_err"]
    names = [n for n in a if g["grad_norm64"][n] > 0]
    # which backend produced the stored ref32_err?  (it should reproduce one of them up to the sketch noise)
    da = sorted(abs(a[n] - stored[n]) / max(stored[n], 1e-12) for n in names)
    db = sorted(abs(b[n] - stored[n]) / max(stored[n], 1e-12) for n in names)
    print(f"  stored ref32_err vs onednn run: median rel diff {da[len(da) // 2]:.3f}; vs native run: {db[len(db) // 2]:.3f}")
    ratio = sorted(max(a[n], b[n]) / max(min(a[n], b[n]), 1e-12) for n in names)
    over = sum(1 for n in names if max(a[n], b[n]) > max(1e-3, 1.5 * min(a[n], b[n])))
    print(f"  per-tensor error ratio between the two backends: median {ratio[len(ratio) // 2]:.2f}  p90 {ratio[int(.9 * len(ratio))]:.2f}  "
          f"max {ratio[-1]:.2f};  tensors where one backend fails the 1.5x bar of the other: {over} of {len(names)}")
    g["ref32b_err"] = {"onednn": a, "native": b}
    torch.save(g, path)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for tag in sys.argv[1:] or ["full"]:
        run(tag)
