"""Two eager search steps of the BASELINE workload (for ncu captures): python tools/prof_step.py [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import genome_nas_b200 as G  # noqa: E402
import genomenas_oracle as O  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(3)
np.random.seed(3)
a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s))) for s in ((14, 9), (14, 9), (36, 5))]
sw = lambda n, k: [[True] * k for _ in range(n)]
m = G.RNNModelSearch(1000, 0.05, 0.1, 8, 919, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sw(14, 9), sw(14, 9),
                     sw(36, 5), 0.0, *a).to(dev).train()
xt, yt = [t.to(dev) for t in O.synthetic_batch(64, 1000, 919, 0)]
xv, yv = [t.to(dev) for t in O.synthetic_batch(64, 1000, 919, 1)]
st = G.SearchStep(m)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = st.step(xt, yt, xv, yv)
torch.cuda.synchronize()
print("ok", float(out[1]))
if os.environ.get("GNAS_REPORT"):
    # event-bracketed per-kernel totals of one eager step (inflated for tiny kernels; fine for the big ones)
    from genome_nas_b200 import _lib
    _lib.profile(True)
    out = st.step(xt, yt, xv, yv)
    torch.cuda.synchronize()
    rep = _lib.profile_report()
    _lib.profile(False)
    for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:int(os.environ["GNAS_REPORT"])]:
        print(f"{ms:9.3f} ms {n:6d} x {1e3 * ms / n:9.1f} us  {k}")
