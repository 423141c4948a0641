"""CUPTI (torch.profiler) per-kernel device time of one search step, eager and CUDA-graph replay."""
import os
import sys
import collections
import re

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
st.capture(xt, yt, xv, yv, warmup=2)
for _ in range(2):
    st.step(xt, yt, xv, yv)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    st.step(xt, yt, xv, yv)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)
        agg[name][0] += 1
        agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        s, e = ev.time_range.start, ev.time_range.end
        t0 = s if t0 is None else min(t0, s)
        t1 = e if t1 is None else max(t1, e)
busy = sum(v[1] for v in agg.values())
print(f"span_us {t1 - t0:.0f} busy_us {busy:.0f} kernels {sum(v[0] for v in agg.values())}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1] / 1e3:9.3f} ms {v[0]:6d} x {v[1] / v[0]:8.1f} us  {k[:90]}")
