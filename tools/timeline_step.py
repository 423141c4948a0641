"""Timeline of one CUDA-graph replay of the search step (CUPTI via torch.profiler): wall span of the segments
(cells forward / RHN forward / head / RHN backward / cells backward, alpha pass and weight pass), time with no kernel
running (dependency gaps), and the average number of concurrently running kernels per segment."""
import os, re, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import genome_nas_b200 as G  # noqa: E402
import genomenas_oracle as O  # noqa: E402
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(3); np.random.seed(3)
a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s))) for s in ((14, 9), (14, 9), (36, 5))]
sw = lambda n, k: [[True] * k for _ in range(n)]
m = G.RNNModelSearch(1000, 0.05, 0.1, 8, 919, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sw(14, 9), sw(14, 9),
                     sw(36, 5), 0.0, *a).to(dev).train()
xt, yt = [t.to(dev) for t in O.synthetic_batch(64, 1000, 919, 0)]
xv, yv = [t.to(dev) for t in O.synthetic_batch(64, 1000, 919, 1)]
st = G.SearchStep(m)
st.capture(xt, yt, xv, yv, warmup=2)
for _ in range(3):
    st.step(xt, yt, xv, yv)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    st.step(xt, yt, xv, yv)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append((e.time_range.start, e.time_range.end, re.sub(r"\(.*", "", e.name).replace("gnas::", "").replace("void ", "")))
ev.sort()
t0 = ev[0][0]
ev = [(s - t0, e - t0, n) for s, e, n in ev]
span = max(e for _, e, _ in ev)
# segment boundaries: RHN forward kernels, head gather, RHN backward first/last kernel
fwd = [i for i, x in enumerate(ev) if "k_rhn_fwd_tc" in x[2] or x[2].startswith("k_rhn_forward")]
dout = [i for i, x in enumerate(ev) if "k_rhn_dout_T" in x[2]]
dprobs = [i for i, x in enumerate(ev) if "k_rhn_dprobs" in x[2]]
marks = []
for p in range(len(fwd)):
    marks += [("cells fwd", ev[fwd[p]][0]), ("rhn fwd", ev[fwd[p]][1]), ("head+loss", ev[dout[p]][0]), ("rhn bwd", ev[dprobs[p]][1])]
    marks.append(("cells bwd + optim", None))
bounds = [0.0] + [t for _, t in marks if t is not None] + [span]
names = []
for p in range(len(fwd)):
    names += [f"pass{p} cells fwd (+stem)", f"pass{p} rhn fwd", f"pass{p} head+loss", f"pass{p} rhn bwd (+wgrad)", f"pass{p} cells bwd + update"]
# reorder bounds: after 'rhn bwd' end comes cells bwd until the next pass's start (= next cells fwd start)
segs = []
cur = 0.0
for p in range(len(fwd)):
    segs.append((f"pass{p} cells fwd (+stem)", cur, ev[fwd[p]][0]))
    segs.append((f"pass{p} rhn fwd", ev[fwd[p]][0], ev[fwd[p]][1]))
    segs.append((f"pass{p} head + loss", ev[fwd[p]][1], ev[dout[p]][0]))
    segs.append((f"pass{p} rhn bwd (+wgrad)", ev[dout[p]][0], ev[dprobs[p]][1]))
    cur = ev[dprobs[p]][1]
    nxt = span
    if p + 1 < len(fwd):
        # the next pass starts with the stem forward: first kernel after the Adam update
        adam = [x for x in ev if "k_adam" in x[2]]
        nxt = adam[0][1] if adam else span
    segs.append((f"pass{p} cells bwd + update", cur, nxt))
    cur = nxt
def stats(lo, hi):
    pts = []
    for s, e, _ in ev:
        s2, e2 = max(s, lo), min(e, hi)
        if e2 > s2:
            pts += [(s2, 1), (e2, -1)]
    pts.sort()
    busy = idle = area = 0.0
    depth, last = 0, lo
    for t, d in pts:
        if depth > 0: busy += t - last
        else: idle += t - last
        area += depth * (t - last)
        depth += d; last = t
    idle += hi - last
    return busy, idle, area
print(f"span {span / 1e3:.2f} ms, {len(ev)} kernels")
for n, lo, hi in segs:
    b, i, ar = stats(lo, hi)
    k = sum(1 for s, e, _ in ev if lo <= s < hi)
    print(f"{n:32s} {(hi - lo) / 1e3:7.2f} ms  kernels {k:5d}  no-kernel time {i / 1e3:6.2f} ms  avg concurrency {ar / max(b, 1e-9):4.2f}  sum of kernel times {ar / 1e3:7.2f} ms")
