import sys, time, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import genome_nas_b200 as G
import genomenas_oracle as O
from genome_nas_b200 import _lib
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda:0')
torch.manual_seed(3); np.random.seed(3)
a = [torch.nn.Parameter(torch.FloatTensor(1e-3*np.random.randn(*s))) for s in ((14,9),(14,9),(36,5))]
sw = lambda n,k: [[True]*k for _ in range(n)]
m = G.RNNModelSearch(1000, 0.05, 0.1, 8, 919, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sw(14,9), sw(14,9), sw(36,5), 0.0, *a).to(dev).train()
B=64
xt, yt = O.synthetic_batch(B, 1000, 919, 0); xv, yv = O.synthetic_batch(B, 1000, 919, 1)
xt,yt,xv,yv = [t.to(dev) for t in (xt,yt,xv,yv)]
st = G.SearchStep(m)
def sync(): torch.cuda.synchronize()
for i in range(3):
    t=time.time(); out = st.step(xt,yt,xv,yv); sync(); print('warm', i, time.time()-t, [float(o) for o in out[:3]])
# phase timing
def timeit(f, n=3):
    sync(); t=time.time()
    for _ in range(n): f()
    sync(); return (time.time()-t)/n*1e3
print('step ms', timeit(lambda: st.step(xt,yt,xv,yv), 5))
print('alpha ms', timeit(lambda: st.alpha_step(xv,yv)))
print('weight ms', timeit(lambda: st.weight_step(xt,yt)))
def fwd():
    with torch.no_grad(): m(xt, m.init_hidden(B))
print('fwd-only ms', timeit(fwd))
# sub-module timing
with torch.no_grad():
    s = m.stem(xt) if False else None
import torch.autograd.profiler as prof
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as p:
    st.step(xt,yt,xv,yv); sync()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
