"""RHN layer alone at the benchmark shape (T=42, B=64, H=512): per-kernel device time (CUDA events around every
launch) and parity against the fp64 oracle.  GNAS_RHN_TC=0/1 selects the fp32 CUDA-core / tensor-core forward."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import genomenas_oracle as O
import genome_nas_b200.rnn as R
from genome_nas_b200 import _lib
from genome_nas_b200.rnn import DARTSCellSearch
T, B, H = int(os.environ.get("T", 42)), int(os.environ.get("B", 64)), int(os.environ.get("H", 512))
dev = torch.device("cuda:0")
torch.manual_seed(5)
switch = [[True] * 5 for _ in range(36)]
m = DARTSCellSearch(H, H, 0.25, 0.25, switch)
m.weights = torch.nn.Parameter(1e-1 * torch.randn(36, 5))
m.to(dev).train()
xm = (torch.rand(B, H) < 0.75).float() / 0.75
hm = (torch.rand(B, H) < 0.75).float() / 0.75
_calls = [0]


def _mask(B_, D, keep, d):      # x mask first, then h mask (utils.py:354-363)
    _calls[0] += 1
    return (xm if _calls[0] % 2 == 1 else hm).to(d)


R.mask2d = _mask
x0, h00, gy = torch.randn(T, B, H), 0.1 * torch.randn(B, H), torch.randn(T, B, H)
def run():
    x = x0.to(dev).requires_grad_(True); h0 = h00.to(dev).requires_grad_(True)
    y, _ = m(x, h0[None])
    (y * gy.to(dev)).sum().backward()
    return x, h0, y
for _ in range(2):
    m.zero_grad(); run()
torch.cuda.synchronize()
# wall time of a forward + backward pass (CUDA events, no per-launch bracketing): what the side-stream overlap changes
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.zero_grad(); run()
e1.record(); torch.cuda.synchronize()
print(f"SIDE={os.environ.get('GNAS_RHN_SIDE', '1')} fwd+bwd wall {e0.elapsed_time(e1) / 5:.3f} ms/pass")
_lib.profile(True)
n = 3
for _ in range(n):
    m.zero_grad(); x, h0, y = run()
rep = _lib.profile_report(); _lib.profile(False)
print(f"TC={os.environ.get('GNAS_RHN_TC', '1')} T={T} B={B} H={H}")
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms / n:9.3f} ms/pass  {c // n:5d} launches  {k[:70]}")
if os.environ.get("CHECK", "1") == "1":
    P = {("rnns.0." + k): v.detach().double().cpu().requires_grad_(True) for k, v in m.state_dict().items() if v.is_floating_point() and "running" not in k}
    P.update({"rnns.0.bn.running_mean": torch.zeros(H, dtype=torch.float64), "rnns.0.bn.running_var": torch.ones(H, dtype=torch.float64)})
    al = m.weights.detach().double().cpu().requires_grad_(True)
    xo, ho = x0.double().requires_grad_(True), h00.double().requires_grad_(True)
    yo, _ = O.rhn_layer(xo, ho, P, "rnns.0", switch, al, xm.double(), hm.double(), True, O.Recorder())
    (yo * gy.double()).sum().backward()
    rel = lambda a, b: float((a.detach().double().cpu() - b).norm() / (b.norm() + 1e-30))
    print(f"  vs fp64 oracle: y {rel(y, yo):.2e} dx {rel(x.grad, xo.grad):.2e} dh0 {rel(h0.grad, ho.grad):.2e} dW0 {rel(m._W0.grad, P['rnns.0._W0'].grad):.2e} "
          f"dWs7 {rel(m._Ws[7].grad, P['rnns.0._Ws.7'].grad):.2e} dalpha {rel(m.weights.grad, al.grad):.2e}")
