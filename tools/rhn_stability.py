import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import torch
from genome_nas_b200.rnn import DARTSCellSearch
dev = torch.device("cuda:0")
m = DARTSCellSearch(512, 512, 0.25, 0.25, [[True]*5 for _ in range(36)])
m.weights = torch.nn.Parameter(1e-1*torch.randn(36, 5)); m.to(dev).train()
x = torch.randn(42, 64, 512, device=dev); h = torch.zeros(1, 64, 512, device=dev)
ref = None
for it in range(40):
    torch.manual_seed(1)
    with torch.no_grad():
        y, _ = m(x, h)
    torch.cuda.synchronize()
    s = float(y.double().sum())
    if ref is None: ref = s
    assert s == ref, (it, s, ref)
print("40 forward passes bit-identical", ref)
