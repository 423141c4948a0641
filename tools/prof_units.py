"""Stand-alone units of the step for ncu captures (small memory footprint => fast kernel replay).
    python tools/prof_units.py rhn        # RHN layer T=42, B=64, H=512 fwd+bwd
    python tools/prof_units.py cell 5     # search cell 5 (C=128, L 125->42, stride 3) fwd+bwd
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genome_nas_b200 as G  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1]
reps = 2
if what == "rhn":
    sw = [[True] * 5 for _ in range(36)]
    m = G.DARTSCellSearch(512, 512, 0.05, 0.1, sw).to(dev).train()
    m.weights = torch.nn.Parameter(1e-3 * torch.randn(36, 5, device=dev))
    x = torch.randn(42, 64, 512, device=dev, requires_grad=True)
    h0 = torch.zeros(1, 64, 512, device=dev)
    for _ in range(reps):
        y, _ = m(x, h0)
        y.square().mean().backward()
else:
    idx = int(sys.argv[2])
    geo = [(24, 24, 8, False, False, False, 1000, 1000), (24, 32, 16, True, False, False, 1000, 1000),
           (32, 64, 32, True, True, False, 1000, 500), (64, 128, 32, False, True, False, 500, 250),
           (128, 128, 64, True, False, False, 250, 250), (128, 256, 128, True, True, True, 250, 125)][idx]
    cpp, cp, c, red, redp, high, l0, l1 = geo
    sw = [[True] * 9 for _ in range(14)]
    m = G.CNN_Cell_search(4, 4, cpp, cp, c, red, redp, sw, 0.0, high).to(dev).train()
    s0 = torch.randn(64, cpp, l0, device=dev, requires_grad=True)
    s1 = torch.randn(64, cp, l1, device=dev, requires_grad=True)
    al = torch.randn(14, 9, device=dev, requires_grad=True)
    for _ in range(reps):
        y = m(s0, s1, torch.softmax(al, -1))
        y.square().mean().backward()
torch.cuda.synchronize()
print("ok")
