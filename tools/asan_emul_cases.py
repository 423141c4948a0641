"""Shapes run by tools/asan_emul.sh (kernel sources on the CPU emulator under AddressSanitizer)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from genome_nas_b200 import _lib
_lib._use_library_for_tests(os.environ["GNAS_ASAN_LIB"], "cpu")
import genome_nas_b200 as G
import genomenas_oracle as O
torch.manual_seed(0)
dev = torch.device("cpu")
# MixedOp shapes incl. multi-segment rows and strides; a search cell; the head
for (B, C, L, S) in [(2, 8, 600, 1), (2, 8, 530, 2), (3, 8, 520, 3), (2, 16, 43, 2), (1, 8, 17, 1), (2, 32, 125, 3)]:
    m = G.MixedOp(C, S, [True] * 9, 0.0).to(dev).train()
    x = torch.randn(B, C, L, requires_grad=True); w = torch.softmax(torch.randn(9), 0).requires_grad_(True)
    y = m(x, w); y.sum().backward()
    print("mixedop", B, C, L, S, tuple(y.shape), flush=True)
sw = [[True] * 9 for _ in range(14)]
for (cpp, cp, c, red, redp, high, l0, l1) in [(12, 12, 4, False, False, False, 96, 96), (12, 16, 8, True, False, False, 96, 96), (16, 32, 8, True, True, True, 96, 48)]:
    m = G.CNN_Cell_search(4, 4, cpp, cp, c, red, redp, sw, 0.0, high).to(dev).train()
    s0 = torch.randn(3, cpp, l0, requires_grad=True); s1 = torch.randn(3, cp, l1, requires_grad=True)
    al = torch.randn(14, 9, requires_grad=True)
    y = m(s0, s1, torch.softmax(al, -1)); y.square().mean().backward()
    print("cell", c, red, tuple(y.shape), flush=True)
sr = [[True] * 5 for _ in range(36)]
r = G.DARTSCellSearch(64, 64, 0.1, 0.1, sr).to(dev).train()
r.weights = torch.nn.Parameter(1e-1 * torch.randn(36, 5))
x = torch.randn(3, 5, 64, requires_grad=True); h0 = torch.zeros(1, 5, 64)
y, _ = r(x, h0); y.square().mean().backward()
print("rhn ok", flush=True)
import genome_nas_b200.head as HD
class M(torch.nn.Module):
    _head_wgrad = True
    def __init__(s):
        super().__init__(); s.decoder = torch.nn.Linear(3 * 64, 70); s.classifier = torch.nn.Linear(70, 19); s.final = torch.nn.Sigmoid(); s.dropout_lin = torch.nn.Dropout(0.1)
mm = M().train()
from genome_nas_b200.flat import region_of
region_of(mm)
hs = torch.randn(3, 5, 64, requires_grad=True)
o = HD.head_forward(mm, hs); l = HD.bce_loss(o, (torch.rand(5, 19) > 0.5).float()) + HD.tar_loss(hs, 0.1); l.backward()
print("head ok", flush=True)
