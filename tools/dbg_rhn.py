import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import genome_nas_b200 as G
dev = torch.device("cuda:0")
H, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sw = [[True] * 5 for _ in range(36)]
m = G.DARTSCellSearch(H, H, 0.05, 0.1, sw).to(dev).train()
m.weights = torch.nn.Parameter(1e-3 * torch.randn(36, 5, device=dev))
x = torch.randn(T, B, H, device=dev, requires_grad=True)
h0 = torch.zeros(1, B, H, device=dev)
y, _ = m(x, h0)
torch.cuda.synchronize(); print("fwd ok", float(y.abs().mean()))
y.square().mean().backward()
torch.cuda.synchronize(); print("bwd ok", float(x.grad.abs().mean()))
