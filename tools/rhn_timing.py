"""Debug: per-phase clock64 marks of CTA 0 in the persistent RHN forward kernel (built with -DGNAS_RHN_TIMING)."""
import ctypes, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = os.path.join(ROOT, "genome-nas_b200", "csrc")
out = "/tmp/libgnas_timing.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                       "--expt-relaxed-constexpr", "-DGNAS_RHN_TIMING", "-shared", "-o", out] +
                      [os.path.join(src, f) for f in sorted(os.listdir(src)) if f.endswith(".cu")])
import genome_nas_b200 as G
from genome_nas_b200 import _lib
import genome_nas_b200.rnn as R
_lib._use_library_for_tests(out, "cuda")
dev = torch.device("cuda:0")
sw = [[True] * 5 for _ in range(36)]
m = G.DARTSCellSearch(512, 512, 0.05, 0.1, sw).to(dev).train()
m.weights = torch.nn.Parameter(1e-3 * torch.randn(36, 5, device=dev))
x = torch.randn(42, 64, 512, device=dev)
h0 = torch.zeros(1, 64, 512, device=dev)
keep = {}
orig = R._RhnFunction.forward
with torch.no_grad():
    for _ in range(2):
        region = R.region_of(m, exclude=('weights',))
        y, _ = m(x, h0)
torch.cuda.synchronize()
# re-run once more capturing the workspace: call the C ABI directly
import ctypes as C
lib = _lib.lib()
desc = m._descriptor(42, 64, 512, torch.softmax(m.weights, -1))
fwd, bwd = C.c_int64(0), C.c_int64(0)
lib.gnas_rhn_workspace(C.byref(desc), C.byref(fwd), C.byref(bwd))
ws = torch.empty(fwd.value, dtype=torch.float32, device=dev)
out_t = torch.empty(42, 64, 512, device=dev)
probs = torch.softmax(m.weights, -1).detach().contiguous()
xm = torch.ones(64, 512, device=dev); hm = torch.ones(64, 512, device=dev)
flat, off = region.flat, region.offsets
rc = lib.gnas_rhn_forward(C.byref(desc), _lib.ptr(x), _lib.ptr(h0[0].contiguous()), _lib.ptr(flat[off["_W0"]:]),
                          _lib.ptr(flat[off["_Ws.0"]:]), _lib.ptr(probs), _lib.ptr(xm), _lib.ptr(hm), None,
                          _lib.ptr(out_t), _lib.ptr(ws), _lib.stream_ptr(x))
torch.cuda.synchronize()
T, H, BP = 42, 512, 64
HB = H * BP
bar_off = T * HB + 3 * HB + 10 * HB + T * 9 * HB + T * 36 * 2 * HB + T * 2 * HB + T * HB + T * 9 * 2 * H
marks = ws[bar_off + 16: bar_off + 16 + 9 * 4 * 2].view(torch.int64).cpu().view(9, 4)
base = int(marks[0, 0])
print("stage: gemm | crit-elem+bn | arrive..wait(end)   [cycles]   total")
for s in range(9):
    a, b, c, d = [int(v) for v in marks[s]]
    print(f"{s}: start {a - base:8d}  gemm {b - a:7d}  crit {c - b:6d}  barrier {d - c:7d}")
print("timestep cycles:", int(marks[8, 3]) - base)
