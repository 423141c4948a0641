"""Debug: clock64 marks of CTA 0 inside the persistent RHN forward kernels (library built with -DGNAS_RHN_TIMING).
GNAS_RHN_TC=1 (default): tensor-core kernel, marks of time step 1 per stage; GNAS_RHN_TC=0: fp32 kernel."""
import ctypes as C, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = os.path.join(ROOT, "genome-nas_b200", "csrc")
out = os.path.join(ROOT, "tools", "libgnas_timing.so")
if not os.path.exists(out) or "--rebuild" in sys.argv:
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "--expt-relaxed-constexpr", "-DGNAS_RHN_TIMING", "-shared", "-o", out] +
                          [os.path.join(src, f) for f in sorted(os.listdir(src)) if f.endswith(".cu")])
if "--build-only" in sys.argv:
    sys.exit(0)
import genome_nas_b200 as G
from genome_nas_b200 import _lib
import genome_nas_b200.rnn as R
_lib._use_library_for_tests(out, "cuda")
dev = torch.device("cuda:0")
T, B, H = 42, 64, 512
sw = [[True] * 5 for _ in range(36)]
m = G.DARTSCellSearch(H, H, 0.05, 0.1, sw).to(dev).train()
m.weights = torch.nn.Parameter(1e-3 * torch.randn(36, 5, device=dev))
x = torch.randn(T, B, H, device=dev)
h0 = torch.zeros(1, B, H, device=dev)
region = R.region_of(m, exclude=('weights',))
lib = _lib.lib()
probs = torch.softmax(m.weights, -1).detach().contiguous()
desc = m._descriptor(T, B, H, probs)
fwd, bwd = C.c_int64(0), C.c_int64(0)
lib.gnas_rhn_workspace(C.byref(desc), C.byref(fwd), C.byref(bwd))
ws = torch.empty(fwd.value, dtype=torch.float32, device=dev)
out_t = torch.empty(T, B, H, device=dev)
xm = torch.ones(B, H, device=dev); hm = torch.ones(B, H, device=dev)
flat, off = region.flat, region.offsets
for _ in range(3):
    rc = lib.gnas_rhn_forward(C.byref(desc), _lib.ptr(x), _lib.ptr(h0[0].contiguous()), _lib.ptr(flat[off["_W0"]:]),
                              _lib.ptr(flat[off["_Ws.0"]:]), _lib.ptr(probs), _lib.ptr(xm), _lib.ptr(hm), None,
                              _lib.ptr(out_t), _lib.ptr(ws), _lib.stream_ptr(x))
    assert rc == 0, rc
torch.cuda.synchronize()
BP = (B + 3) // 4 * 4
HB = H * BP
bar_off = T * HB + 3 * HB + 10 * HB + T * 9 * HB + T * 36 * 2 * HB + T * 2 * HB + T * HB + T * 9 * 2 * H
if os.environ.get("GNAS_RHN_TC", "1") != "0":
    marks = ws[bar_off + 16: bar_off + 16 + 9 * 8 * 2].view(torch.int64).cpu().view(9, 8)
    base = int(marks[0, 4]) if int(marks[0, 4]) else int(marks[0, 0])
    print("tensor-core kernel, CTA 0, time step 1 [cycles relative to the stage-0 barrier release]")
    print("stage:  barrier-passed  mma-first-chunk  first-chunk-done  drained+dumped  published  stage-end  mma-last-issued")
    prev_pub = None
    for s in range(9):
        v = [int(marks[s, k]) - base for k in range(8)]
        line = f"{s}: {v[4]:9d} {v[5]:9d} {v[0]:9d} {v[1]:9d} {v[2]:9d} {v[3]:9d} {v[6]:9d}"
        if prev_pub is not None:
            line += (f"   publish->barrier {v[4] - prev_pub:6d}  barrier->dumped {v[1] - v[4]:6d}  dumped->bn+stores {v[7] - v[1]:6d}"
                     f"  ->published {v[2] - v[7]:6d}")
        prev_pub = v[2]
        print(line)
    print("time step cycles (published stage 8 - barrier stage 0):", int(marks[8, 2]) - base)
    ch = ws[bar_off + 16 + 72 * 2: bar_off + 16 + (72 + 40) * 2].view(torch.int64).cpu().view(8, 5)
    b1 = int(marks[1, 4])
    print("stage 1 chunks [cycles after the stage's barrier release]: mma: acc-empty ok | operands ok | issued ;  epilogue warp 0: acc-full seen | drained")
    for c in range(8):
        v = [int(ch[c, k]) - b1 for k in range(5)]
        print(f"  chunk {c}: {v[0]:7d} {v[1]:7d} {v[2]:7d}   {v[3]:7d} {v[4]:7d}")
else:
    marks = ws[bar_off + 16: bar_off + 16 + 9 * 4 * 2].view(torch.int64).cpu().view(9, 4)
    base = int(marks[0, 0])
    print("stage: gemm | crit-elem+bn | arrive..wait(end)   [cycles]   total")
    for s in range(9):
        a, b, c, d = [int(v) for v in marks[s]]
        print(f"{s}: start {a - base:8d}  gemm {b - a:7d}  crit {c - b:6d}  barrier {d - c:7d}")
    print("timestep cycles:", int(marks[8, 3]) - base)
