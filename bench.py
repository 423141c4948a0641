#!/usr/bin/env python
"""Benchmark of the genomeDARTS supernet search step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload darts|pdarts3]

One "step" = one search iteration: first-order alpha step on a validation batch (fwd + bwd + Adam on
the alphas) + weight step on a training batch (fwd + bwd with TAR, global-norm clip, SGD), i.e. the body
of generalNAS_tools/train_and_validate.py:55-152.  Workload `darts` = BASELINE configs[1]: full supernet
(C=8, 6 layers, 9 CNN / 5 RNN candidates, 38.0 M parameters), batch 64 per GPU, one-hot 4x1000 input,
919 labels, fp32.  Workload `pdarts3` = BASELINE configs[3]: the P-DARTS / CWP-DARTS stage-3 supernet
(3 CNN / 2 RNN candidates per edge, switches produced by the reference's own discard_* functions and stored
in tests/golden/pdarts3_switches.json, skip-connect dropout 0.1, inline Adam(6e-4, (0.5, 0.999)) alpha step
with its own clip: train_and_validate.py:87-101).  N > 1 (torchrun): batch-sharded data parallel, NCCL
allreduce of the flat gradient buffer, weak scaling (64 sequences per GPU).  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's own CPU implementation (oracle/_ref: the unmodified reference,
byte-compiled by oracle/build_ref.py; its own RNNModelSearch + Architect + train()) on the host cores at
the SAME batch of 64 sequences per step; when oracle/_ref is missing it falls back to the oracle port.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "darts": "genomeDARTS supernet search step, batch 64/GPU, seq 1000, 919 labels (BASELINE configs[1])",
    "pdarts3": "genomeP-DARTS/CWP-DARTS stage-3 pruned supernet search step (3 CNN / 2 RNN candidates per edge, "
               "skip dropout 0.1), batch 64/GPU, seq 1000, 919 labels (BASELINE configs[3])",
}
METRIC = "supernet train seqs/sec (fwd+bwd+alpha) @1000bp"
SEQ, NCLS, BATCH = 1000, 919, 64
REF_BUDGET_S = float(os.environ.get("GNAS_REF_BUDGET_S", "1500"))     # the driver allows 1800 s per arm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="darts", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    return ap.parse_args()


def switches(workload):
    if workload == "pdarts3":
        sw = json.load(open(os.path.join(ROOT, "tests", "golden", "pdarts3_switches.json")))
        return sw["normal"], sw["reduce"], sw["rnn"]
    return [[True] * 9 for _ in range(14)], [[True] * 9 for _ in range(14)], [[True] * 5 for _ in range(36)]


def hyper_for(workload):
    """README.md:72 / train_genomicDARTS.py:53-140; the P-DARTS alpha optimiser: train_genomicPDARTS.py:300-301."""
    import genome_nas_b200 as G
    h = G.Hyper()
    if workload == "pdarts3":
        h.arch_lr, h.adam_betas, h.arch_clip = 6e-4, (0.5, 0.999), h.clip
    return h


# --------------------------------------------------------------------------- reference arms
def _ref_model_and_optim(ref, torch, np, workload, dev):
    """The reference's own model / optimiser construction (train_genomicDARTS.py:206-292)."""
    torch.manual_seed(3)
    np.random.seed(3)
    sn, sr, sw = switches(workload)
    nc, nr = sum(sn[0]), sum(sw[0])
    a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s)).to(dev)) for s in ((14, nc), (14, nc), (36, nr))]
    p_skip = 0.1 if workload == "pdarts3" else 0.0
    model = ref.model_search.RNNModelSearch(SEQ, 0.05, 0.1, 8, NCLS, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings',
                                            sn, sr, sw, p_skip, *a).to(dev)
    conv, rhn = [], []
    for name, p in model.named_parameters():
        (rhn if "rnns" in name else conv).append(p)
    opt = torch.optim.SGD([{"params": conv}, {"params": rhn}], lr=0.025, weight_decay=3e-4)
    opt.param_groups[0]["momentum"] = 0.9
    opt.param_groups[1]["lr"] = 8.0
    opt.param_groups[1]["weight_decay"] = 5e-7
    from types import SimpleNamespace
    arch = ref.architect.Architect(model, torch.nn.BCELoss(), SimpleNamespace(wdecay=5e-7, clip=0.25, arch_lr=3e-3,
                                                                              arch_wdecay=1e-3))
    opt_a = torch.optim.Adam(model.arch_parameters(), lr=6e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    return model, conv, rhn, opt, opt_a, arch


def run_reference(args, rank, on_gpu=False):
    """CPU arm (`--impl reference`): the unmodified reference on all host cores, B sequences per step.
    `on_gpu` (internal `--impl reference-gpu`): the same code in PyTorch eager on the B200 with TF32 off --
    the `gpu_eager_baseline` of the native line ("the real kernel to beat")."""
    if rank != 0:
        return
    if not on_gpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = ""          # the reference picks `cuda` on its own when it sees one
    import logging
    logging.disable(logging.CRITICAL)                    # the reference logs every step
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import genomenas_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, K, W = args.batch, args.steps, args.warmup
    ref = build_ref.load()
    pdarts = args.workload == "pdarts3"
    if ref is None and on_gpu:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref not built"}), flush=True)
        return
    dev = torch.device("cuda" if on_gpu else "cpu")
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    batches = lambda k: (list(O.synthetic_batch(B, SEQ, NCLS, 300 + 2 * k)), list(O.synthetic_batch(B, SEQ, NCLS, 301 + 2 * k)))
    if ref is not None:
        kind = "reference"
        model, conv, rhn, opt, opt_a, arch = _ref_model_and_optim(ref, torch, np, args.workload, dev)
        crit = torch.nn.BCELoss()

        def run(n, seed0):
            tq, vq = zip(*[batches(seed0 + i) for i in range(n)])
            sync()
            t0 = time.perf_counter()
            # the reference's train() runs num_steps + 1 iterations (quirk Q3)
            ref.train_and_validate.train(list(tq), list(vq), model, rhn, conv, crit, opt, opt_a, arch, False, 0.025,
                                         0, n - 1, [5, 0.25, 0.25], 10 ** 9, 1e-3, True, True, pdarts, "TF_bindings")
            sync()
            return time.perf_counter() - t0
    else:
        kind = "port"
        shapes = torch.load(os.path.join(ROOT, "tests", "golden", "model_full.pt"), weights_only=False)["shapes"]
        P = O.synthetic_params(shapes, 43)
        cfg = O.SupernetConfig(dropouth=0.05, dropoutx=0.1)
        hyper, state = O.Hyper(), {}

        def run(n, seed0):
            nonlocal P
            t = 0.0
            for i in range(n):
                (xt, yt), (xv, yv) = batches(seed0 + i)
                mv = O.draw_rhn_masks(B, 512, cfg.dropoutx, cfg.dropouth)
                mt = O.draw_rhn_masks(B, 512, cfg.dropoutx, cfg.dropouth)
                t0 = time.perf_counter()
                P, _ = O.search_step(P, cfg, hyper, xt, yt, xv, yv, state, mv, mt)
                t += time.perf_counter() - t0
            return t

    # The batch is never reduced.  If K + W steps of B sequences do not fit the time budget, the STEP COUNT is
    # cut (and said so in the line): one untimed probe step sizes it.
    t_probe = run(1, 0)
    cut = False
    if t_probe * (K + W) > REF_BUDGET_S:
        cut = True
        W = min(W, 1)
        K = max(1, min(K, int(REF_BUDGET_S / t_probe) - W - 1))
    if W > 1:
        run(W - 1, 1000)                                  # the probe step is the first warm-up step
    total = run(K, 2000)
    value = B * K / total
    line = {
        "impl": "reference-gpu" if on_gpu else "reference", "metric": METRIC, "value": value, "unit": "seq/s",
        "n_gpus": args.gpus, "steps": K, "warmup": max(W, 1), "ms_per_step": 1e3 * total / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "global_batch": B, "device": str(dev),
                   "note": ("unmodified reference (byte-compiled into oracle/_ref): its RNNModelSearch, Architect and "
                            "train(), " if kind == "reference" else "CPU port of the reference path (oracle/), ")
                           + f"{B} sequences per step, {cores} host threads"
                           + ("; step count cut to fit the time budget" if cut else "")},
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": kind,
                         "sample": f"{K} search steps of {B} sequences each ({total:.1f} s)"},
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if on_gpu:
        line.pop("cpu_baseline")
        line["config"]["note"] = ("unmodified reference in PyTorch eager on the GPU (cuDNN/cuBLAS, TF32 off), "
                                  f"{B} sequences per step")
    print(json.dumps(line), flush=True)


def sub_bench(extra, timeout_s, hide_cuda):
    """Run another arm of this script in a child process and return its JSON line (or a reason)."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    if hide_cuda:
        env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True,
                             timeout=timeout_s, env=env, cwd=ROOT)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"no JSON line (rc {out.returncode}): {out.stderr.strip()[-300:]}"}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout_s} s"}


# --------------------------------------------------------------------------- clocks sampler
class Clocks(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------- algorithmic work
def cell_geometry():
    # (C, L_in of the edges from s0'/s1', L_out, stride)
    return [(8, 1000, 1000, 1), (16, 1000, 500, 2), (32, 500, 250, 2), (32, 250, 250, 1), (64, 250, 125, 2), (128, 125, 42, 3)]


PRIMS = ['none', 'max_pool_5', 'avg_pool_5', 'skip_connect', 'sep_conv_15', 'sep_conv_9', 'dil_conv_15', 'dil_conv_9',
         'conv_15']


def algorithmic_work(B, sw_normal, sw_reduce, T=42, H=512, n_rnn_edges=36):
    """Algorithmic FLOPs (1 multiply-add = 2) and bytes (fp32 tensors that MUST cross HBM once) of every kernel
    family for ONE search step = alpha pass (forward + data gradients) + weight pass (forward + data gradients +
    weight gradients).  Formulas: SURVEY.md 8(d), restated per family in DESIGN.md section 4.  An edge reads its
    input at L_e (= L_in for the 8 edges from s0'/s1', L_out for the 6 edges between nodes) and writes L_o."""
    w = {k: 0.0 for k in ("conv15_fb", "conv15_wg", "pw_fb", "pw_wg", "dw_fwd_b", "dw_bwd_b", "pool_fwd_b",
                          "pool_bwd_b", "mix_fwd_b", "bwd_reduce_b", "pre_fb", "pre_wg")}
    for ci, (C, Li, Lo, s) in enumerate(cell_geometry()):
        sw = sw_reduce if ci % 3 != 0 else sw_normal
        e = 0
        for node in range(4):
            n_terms = 0
            for j in range(2 + node):
                Le = Li if j < 2 else Lo
                on = lambda name: bool(sw[e][PRIMS.index(name)])
                el_i, el_o = 4.0 * B * C * Le, 4.0 * B * C * Lo
                if on('conv_15'):
                    f = 2 * (2.0 * B * C * C * Lo * 15)
                    w["conv15_fb"] += 2 * (f + f)          # two passes x (forward + data gradient)
                    w["conv15_wg"] += f
                n_pw = 2 * (on('sep_conv_15') + on('sep_conv_9')) + on('dil_conv_15') + on('dil_conv_9')
                if on('skip_connect') and j < 2 and s > 1:
                    n_pw += 1                               # FactorizedReduce = one C x C 1x1 conv at L_o
                w["pw_fb"] += 2 * 2 * n_pw * 2.0 * B * C * C * Lo
                w["pw_wg"] += n_pw * 2.0 * B * C * C * Lo
                n_dw1 = on('sep_conv_15') + on('sep_conv_9') + on('dil_conv_15') + on('dil_conv_9')
                n_dw2 = on('sep_conv_15') + on('sep_conv_9')
                # depthwise forward (fused per edge): read the edge input once, write one output per candidate;
                # second stage of sep_conv: read z1, write u2
                w["dw_fwd_b"] += 2 * ((el_i if n_dw1 else 0) + n_dw1 * el_o + n_dw2 * 2 * el_o)
                # depthwise backward (fused per edge): read every candidate's du and the edge input once,
                # read-modify-write dX once; second stage of sep_conv: read du2 and z1, write dyh1
                w["dw_bwd_b"] += 2 * (n_dw1 * el_o + (3 * el_i if n_dw1 else 0) + n_dw2 * 3 * el_o)
                n_pool = on('max_pool_5') + on('avg_pool_5')
                if n_pool:
                    w["pool_fwd_b"] += 2 * (el_i + n_pool * el_o)
                    w["pool_bwd_b"] += 2 * (el_o + 2 * el_i + n_pool * el_o)
                n_terms += sum(bool(x) for x in sw[e][1:])
                e += 1
            # node = sum of its terms: read every term once, write the node once (forward); the backward
            # reduction reads g and every term once (two passes each)
            w["mix_fwd_b"] += 2 * 4.0 * B * C * Lo * (n_terms + 1)
            w["bwd_reduce_b"] += 2 * 4.0 * B * C * Lo * (n_terms + 1)
    prev = [(24, 24), (24, 32), (32, 64), (64, 128), (128, 128), (128, 256)]   # (C_pp, C_p) per cell
    for (C, Li, Lo, s), (cpp, cp) in zip(cell_geometry(), prev):
        f = 2.0 * B * Li * C * (cpp + cp)
        w["pre_fb"] += 2 * 2 * f
        w["pre_wg"] += f
    rhn = T * (2.0 * B * 2 * H * 2 * H + n_rnn_edges * 2.0 * B * H * 2 * H)
    w["rhn_fwd"] = 2 * rhn
    w["rhn_bwd"] = 2 * rhn
    w["rhn_wg"] = rhn
    w["head_fb"] = 2 * 2 * 2.0 * B * (21504 * 925 + 925 * NCLS)
    w["head_wg"] = 2.0 * B * (21504 * 925 + 925 * NCLS)
    w["head_f"] = w["head_fb"] + w["head_wg"]                # the five GEMM pairs of the head per step (csrc/head.cu)
    w["optim_b"] = 38035471 * 4.0 * 6                       # sumsq read g; sgd read p, g, m, write p, m
    return w


# kernel name -> (family, bound, work key or None).  bound: "tensor" = tcgen05 3xTF32 kernels (graded against the
# measured bf16 tensor peak: every algorithmic multiply-add costs three TF32 MMAs at half the bf16 rate, so 1/6
# of that peak is the ceiling), "fp32" = CUDA-core FFMA kernels (nominal 72 TFLOP/s), "hbm" = memory-bound.
FAMILIES = [
    ("k_conv_tc", ",15>", "conv15_tc(fwd+dx)", "tensor", "conv15_fb"), ("k_conv_tc", ",5>", "conv15_tc(fwd+dx)", "tensor", "conv15_fb"),
    ("k_conv_tc", ",8>", "conv15_tc(fwd+dx)", "tensor", "conv15_fb"), ("k_conv_tc", "", "pw_tc(fwd+dx)", "tensor", "pw_fb"),
    ("k_conv_wgrad_tc", "", "conv15_wgrad_tc", "tensor", "conv15_wg"), ("k_pw_wgrad_tc", "", "pw_wgrad_tc", "tensor", "pw_wg"),
    ("k_rhn_wgrad_tc", "", "rhn_wgrad_tc", "tensor", "rhn_wg"), ("k_rhn_fwd_tc", "", "rhn_fwd_tc", "tensor", "rhn_fwd"),
    ("k_rhn_forward", "", "rhn_fwd", "fp32", "rhn_fwd"), ("k_rhn_bwd", "", "rhn_bwd", "fp32", "rhn_bwd"),
    ("k_rhn_wgrad", "", "rhn_wgrad", "fp32", "rhn_wg"),
    ("k_conv_wgrad", "", "conv15_wgrad", "fp32", None), ("k_pw_wgrad", "", "pre_wgrad", "fp32", "pre_wg"),
    ("k_dense_fwd", "KW = 15", "conv15_simt", "fp32", None), ("k_dense_bwd_dx", "KW = 15", "conv15_simt", "fp32", None),
    ("k_dense_fwd", "KW = 1;", "pre_1x1(fwd+dx)", "fp32", "pre_fb"), ("k_dense_bwd_dx", "KW = 1;", "pre_1x1(fwd+dx)", "fp32", "pre_fb"),
    ("k_dw2_fwd", "", "dw_fwd", "hbm", "dw_fwd_b"), ("k_dw2_bwd", "", "dw_bwd", "hbm", "dw_bwd_b"),
    ("k_edge_fwd", "", "dw_fwd", "hbm", "dw_fwd_b"), ("k_edge_bwd", "", "dw_bwd", "hbm", "dw_bwd_b"),
    ("k_pool_fwd", "", "pool_fwd", "hbm", "pool_fwd_b"), ("k_pool_bwd", "", "pool_bwd", "hbm", "pool_bwd_b"),
    ("k_mix_fwd", "", "mix_fwd", "hbm", "mix_fwd_b"), ("k_bwd_reduce", "", "bwd_reduce", "hbm", "bwd_reduce_b"),
    ("k_head", "", "head", "fp32", "head_f"), ("k_sgd", "", "optim", "hbm", "optim_b"), ("k_sumsq", "", "optim", "hbm", "optim_b"),
]


def family_of(name):
    for prefix, needle, fam, bound, key in FAMILIES:
        if name.startswith(prefix) and needle in name:
            return fam, bound, key
    return name.split(" ")[0].split("<")[0], None, None


def dbg(msg):
    if os.environ.get("GNAS_BENCH_DEBUG"):
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}\n")
        sys.stderr.flush()


def main():
    args = parse()
    if os.environ.get("GNAS_BENCH_DEBUG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["GNAS_BENCH_DEBUG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl != "native":
        run_reference(args, rank, on_gpu=args.impl == "reference-gpu")
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the native arm has no CPU path")
    import genome_nas_b200 as G
    from genome_nas_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import genomenas_oracle as O          # input synthesis only (before the timed region)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(3)
    np.random.seed(3)
    sn, sr, sw = switches(args.workload)
    nc, nr = sum(sn[0]), sum(sw[0])
    a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s))) for s in ((14, nc), (14, nc), (36, nr))]
    p_skip = 0.1 if args.workload == "pdarts3" else 0.0
    model = G.RNNModelSearch(SEQ, 0.05, 0.1, 8, NCLS, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sn, sr, sw, p_skip, *a)
    model.to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    dbg("model built")
    stepper = G.SearchStep(model, hyper=hyper_for(args.workload), process_group=None, world_size=world)
    if world > 1:
        dist.broadcast(stepper.region.flat, 0)
        torch.cuda.synchronize()
    dbg("parameters broadcast")
    torch.manual_seed(1000 + rank)       # per-rank dropout masks
    host = [t.pin_memory() for pair in (O.synthetic_batch(B, SEQ, NCLS, 2 * rank), O.synthetic_batch(B, SEQ, NCLS, 2 * rank + 1))
            for t in pair]
    xt, yt, xv, yv = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident():
        stepper.step(xt, yt, xv, yv)

    dbuf = [torch.empty_like(t, device=dev) for t in host]
    out_host = torch.empty(B, NCLS).pin_memory()
    loss_host = torch.empty(2).pin_memory()

    def step_e2e():
        for d, h in zip(dbuf, host):
            d.copy_(h, non_blocking=True)
        la, lw, raw, logits = stepper.step(dbuf[0], dbuf[1], dbuf[2], dbuf[3])
        out_host.copy_(logits, non_blocking=True)          # the reference reads logits back every step (:167)
        loss_host.copy_(torch.stack([la, lw]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_resident()
    torch.cuda.synchronize()
    dbg("first eager step done")
    mode = "eager"
    if not args.no_graph:
        try:
            stepper.capture(xt, yt, xv, yv, warmup=2)
            mode = "cuda-graph"
        except Exception as e:     # noqa: BLE001 -- report and keep measuring in eager mode
            sys.stderr.write(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); eager launches\n")
    dbg(f"launch mode {mode}")
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    dbg("warm-up done")
    clocks = Clocks(local)
    clocks.start()
    l0 = _lib.kernel_launches()
    ms = timed(step_resident, args.steps)
    dbg("timed region done")
    launches = _lib.kernel_launches() - l0
    if stepper.graph is not None:      # replays do not pass through the host-side launch counter
        launches = stepper.graph_launches * args.steps
    clocks.stop_flag = True
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clock_summary = clocks.summary()
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- N > 1: the NCCL-reduced gradient must equal the mean of the per-rank gradients (untimed self-check)
    dp_parity = None
    if world > 1:
        stepper._draw_static_masks() if stepper.graph is not None else None
        stepper.weight_backward(xt, yt)
        g = stepper.region.ensure_grads()
        local_g = g.clone()
        stepper.reduce_gradients()
        gathered = [torch.empty_like(local_g) for _ in range(world)]
        dist.all_gather(gathered, local_g)
        mean = torch.zeros_like(local_g, dtype=torch.float64)
        for t in gathered:
            mean += t.double()
        mean /= world
        err = float((g.double() - mean).norm() / (mean.norm() + 1e-30))
        dp_parity = {"rel_err": err, "ok": bool(err < 1e-6), "what": "NCCL allreduce / N vs fp64 mean of the all-gathered "
                     f"per-rank flat gradients ({local_g.numel()} floats, {world} ranks)"}
        del gathered, mean, local_g
        barrier()

    # ---- per-kernel device time of two more (untimed) steps -> roofline per kernel family
    roofline, fam_out = None, {}
    if rank == 0:
        ev_ms = float(_lib.lib().gnas_profile_event_overhead_ms(512, _lib.stream_ptr(xt)))
        # the timed steps overlap independent kernel chains on side streams; for per-kernel times the chains are
        # serialised on one stream, otherwise the event brackets of concurrent kernels would count the same time twice
        side_prev = _lib.side_streams(False)
        _lib.profile(True)
        nprof = 2
        stepper.world = 1            # rank-0-only pass: no collectives (the other ranks are already done)
        for _ in range(nprof):
            stepper._draw_static_masks() if stepper.graph is not None else None
            stepper._step_eager(xt, yt, xv, yv)
        rep = _lib.profile_report()
        _lib.profile(False)
        _lib.side_streams(side_prev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_bw = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json: bf16 sustained / copy bandwidth)" if peaks else \
            "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained, 6.65 TB/s)"
        work = algorithmic_work(B, sn, sr, n_rnn_edges=36)
        # ncu DRAM traffic per launch, when a capture of this build is committed (profiles/ncu_traffic.json:
        # {family: {"bytes_per_launch": ..., "source": "profiles/<file>"}}); never a hard-coded number
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        fam = {}
        for name, (cnt, tms) in rep.items():
            f, bound, key = family_of(name)
            d = fam.setdefault(f, {"launches": 0.0, "ms": 0.0, "bound": bound, "key": key})
            d["launches"] += cnt / nprof
            # every bracketed launch reads one empty event pair (measured live above) more than the kernel ran
            d["ms"] += max(tms - ev_ms * cnt, 0.0) / nprof
        for f, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            o = {"ms_per_step": round(d["ms"], 3), "launches_per_step": d["launches"], "bound": d["bound"]}
            wk = work.get(d["key"]) if d["key"] else None
            if wk and d["ms"] > 0:
                if d["bound"] == "hbm":
                    ach = wk / (d["ms"] / 1e3) / 1e9
                    o.update(achieved=round(ach, 1), peak=peak_bw, unit="GB/s", frac=round(ach / peak_bw, 4),
                             algorithmic_bytes_per_step=wk)
                else:
                    ach = wk / (d["ms"] / 1e3) / 1e12
                    o.update(achieved=round(ach, 2), peak=peak_tf, unit="TFLOP/s", frac=round(ach / peak_tf, 4),
                             algorithmic_flops_per_step=wk)
                    if d["bound"] == "fp32":
                        o["fp32_pipe_frac"] = round(ach / 72.0, 4)     # nominal FFMA rate of the pipe it runs on
                    else:
                        o["tf32x3_frac"] = round(ach / (peak_tf / 6.0), 4)   # 3 TF32 MMAs per multiply-add, TF32 = bf16/2
            t = traffic.get(f)
            o["traffic"] = t.get("bytes_per_launch") if isinstance(t, dict) else None
            if isinstance(t, dict):
                o["traffic_source"] = t.get("source")
            fam_out[f] = o
        graded = [f for f in fam_out if "frac" in fam_out[f]]
        if graded:
            top = max(graded, key=lambda f: fam_out[f]["ms_per_step"])
            t = fam_out[top]
            per_launch = t["launches_per_step"] or 1.0
            roofline = {"kernel": top, "bound": t["bound"], "achieved": t["achieved"], "peak": t["peak"],
                        "unit": t["unit"], "frac": t["frac"], "traffic": t["traffic"], "peak_source": peak_src,
                        "launches_per_step": t["launches_per_step"], "ms_per_step": t["ms_per_step"],
                        "ms_per_launch": round(t["ms_per_step"] / per_launch, 4),
                        "event_overhead_us": round(1e3 * ev_ms, 2),
                        "note": "achieved = algorithmic work of the family per step / CUDA-event time of its launches "
                                "(events on the launching stream, two untimed eager steps after the timed region with the "
                                "library's side streams switched off so that every kernel runs alone, one live-measured "
                                "empty event pair subtracted per launch; the timed steps overlap independent chains, so "
                                "the family times add up to more than ms_per_step); per-family figures incl. the "
                                "HBM-bound MixedOp kernels under `families`"}
            for k in ("fp32_pipe_frac", "tf32x3_frac", "traffic_source"):
                if k in t:
                    roofline[k] = t[k]
            hbm = [f for f in graded if fam_out[f]["bound"] == "hbm"]
            if hbm:
                tot_b = sum(fam_out[f]["algorithmic_bytes_per_step"] for f in hbm)
                tot_ms = sum(fam_out[f]["ms_per_step"] for f in hbm)
                roofline["mixedop_hbm_kernels"] = {"families": hbm, "ms_per_step": round(tot_ms, 3),
                                                   "achieved_GBps": round(tot_b / (tot_ms / 1e3) / 1e9, 1),
                                                   "frac_of_hbm_peak": round(tot_b / (tot_ms / 1e3) / 1e9 / peak_bw, 4)}

    cpu_baseline = gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the reference's own CPU path on this box's host cores: ONE full search step of 64 sequences (~15-30 s)
        r = sub_bench(["--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", args.workload,
                       "--batch", str(B)], 900, hide_cuda=True)
        cpu_baseline = r.get("cpu_baseline") or {"unavailable": r.get("unavailable", "no result")}
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        torch.cuda.empty_cache()
        r = sub_bench(["--impl", "reference-gpu", "--steps", "3", "--warmup", "2", "--workload", args.workload,
                       "--batch", str(B)], 900, hide_cuda=False)
        gpu_eager = {k: r.get(k) for k in ("value", "unit", "steps", "ms_per_step", "unavailable") if k in r}
        gpu_eager["what"] = "unmodified reference, PyTorch eager on this B200 (cuDNN/cuBLAS, TF32 off), same search step"

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = out_host.numel() * 4 + loss_host.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "global_batch": world * B, "parallelism": f"dp{world}",
                       "launch": mode, "params": n_params, "l2": "per-step working set (~6 GB of activations/workspace + "
                       "152 MB weights) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clock_summary,
            "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "families": fam_out,
            "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager,
        }
        if dp_parity is not None:
            line["dp_parity"] = dp_parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
