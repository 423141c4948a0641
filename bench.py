#!/usr/bin/env python
"""Benchmark of the genomeDARTS supernet search step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one search iteration: first-order alpha step on a validation batch (fwd + bwd + Adam on
the alphas) + weight step on a training batch (fwd + bwd with TAR, global-norm clip, SGD), i.e. the body
of generalNAS_tools/train_and_validate.py:55-152.  Workload = BASELINE configs[1]: full supernet
(C=8, 6 layers, 9 CNN / 5 RNN candidates, 38.0 M parameters), batch 64 per GPU, one-hot 4x1000 input,
919 labels, fp32.  N > 1 (torchrun): batch-sharded data parallel, NCCL allreduce of the flat gradient
buffer, weak scaling (64 sequences per GPU).  Prints ONE JSON line on rank 0.
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

WORKLOAD = "genomeDARTS supernet search step, batch 64/GPU, seq 1000, 919 labels (BASELINE configs[1])"
METRIC = "supernet train seqs/sec (fwd+bwd+alpha) @1000bp"
SEQ, NCLS, BATCH = 1000, 919, 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    return ap.parse_args()


def switches():
    return [[True] * 9 for _ in range(14)], [[True] * 9 for _ in range(14)], [[True] * 5 for _ in range(36)]


# --------------------------------------------------------------------------- reference arm (CPU oracle)
def run_reference(args, rank):
    """Times the CPU restatement of the reference path (oracle/, the reference is pure Python/PyTorch and is
    not present on the GPU box) on all host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import genomenas_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(3)
    # parameter shapes of the full supernet as recorded from the reference's own state_dict (golden fixture)
    shapes = torch.load(os.path.join(ROOT, "tests", "golden", "model_full.pt"), weights_only=False)["shapes"]
    P = O.synthetic_params(shapes, 43)
    cfg = O.SupernetConfig(dropouth=0.05, dropoutx=0.1)
    hyper, state = O.Hyper(), {}

    def one(bs, seed):
        nonlocal P
        xt, yt = O.synthetic_batch(bs, SEQ, NCLS, seed)
        xv, yv = O.synthetic_batch(bs, SEQ, NCLS, seed + 1)
        mv = O.draw_rhn_masks(bs, 512, cfg.dropoutx, cfg.dropouth)
        mt = O.draw_rhn_masks(bs, 512, cfg.dropoutx, cfg.dropouth)
        t0 = time.perf_counter()
        P, _ = O.search_step(P, cfg, hyper, xt, yt, xv, yv, state, mv, mt)
        return time.perf_counter() - t0

    bs = 8
    t_probe = one(bs, 100)                      # untimed probe: size the per-step sample
    budget = 200.0
    n_total = args.steps + args.warmup
    while bs > 2 and t_probe * n_total * (bs / 8.0) > budget:
        bs //= 2
    for w in range(args.warmup):
        one(bs, 200 + 2 * w)
    times = [one(bs, 300 + 2 * k) for k in range(args.steps)]
    total = sum(times)
    value = bs * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "cpu", "note": "CPU port of the reference path (oracle/), full "
                   "supernet, same search step; per-step sample reduced to bound the run"},
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} search steps of {bs} sequences (of 64) each"},
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
    # (C, L_edge_in for edges 0/1, L_out, stride)
    return [(8, 1000, 1000, 1), (16, 1000, 500, 2), (32, 500, 250, 2), (32, 250, 250, 1), (64, 250, 125, 2), (128, 125, 42, 3)]


def algorithmic_work(B):
    """FLOPs / bytes per launch family for one pass over the full supernet (all 84 edges, 9 candidates)."""
    w = {}
    conv15 = sum(2 * 14 * 2 * B * C * C * Lo * 15 for C, _, Lo, _ in cell_geometry())       # two dense convs per edge
    pw = sum(14 * 6 * 2 * B * C * C * Lo for C, _, Lo, _ in cell_geometry())                # 4 sep + 2 dil pointwise
    w["conv15_fwd_flops"] = conv15
    w["pw_fwd_flops"] = pw
    # conv_15 convolutions that run on the tcgen05 kernels (C >= 16: all 28 per cell; strided first convs count at
    # their output length), forward + data gradient
    w["conv15_tc_flops"] = sum(2 * 14 * 2 * B * C * C * Lo * 15 for C, _, Lo, _ in cell_geometry() if C >= 16)
    w["mix_bytes"] = sum(4 * B * C * Lo * (sum(8 * (2 + i) for i in range(4)) + 4) for C, _, Lo, _ in cell_geometry())
    return w


def family_of(name):
    if name.startswith("k_conv_tc"):
        return "conv15_tc(fwd+dx)" if ",15>" in name else "pw_tc(fwd+dx)"
    if name.startswith("k_conv_wgrad_tc"):
        return "conv15_wgrad_tc"
    if name.startswith("k_pw_wgrad_tc"):
        return "pw_wgrad_tc"
    if name.startswith("k_rhn_wgrad_tc"):
        return "rhn_wgrad_tc"
    if name.startswith("k_conv_wgrad"):
        return "conv15_wgrad"
    if name.startswith("k_dense_fwd") and "KW = 15" in name:
        return "conv15_fwd"
    if name.startswith("k_dense_bwd_dx") and "KW = 15" in name:
        return "conv15_dx"
    if name.startswith("k_pw_wgrad"):
        return "pw_wgrad"
    if name.startswith("k_dense_fwd") and "KW = 1;" in name:
        return "pw_fwd"
    if name.startswith("k_dense_bwd_dx") and "KW = 1;" in name:
        return "pw_dx"
    if name.startswith("k_rhn_forward"):
        return "rhn_fwd"
    if name.startswith("k_rhn_bwd_gemm"):
        return "rhn_bwd_gemm"
    if name.startswith("k_rhn_wgrad"):
        return "rhn_wgrad"
    if name.startswith("k_mix_fwd"):
        return "mix_fwd"
    return name.split(" ")[0]


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
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the native arm has no CPU path")
    import genome_nas_b200 as G
    from genome_nas_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import genomenas_oracle as O
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
    a = [torch.nn.Parameter(torch.FloatTensor(1e-3 * np.random.randn(*s))) for s in ((14, 9), (14, 9), (36, 5))]
    sn, sr, sw = switches()
    model = G.RNNModelSearch(SEQ, 0.05, 0.1, 8, NCLS, 6, 4, 4, 3, True, 0.2, None, 'TF_bindings', sn, sr, sw, 0.0, *a)
    model.to(dev).train()
    dbg("model built")
    stepper = G.SearchStep(model, process_group=None, world_size=world)
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

    # ---- per-kernel device time of two more (untimed) steps -> roofline of the dominant kernel family
    roofline = None
    fam = {}
    if rank == 0:
        _lib.profile(True)
        nprof = 2
        stepper.world = 1            # rank-0-only pass: no collectives (the other ranks are already done)
        for _ in range(nprof):
            stepper._draw_static_masks() if stepper.graph is not None else None
            stepper._step_eager(xt, yt, xv, yv)
        rep = _lib.profile_report()
        _lib.profile(False)
        # An event pair around a launch reads ~5.2 us more than the kernel ran (calibrated against the CUPTI
        # trace of the same step, profiles/r01_cupti_step_v3.txt: k_rhn_bwd_gemm 756 x 17.9 us vs 12.6 us,
        # k_rhn_bwd_elem 672 x 10.2 vs 5.1); without the correction families of tiny kernels look dominant.
        ev_overhead_ms = 5.2e-3
        for name, (cnt, tms) in rep.items():
            f = fam.setdefault(family_of(name), [0, 0.0])
            f[0] += cnt / nprof
            f[1] += max(tms - ev_overhead_ms * cnt, 0.2 * tms) / nprof
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        work = algorithmic_work(B)
        top = max(fam.items(), key=lambda kv: kv[1][1])[0] if fam else None
        # algorithmic FLOPs per step of each compute family (weight pass + alpha pass where applicable)
        flops = {"conv15_wgrad": work["conv15_fwd_flops"], "conv15_fwd": 2 * work["conv15_fwd_flops"],
                 "conv15_dx": 2 * work["conv15_fwd_flops"], "pw_wgrad": work["pw_fwd_flops"],
                 "pw_fwd": 2 * work["pw_fwd_flops"], "pw_dx": 2 * work["pw_fwd_flops"],
                 "rhn_fwd": 2 * 42 * (2 * B * 1024 * 1024 + 36 * 2 * B * 512 * 1024),
                 "rhn_bwd_gemm": 2 * 42 * (2 * B * 1024 * 1024 + 36 * 2 * B * 512 * 1024),
                 "rhn_wgrad": 42 * (2 * B * 1024 * 1024 + 36 * 2 * B * 512 * 1024),
                 "rhn_wgrad_tc": 42 * (2 * B * 1024 * 1024 + 36 * 2 * B * 512 * 1024)}
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the ncu captures under
        # profiles/ (r01_ncu_rhn_layer_T42_B64_H512.csv, r01_ncu_cell5_C128_L125to42.csv); None = not captured
        ncu_traffic = {"rhn_fwd": 2.826e7 + 4.126e8, "rhn_bwd_gemm": 4.121e6 + 5.1, "rhn_wgrad": 2.75e8 + 4.04e7}
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16 sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        if top in flops:
            launches_top, t_top = fam[top]
            ach = flops[top] / (t_top / 1e3) / 1e12
            roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": ach / peak_tf, "traffic": ncu_traffic.get(top), "peak_source": peak_src,
                        "launches_per_step": launches_top, "ms_per_step": t_top,
                        "fp32_pipe_frac": ach / 72.0,
                        "note": "algorithmic fp32 FLOPs (1 multiply-add = 2) / CUDA-event time of the launches; "
                                + ("fp32 CUDA-core kernel: the 336 dependent RHN steps stream the state through L2 and "
                                   "cannot use tcgen05 (hi/lo weight copies do not fit the weight-stationary shared "
                                   "memory); fp32_pipe_frac is against the ~72 TFLOP/s nominal FFMA rate"
                                   if not top.endswith("_tc") and "tc(" not in top else
                                   "tcgen05 3xTF32 kernel: 3 MMAs per algorithmic multiply-add")}
        elif top is not None:
            launches_top, t_top = fam[top]
            roofline = {"kernel": top, "bound": "hbm", "achieved": None, "peak": peaks.get("hbm_gbs", 6650.0),
                        "unit": "GB/s", "frac": None, "traffic": None, "launches_per_step": launches_top,
                        "ms_per_step": t_top}
        # the tensor-core families next to it (algorithmic fp32 FLOPs; every multiply-add costs three TF32 MMAs)
        tc_fl = {"conv15_tc(fwd+dx)": 4 * work["conv15_tc_flops"], "conv15_wgrad_tc": work["conv15_fwd_flops"],
                 "rhn_wgrad_tc": flops["rhn_wgrad_tc"]}
        if roofline is not None:
            roofline["tensor_core_kernels"] = {k: {"ms_per_step": round(fam[k][1], 3), "tflops": round(v / (fam[k][1] / 1e3) / 1e12, 1)}
                                               for k, v in tc_fl.items() if k in fam and fam[k][1] > 0}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        P = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        cfg = O.SupernetConfig(dropouth=0.05, dropoutx=0.1)
        bs = 16
        hx = [t[:bs].cpu() for t in host]
        mv = O.draw_rhn_masks(bs, 512, cfg.dropoutx, cfg.dropouth)
        mt = O.draw_rhn_masks(bs, 512, cfg.dropoutx, cfg.dropouth)
        t0 = time.perf_counter()
        O.search_step(P, cfg, O.Hyper(), hx[0], hx[1], hx[2], hx[3], {}, mv, mt)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": bs / dt, "unit": "seq/s", "cores": cores, "kind": "port",
                        "sample": f"1 search step of {bs} sequences (of 64), {dt:.1f} s on {cores} threads"}

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = out_host.numel() * 4 + loss_host.numel() * 4
        line = {
            "metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world}", "launch": mode,
                       "params": 38035471, "l2": "per-step working set (~6 GB of activations/workspace + 152 MB "
                       "weights) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clock_summary,
            "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "kernel_ms_per_step": {k: round(v[1], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:12]},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
