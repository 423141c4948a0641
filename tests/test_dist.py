"""Batch-sharded data parallelism (N>1 path) on CPU: world size 2 over gloo, kernels on the emulator.
Allreduced gradient == mean of the per-shard gradients; both ranks end with identical parameters that equal a
single-process step on the averaged gradient (SURVEY.md 8(e) parity definition)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(device):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "emul")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import build_emul
    import genomenas_oracle as O
    import genome_nas_b200 as G
    from genome_nas_b200 import _lib
    _lib._use_library_for_tests(build_emul.build(), "cpu")
    g = torch.load(os.path.join(GOLDEN, "model_tiny.pt"), weights_only=False)
    cfg = g["cfg"]
    P = O.synthetic_params(g["shapes"], g["seed_params"])
    mk = lambda k: torch.nn.Parameter(P[k].clone())
    m = G.RNNModelSearch(cfg.seq_len, 0.0, 0.0, cfg.C, cfg.num_classes, cfg.layers, cfg.steps, cfg.multiplier,
                         cfg.stem_multiplier, True, 0.2, None, cfg.task, cfg.switches_normal, cfg.switches_reduce,
                         cfg.switches_rnn, 0.0, mk("alphas_normal"), mk("alphas_reduce"), mk("weights"))
    m.load_state_dict(P)
    m.dropout_lin.p = 0.0
    m.train()
    shards = [O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 10 + r) for r in range(2)]
    return G, m, shards


def _worker(rank, port, out_dir):
    import torch.distributed as dist
    torch.set_num_threads(2)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)
    G, m, shards = _build("cpu")
    st = G.SearchStep(m, world_size=2)
    x, y = shards[rank]
    st.weight_backward(x, y)
    local = st.region.gflat.clone()
    st.reduce_gradients()
    st.apply_weight_update()
    torch.save({"local_grad": local, "reduced_grad": st.region.gflat.clone(), "flat": st.region.flat.clone()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    mean = 0.5 * (r0["local_grad"] + r1["local_grad"])
    assert torch.allclose(r0["reduced_grad"], mean, rtol=0, atol=1e-7 * float(mean.abs().max()) + 1e-12)
    assert torch.equal(r0["reduced_grad"], r1["reduced_grad"])
    assert torch.equal(r0["flat"], r1["flat"])
    # single process: accumulate both shards, halve, same optimiser kernels
    G, m, shards = _build("cpu")
    st = G.SearchStep(m)
    st.weight_backward(*shards[0])
    st.weight_backward(*shards[1], zero=False)
    st.reduce_gradients(scale=0.5)
    st.apply_weight_update()
    err = float((st.region.flat - r0["flat"]).norm() / r0["flat"].norm())
    assert err < 1e-6, err
