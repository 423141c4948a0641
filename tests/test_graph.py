"""CUDA-graph mode of SearchStep (GPU only): captured replay == eager steps on the same batches and masks."""
import pytest
import torch

import genomenas_oracle as O
from test_model import build, load


@pytest.mark.gpu
def test_graph_replay_matches_eager():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import genome_nas_b200 as G
    from genome_nas_b200 import _lib
    _lib._use_library_for_tests(None, "cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = load("tiny")
    cfg = g["cfg"]
    batches = [t.to(dev) for pair in (O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 0),
                                      O.synthetic_batch(g["B"], cfg.seq_len, cfg.num_classes, 1)) for t in pair]
    results = []
    for mode in ("eager", "graph"):
        m, _ = build(g, dev)
        st = G.SearchStep(m)
        torch.manual_seed(5)          # CPU generator: RHN masks
        if mode == "graph":
            st.capture(*batches, warmup=2)
            for _ in range(2):
                out = st.step(*batches)
        else:
            for _ in range(4):
                out = st.step(*batches)
        torch.cuda.synchronize()
        results.append((st.region.flat.clone(), [o.clone() for o in out]))
    (fa, oa), (fb, ob) = results
    assert float((fa - fb).norm() / fa.norm()) < 1e-5
    for a, b in zip(oa, ob):
        assert float((a - b).abs().max()) < 1e-4 * max(1.0, float(a.abs().max()))
