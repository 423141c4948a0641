"""Helper run in a subprocess by tests/test_simt_path.py with GNAS_TC=0: MixedOp / RHN cases at the channel counts the
tcgen05 kernels normally take, forced through the fp32 SIMT kernels, against the fp64 oracle."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import genomenas_oracle as O  # noqa: E402
import genome_nas_b200 as G  # noqa: E402


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def shapes(module, prefix):
    kinds = {"running_mean": "rm", "running_var": "rv", "num_batches_tracked": "nbt"}
    return {prefix + k: (tuple(v.shape), next((t for s, t in kinds.items() if k.endswith(s)), "w"))
            for k, v in module.state_dict().items()}


def main():
    assert os.environ.get("GNAS_TC") == "0"
    worst = 0.0
    for B, C, L, S in [(3, 32, 90, 1), (2, 64, 61, 2), (2, 128, 42, 3), (2, 16, 100, 1)]:
        torch.manual_seed(B * 1000 + C + L)
        sw = [True] * 9
        m = G.MixedOp(C, S, sw, 0.0)
        P = O.synthetic_params(shapes(m, "cell_ops.0."), 5)
        m.load_state_dict({k[len("cell_ops.0."):]: v for k, v in P.items()})
        m.to("cuda").train()
        x0, w0 = torch.randn(B, C, L), torch.softmax(torch.randn(9), 0)
        x, w = x0.clone().cuda().requires_grad_(True), w0.clone().cuda().requires_grad_(True)
        y = m(x, w)
        gy = torch.randn(y.shape)
        (y * gy.cuda()).sum().backward()
        Q = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
        xo, wo = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
        yo = O.mixed_op(xo, wo, sw, S, Q, "cell_ops.0", True, O.Recorder(), 0.0, None)
        (yo * gy.double()).sum().backward()
        errs = [relerr(y, yo), relerr(x.grad, xo.grad), relerr(w.grad, wo.grad)]
        errs += [relerr(p.grad, Q["cell_ops.0." + n].grad) for n, p in m.named_parameters()]
        worst = max(worst, max(errs))
        assert max(errs) < 5e-5, (B, C, L, S, max(errs))
    # the point of this case: the CUDA-core kernels really ran (GNAS_TC=0), no tensor-core kernel did
    from genome_nas_b200 import _lib
    _lib.profile(True)
    m.zero_grad()
    y = m(x.detach().requires_grad_(True), w.detach().requires_grad_(True))
    y.sum().backward()
    names = list(_lib.profile_report().keys())
    _lib.profile(False)
    assert any(n.startswith("k_dense_fwd") for n in names) and any(n.startswith("k_dense_bwd_dx") for n in names), names
    assert not any("_tc" in n for n in names), names
    print(f"simt path ok: worst rel err {worst:.2e}; {len(names)} distinct kernels, none on tensor cores")


if __name__ == "__main__":
    main()
