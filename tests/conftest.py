import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(params=[pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """'cuda': the product path (libgnas.so on a B200).  'emul': the same kernel sources compiled
    against the CUDA execution-model emulator -- CPU tier, checks kernel logic only."""
    import torch
    from genome_nas_b200 import _lib
    if request.param == "cuda":
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        _lib._use_library_for_tests(None, "cuda")
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        yield torch.device("cuda:0")
    else:
        import build_emul
        _lib._use_library_for_tests(build_emul.build(), "cpu")
        yield torch.device("cpu")
        _lib._use_library_for_tests(None, "cuda")


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def leaf(t, device):
    return t.detach().clone().to(device).requires_grad_(True)
