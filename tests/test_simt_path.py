"""The fp32 SIMT kernels stay in the library behind GNAS_TC=0 (and for shapes the tcgen05 kernels do not take).  The
switch is read once per process, so the SIMT run happens in a subprocess."""
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_simt_kernels_match_oracle_at_tensor_core_shapes():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, GNAS_TC="0")
    r = subprocess.run([sys.executable, os.path.join(HERE, "_simt_case.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "simt path ok" in r.stdout
