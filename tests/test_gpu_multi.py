"""GPU, needs >= 2 GPUs on the box (skipped otherwise): the row-sharded multi-GPU path, one rank per GPU
under torchrun over NCCL, every rank's columns bit-exact against the oracle (tests/mg_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["nvlink", "exchange", "replicate"])
def test_sharded_path_matches_oracle(mode):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tests", "mg_worker.py"), mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mg ok" in r.stdout
