"""Multi-GPU parity (NCCL all-to-all pencil transposes): needs >= 2 GPUs on the box, skipped
otherwise.  Launches tests/mgpu_worker.py under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multigpu_parity(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
