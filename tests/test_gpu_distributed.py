"""Multi-GPU parity (one process per GPU, NCCL): runs tests/dist_worker.py under torchrun when at least two
GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29541 + world), os.path.join(here, "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(res.stdout[-3000:])
    print(res.stderr[-3000:])
    assert res.returncode == 0
