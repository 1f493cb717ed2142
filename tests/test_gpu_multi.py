"""N > 1 GPU path: slab decomposition with NCCL halo exchange must reproduce the single-GPU result.
Needs at least two visible GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"MG_CHUNKS": "3"}, {"MG_CHUNKS": "3", "MG_OVERLAP": "0"}],
                         ids=["default", "split-interior-boundary-chunks", "stream-ordered-exchange"])
def test_two_rank_slab_decomposition_matches_single_gpu(env):
    """MG_CHUNKS=3 forces three k-chunks per rank so that the overlapped exchange really launches the interior
    chunk before, and the two boundary chunks after, the halo planes have arrived."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "max rel diff" in r.stderr
