"""Slab decomposition on real GPUs (needs >= 2): one lattice over 2 ranks, NCCL all-gather of the
interface rows, halo exchange + all-reduce in the Kirchhoff solve; against the single-GPU run."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_slab_two_gpus_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL does not accept two ranks on one device)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "slab_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "slab-rank0-ok" in out.stdout and "slab-rank1-ok" in out.stdout
