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


def test_two_handles_on_two_devices_in_one_process():
    """one process, one handle per GPU, interleaved calls (include/perc_abi.h: 'one handle per host thread / GPU'): the
    kernels' shared-memory opt-in is a per-device function attribute and is tracked per handle -- the second device's
    labeling and solves must work and give the first device's results"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.dirname(HERE))
    import percolation_b200 as P
    m = n = 512
    with P.Lattice(P.SQUARE, m, n, 0, device=0) as A, P.Lattice(P.SQUARE, m, n, 0, device=1) as B:
        out = {}
        for stream in (0, 1):
            for name, L in (("a", A), ("b", B)):
                L.generate(4711, stream, int(0.8 * L.t), int(0.75 * L.nb))
            for name, L in (("b", B), ("a", A)):                          # interleaved: label on 1, then on 0
                L.label(P.MIXED)
            for mode in (0, 2, 1):                                         # deflated one-pass, plain one-pass, two-kernel form
                for name, L in (("a", A), ("b", B)):
                    L.set_solver(mode)
                    out[(name, stream, mode)] = (L.summary(), L.conduct(0, tol=1e-12, itmax=400000, voltages=False))
                assert out[("a", stream, mode)] == out[("b", stream, mode)]
            assert (A.site_labels() == B.site_labels()).all() and (A.sizes() == B.sizes()).all()
