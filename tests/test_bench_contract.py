"""bench.py contract checks that need no GPU: the CPU arm (`--impl reference`, the oracle port of the
reference algorithm on the host cores) runs and prints ONE JSON line with the keys the driver reads;
our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--L", "96", "--steps", "1",
                          "--warmup", "0", "--cpu-cg-iters", "2", "--cpu-threads", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "realizations/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C3 square mixed site/bond")


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stdout + out.stderr)


def test_e2e_host_inputs_are_the_drivers_orders():
    """the host-side inputs of bench.py's e2e leg: a site permutation of 1..t and the bond list permuted as pairs"""
    import importlib.util
    import numpy as np
    import torch
    import percolation_b200 as P
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    m = n = 24
    t = m * n
    b1, b2 = P.geom_bondlist(P.SQUARE, m, n, 0)
    nb = len(b1)
    ins = bench.e2e_host_inputs(np.random.default_rng(5), t, nb, b1, b2, 2, lambda k: torch.empty(k, dtype=torch.int32))
    assert len(ins) == 2
    for hs, hb in ins:
        so, bo = hs.numpy(), hb.numpy()
        assert so.dtype == np.int32 and sorted(so.tolist()) == list(range(1, t + 1))
        assert sorted(zip(bo[:nb].tolist(), bo[nb:].tolist())) == sorted(zip(b1.tolist(), b2.tolist()))
    assert (ins[0][0].numpy() != ins[1][0].numpy()).any()


def test_sweep_points_follow_the_reference_table_rule():
    """bench.py's sweep points: pb accumulated in fp64 by repeated + 5e-3 from pb* = k*/nb and truncated -- the rule of the
    reference's table (Sq/bond_cond.f:89-94, restated by the oracle's sweep_table); point 0 is the first-spanning fill itself"""
    import importlib.util
    from oracle import pyoracle as O
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    nb = 2 * 4096 * 4096 - 2 * 4096
    for kstar in (22388211, 22390015, 1, nb - 5):
        pbarr, nbarr = O.sweep_table(float(kstar) / float(nb), 5.0e-3, 9, nb)
        got = [bench.sweep_fill(kstar, nb, j) for j in range(9)]
        assert got[0] == kstar
        for j in range(1, 9):
            assert got[j] == min(nb, max(kstar, int(nbarr[j]))), (kstar, j, got[j], nbarr[j])
        assert all(a <= b for a, b in zip(got, got[1:]))
