"""multi-rank host logic on CPU: world_size 2, gloo backend (SURVEY §8e mode 1)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from percolation_b200.shard import Stats, my_realizations, stream_id
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = my_realizations(11, rank, world)
    st = Stats(nbins=8)
    for i in mine:
        h = np.zeros(8, np.int64)
        h[i % 8] = i + 1
        st.add(G=0.5 * (i + 1), f=0.1 * i, iters=100 + i, maxcs=i, spans=(i % 2 == 0), hist=h)
    st.allreduce(dist)
    q.put((rank, mine, [stream_id(rank, i) for i in mine], st.v.tolist(), st.hist.tolist()))
    dist.destroy_process_group()


def test_round_robin_sharding_and_stats_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    all_idx = sorted(res[0][1] + res[1][1])
    assert all_idx == list(range(11))                       # every realization exactly once
    assert not set(res[0][2]) & set(res[1][2])              # disjoint Philox streams
    # both ranks hold the same merged block, equal to the serial sums
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]
    want_G = sum(0.5 * (i + 1) for i in range(11))
    assert res[0][3][0] == 11 and abs(res[0][3][2] - want_G) < 1e-12
    assert res[0][3][1] == 6 and res[0][3][6] == sum(100 + i for i in range(11))
    want_h = np.zeros(8, np.int64)
    for i in range(11):
        want_h[i % 8] += i + 1
    assert res[0][4] == want_h.tolist()
