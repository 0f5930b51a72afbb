"""Realization sharding across ranks (multi-GPU mode 1): independent realizations
go round-robin to GPUs, there is no data-path collective, and ONE all-reduce merges the
statistics block at the end.  Pure host logic -- works with any torch.distributed backend
(NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np

STREAM_STRIDE = 1000003      # realization i of rank r uses Philox stream r * STREAM_STRIDE + i


def stream_id(rank, i):
    return rank * STREAM_STRIDE + i


def my_realizations(n_total, rank, world):
    """indices of the realizations rank `rank` runs (round-robin, SURVEY §8e mode 1)"""
    return list(range(rank, n_total, world))


class Stats:
    """order-independent statistics of a batch of realizations: integer counts are exact for any
    number of ranks; floating sums are accumulated in fp64"""
    FIELDS = ("count", "spanning", "sum_G", "sum_G2", "sum_f", "sum_f2", "iters", "maxcs_sum")

    def __init__(self, nbins=0):
        self.v = np.zeros(len(self.FIELDS), np.float64)
        self.hist = np.zeros(nbins, np.int64)

    def add(self, G=None, f=None, iters=0, maxcs=0, spans=False, hist=None):
        self.v[0] += 1
        self.v[1] += 1 if spans else 0
        if G is not None:
            self.v[2] += G
            self.v[3] += G * G
        if f is not None:
            self.v[4] += f
            self.v[5] += f * f
        self.v[6] += iters
        self.v[7] += maxcs
        if hist is not None:
            self.hist += hist

    def allreduce(self, dist=None, device="cpu"):
        """one collective for the floating block and one for the integer histogram"""
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return self
        import torch
        tv = torch.tensor(self.v, dtype=torch.float64, device=device)
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
        self.v = tv.cpu().numpy()
        if self.hist.size:
            th = torch.tensor(self.hist, dtype=torch.int64, device=device)
            dist.all_reduce(th, op=dist.ReduceOp.SUM)
            self.hist = th.cpu().numpy()
        return self

    def allreduce_native(self, lattice):
        """the same single reduction through the library's own NCCL communicator
        (perc_comm_init_rank + perc_allreduce_stats): what a Fortran driver calls"""
        iv, dv = lattice.allreduce_stats(self.hist, self.v)
        self.v = dv
        if self.hist.size:
            self.hist = iv
        return self

    def asdict(self):
        d = dict(zip(self.FIELDS, self.v.tolist()))
        n = max(d["count"], 1)
        d["mean_G"] = d["sum_G"] / n
        d["mean_f"] = d["sum_f"] / n
        return d
