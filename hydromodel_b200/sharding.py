"""Sharding of an ensemble over the GPUs of one box, one process per GPU.

Columns are independent (the Richards path has no lateral coupling: the only
"lateral" term is a local sink, reference richards_pde.py:305-377), so the data
path needs no collective: each rank integrates a contiguous block of columns.
The only exchange is the optional ensemble statistic -- an all-reduce of
(count, sum, sum of squares) per output row -- which works on any
torch.distributed backend (NCCL on GPUs, gloo on CPU for the tests).
"""
import numpy as np


def column_range(n_total, rank, world):
    """Contiguous block [first, first + count) of rank's columns; blocks differ
    by at most one column."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("column_range: bad rank / world / n_total")
    base, extra = divmod(n_total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def local_moments(values, valid=None):
    """(count, sum, sumsq) over the column axis 0 of ``values[col, row]``."""
    v = np.asarray(values, dtype=np.float64)
    ok = np.ones(v.shape, bool) if valid is None else np.asarray(valid, bool)
    v = np.where(ok, v, 0.0)
    return ok.sum(axis=0).astype(np.float64), v.sum(axis=0), (v * v).sum(axis=0)


def reduce_moments(count, total, total_sq, group=None):
    """All-reduce the three moment tensors (torch tensors on the backend's
    device) and return mean, standard deviation and count per row."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        for t in (count, total, total_sq):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = torch.clamp(count, min=1.0)
    mean = total / n
    var = torch.clamp(total_sq / n - mean * mean, min=0.0)
    return mean, var.sqrt(), count
