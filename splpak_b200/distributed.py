"""Multi-GPU plumbing: one process per GPU, points sharded by contiguous index range, ONE all-reduce
(sum) of the partial normal equations before the replicated solve.  Evaluation shards the
queries the same way and needs no communication.

The only collective on the path is `torch.distributed.all_reduce` of the handle's partial buffer
[G in stencil storage | g | node histogram | totlwt | nrows] -- NCCL over NVLink on GPUs; the same
host logic runs over gloo with CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def partial_layout(ndim: int, nodes):
    """Offsets of the fields inside the partial buffer (float64 counts)."""
    ncol = int(np.prod(np.asarray(nodes, dtype=np.int64)[:ndim]))
    nst = 4 ** ndim
    off = {"S": (0, ncol * nst), "g": (ncol * nst, ncol * nst + ncol),
           "cnt": (ncol * nst + ncol, ncol * nst + 2 * ncol),
           "totlwt": (ncol * nst + 2 * ncol, ncol * nst + 2 * ncol + 1),
           "nrows": (ncol * nst + 2 * ncol + 1, ncol * nst + 2 * ncol + 2)}
    return off, ncol * nst + 2 * ncol + 2


def allreduce_partials(buf, group=None):
    """In-place sum of the partial buffer across ranks (no-op for a single process)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def fit_sharded(handle, x, y, w, weighted=True, device_resident=False, l1x=None, n=None, group=None, stream=None):
    """Every rank adds ITS shard to `handle`, the partial sums are all-reduced, every rank solves.
    Returns (coef, ierror); coef is identical on every rank (identical inputs to a deterministic solve)."""
    import torch

    if device_resident:
        rc = handle.add_points_device(x, l1x, y, w, n, weighted)
    else:
        rc = handle.add_points(x, y, w, weighted=weighted)
    if rc != 0:
        return None, rc
    part = handle.partial_tensor()
    if stream is not None:
        with torch.cuda.stream(stream):
            allreduce_partials(part, group)
    else:
        handle.synchronize()
        allreduce_partials(part, group)
        torch.cuda.synchronize()
    return handle.compute()
