"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, partial-buffer layout and
the one all-reduce.  The per-rank partial normal equations come from the ORACLE's rows here (checker
only -- the product assembles them in CUDA); what is tested is that sharding + sum == unsharded."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from splpak_b200.distributed import allreduce_partials, partial_layout, shard_range
from util import make_problem


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 100, 10**8 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_partial_layout_matches_header_order():
    off, n = partial_layout(3, [24, 24, 24])
    assert off["S"] == (0, 13824 * 64) and off["g"][1] - off["g"][0] == 13824
    assert off["nrows"][1] == n == 13824 * 64 + 2 * 13824 + 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stencil_from_rows(A, nodes):
    """Orthant-stencil packing of A^T A (numpy, test side) -- same layout the CUDA assembly produces."""
    nodes = [int(v) for v in nodes]
    ndim = len(nodes)
    ncol = int(np.prod(nodes))
    G = A.T @ A
    S = np.zeros((ncol, 4 ** ndim))
    idx = np.arange(ncol)
    multi = np.stack([(idx // int(np.prod(nodes[:d]))) % nodes[d] for d in range(ndim)], axis=1)
    strides = np.array([int(np.prod(nodes[:d])) for d in range(ndim)])
    for i in range(ncol):
        for delta in np.ndindex(*([4] * ndim)):
            j = multi[i] + np.array(delta)
            if (j < nodes).all():
                S[i, int(sum(dd * 4 ** d for d, dd in enumerate(delta)))] = G[i, int((j * strides).sum())]
    return S


def _worker(rank, world, port, ndim, nodes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle

    o = Oracle()
    x, y, w, mn, mx = make_problem(ndim, nodes, 1200, seed=3)
    lo, hi = shard_range(len(x), rank, world)
    A, r = o.rows(ndim, x[lo:hi], y[lo:hi], w[lo:hi], mn, mx, nodes, 0.0)
    off, n = partial_layout(ndim, nodes)
    buf = torch.zeros(n, dtype=torch.float64)
    buf[off["S"][0]:off["S"][1]] = torch.from_numpy(_stencil_from_rows(A, nodes).reshape(-1))
    buf[off["g"][0]:off["g"][1]] = torch.from_numpy(A.T @ r)
    buf[off["totlwt"][0]] = float(w[lo:hi].sum())
    buf[off["nrows"][0]] = float(A.shape[0])
    allreduce_partials(buf)
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ndim,nodes", [(1, [8]), (2, [5, 6])])
def test_two_rank_allreduce_equals_single_rank(oracle, ndim, nodes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ndim, nodes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, y, w, mn, mx = make_problem(ndim, nodes, 1200, seed=3)
    A, r = oracle.rows(ndim, x, y, w, mn, mx, nodes, 0.0)
    off, n = partial_layout(ndim, nodes)
    S = _stencil_from_rows(A, nodes).reshape(-1)
    np.testing.assert_allclose(got[off["S"][0]:off["S"][1]], S, rtol=0, atol=1e-11 * np.abs(S).max())
    np.testing.assert_allclose(got[off["g"][0]:off["g"][1]], A.T @ r, rtol=0, atol=1e-11 * np.abs(A.T @ r).max())
    assert abs(got[off["totlwt"][0]] - w.sum()) < 1e-9 and got[off["nrows"][0]] == A.shape[0]
