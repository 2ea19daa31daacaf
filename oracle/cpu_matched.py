"""'Algorithm-matched' CPU baseline (optional second CPU row) -- TEST / BENCH INFRASTRUCTURE.

The reference pushes a dense ncol-wide row per point through a streaming QR (~2 ncol^2 flops per point); the GPU
path assembles the sparse normal equations (4^ndim nonzeros per row) and solves them with a band Cholesky.  To
separate the ALGORITHMIC gain from the hardware gain, this module runs the GPU path's algorithm on the host with
numpy / BLAS / LAPACK on all cores the BLAS uses:

    window-sorted points -> per-window P^T W^2 P (BLAS dsyrk-like matmul of the (m x 4^ndim) basis block)
    -> lower band storage -> scipy.linalg.cholesky_banded + cho_solve_banded (LAPACK dpbtrf / dpbtrs)

Value-only basis arithmetic is oracle/numpy_model.window_weights_numpy (same formulas as bascmp).  No derivative
constraint rows (the benchmark's cfg3 has none firing).  Only bench.py's CPU leg and tests/ import this.
"""
from __future__ import annotations

import time

import numpy as np

from .numpy_model import window_weights_numpy


def assemble_banded(ndim, x, y, w, xmin, xmax, nodes):
    """Returns (ab, g, seconds): lower band storage ab[i-j, j] of G = B^T W^2 B and g = B^T W^2 y."""
    nodes = [int(v) for v in nodes]
    n = int(np.prod(nodes))
    strides = [int(np.prod(nodes[:d])) for d in range(ndim)]
    bw = min(n - 1, 3 * sum(strides))
    t0 = time.perf_counter()
    ws, b = [], []
    for d in range(ndim):
        s, v = window_weights_numpy(np.ascontiguousarray(x[:, d]), float(xmin[d]), float(xmax[d]), nodes[d])
        ws.append(s)
        b.append(v)
    # window id and the point order sorted by it
    wid = np.zeros(len(y), dtype=np.int64)
    mult = 1
    for d in range(ndim):
        wid += ws[d] * mult
        mult *= nodes[d] - 3
    order = np.argsort(wid, kind="stable")
    wid_s = wid[order]
    starts = np.flatnonzero(np.r_[True, wid_s[1:] != wid_s[:-1]])
    ends = np.r_[starts[1:], len(wid_s)]
    # basis block of every point: P[p, (k_ndim..k_1)] = prod_d b_d[k_d], dimension 1 fastest
    P = np.ones((len(y), 1))
    for d in range(ndim):
        P = (b[d].T[:, :, None] * P[:, None, :]).reshape(len(y), -1)      # new dim is the slower index
    if w is not None:
        P = P * w[:, None]
        ry = w * y
    else:
        ry = y
    P = P[order]
    ry = ry[order]
    local = np.array([sum(((idx >> (2 * d)) & 3) * strides[d] for d in range(ndim)) for idx in range(4 ** ndim)],
                     dtype=np.int64)
    li, lj = np.meshgrid(local, local, indexing="ij")
    low = li >= lj
    off_i, off_j = (li - lj)[low], lj[low]
    ab = np.zeros((bw + 1, n))
    g = np.zeros(n)
    flat = ab.reshape(-1)
    for s, e in zip(starts, ends):
        Pw = P[s:e]
        base = 0
        wv = int(wid_s[s])
        for d in range(ndim):
            base += (wv % (nodes[d] - 3)) * strides[d]
            wv //= nodes[d] - 3
        blk = Pw.T @ Pw
        np.add.at(flat, off_i * n + (off_j + base), blk[low])
        np.add.at(g, local + base, Pw.T @ ry[s:e])
    return ab, g, time.perf_counter() - t0


def solve_banded(ab, g):
    """Returns (coef, seconds) through LAPACK's band Cholesky."""
    from scipy.linalg import cho_solve_banded, cholesky_banded

    t0 = time.perf_counter()
    c = cholesky_banded(ab, lower=True, overwrite_ab=False, check_finite=False)
    coef = cho_solve_banded((c, True), g, check_finite=False)
    return coef, time.perf_counter() - t0


def fit(ndim, x, y, w, xmin, xmax, nodes):
    ab, g, t_asm = assemble_banded(ndim, x, y, w, xmin, xmax, nodes)
    coef, t_solve = solve_banded(ab, g)
    return coef, t_asm, t_solve
