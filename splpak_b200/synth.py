"""Counter-based synthetic scattered data: u01(splitmix64(seed, index)).

The same stream is produced by numpy on the host and by torch on the device, so every rank and the
CPU baseline see identical points without any transfer: point i of the global stream depends only
on (seed, i).  x ~ U[0,1)^ndim, y = smooth(x) + noise, w ~ U(0.5, 1.5).
"""
from __future__ import annotations

import numpy as np

_G = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB


def _s64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _u01_numpy(idx, seed):
    z = idx.astype(np.uint64) * np.uint64(_G) + np.uint64((seed * _M2) & ((1 << 64) - 1))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _u01_torch(idx, seed):
    import torch

    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)

    z = idx * _s64(_G) + _s64(seed * _M2)
    z = (z ^ lsr(z, 30)) * _s64(_M1)
    z = (z ^ lsr(z, 27)) * _s64(_M2)
    z = z ^ lsr(z, 31)
    return lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def _smooth(x, xp):
    f = None
    for d in range(x.shape[1]):
        t = xp.sin(2.0 * x[:, d] + 0.3 * d) + 0.25 * x[:, d]
        f = t if f is None else f * t
    return f


def points_numpy(ndim, n, start=0, seed=42, weighted=True, dtype=np.float64):
    """Host copy of points [start, start+n): (x[n, ndim], y[n], w[n] or None)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    x = np.stack([_u01_numpy(idx * np.uint64(8) + np.uint64(d), seed) for d in range(ndim)], axis=1)
    noise = (_u01_numpy(idx * np.uint64(8) + np.uint64(6), seed) - 0.5) * 0.1
    y = _smooth(x, np) + noise
    w = 0.5 + _u01_numpy(idx * np.uint64(8) + np.uint64(7), seed) if weighted else None
    return x.astype(dtype), y.astype(dtype), (w.astype(dtype) if w is not None else None)


def points_torch(ndim, n, start=0, seed=42, weighted=True, device="cuda", dtype=None, chunk=1 << 24):
    """Device-resident points [start, start+n), generated in chunks; same values as points_numpy."""
    import torch

    dtype = dtype or torch.float64
    x = torch.empty((n, ndim), dtype=dtype, device=device)
    y = torch.empty(n, dtype=dtype, device=device)
    w = torch.empty(n, dtype=dtype, device=device) if weighted else None
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(start + lo, start + hi, dtype=torch.int64, device=device)
        xc = torch.stack([_u01_torch(idx * 8 + d, seed) for d in range(ndim)], dim=1)
        noise = (_u01_torch(idx * 8 + 6, seed) - 0.5) * 0.1
        x[lo:hi] = xc.to(dtype)
        y[lo:hi] = (_smooth(xc, torch) + noise).to(dtype)
        if weighted:
            w[lo:hi] = (0.5 + _u01_torch(idx * 8 + 7, seed)).to(dtype)
    return x, y, w


def queries_numpy(ndim, n, start=0, seed=43, dtype=np.float64):
    idx = np.arange(start, start + n, dtype=np.uint64)
    return np.stack([_u01_numpy(idx * np.uint64(8) + np.uint64(d), seed) for d in range(ndim)], axis=1).astype(dtype)


def queries_torch(ndim, n, start=0, seed=43, device="cuda", dtype=None, chunk=1 << 24, raster=False, out=None):
    """Uniform random queries (default) or raster order of a regular output grid (csagrid use case)."""
    import torch

    dtype = dtype or torch.float64
    x = out if out is not None else torch.empty((n, ndim), dtype=dtype, device=device)
    if raster:
        m = int(round(n ** (1.0 / ndim)))
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            idx = torch.arange(start + lo, start + hi, dtype=torch.int64, device=device)
            k = idx
            for d in range(ndim):
                x[lo:hi, d] = ((k % m).to(torch.float64) / max(m - 1, 1)).to(dtype)
                k = k // m
        return x
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(start + lo, start + hi, dtype=torch.int64, device=device)
        for d in range(ndim):
            x[lo:hi, d] = _u01_torch(idx * 8 + d, seed).to(dtype)
    return x
