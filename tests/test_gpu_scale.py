"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle's dense
O(ndata * ncol^2) QR cannot run these sizes):

  * analytic known answers that hold at any size: a constant / multilinear function is reproduced, and its
    coefficients are the Kronecker product of the 1-D closed forms K2/K3 of SURVEY 8c;
  * G c_one = g_one: the Gram matrix applied to the known coefficients of f = 1 must equal the right-hand side
    accumulated from y = 1 (two independent accumulators of the assembly kernel checked against each other);
  * the solver's backward error ||G c - g|| on the assembled system, and agreement with LAPACK's banded Cholesky
    (scipy.linalg.solveh_banded) on the same system;
  * streaming in chunks == one shot; evaluation of a subsample against the oracle's scalar splfe/splde.
cfg1 (the reference's own CPU-runnable case) is compared with the oracle directly, coefficients included.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import splpak_b200 as sp  # noqa: E402
from splpak_b200 import synth  # noqa: E402
from util import dense_from_stencil  # noqa: E402

EPS = np.finfo(float).eps


def c1d_linear(n, xmin, xmax, a, b):
    """K2: spline coefficients of f(x) = a + b x on n nodes."""
    dx = (xmax - xmin) / (n - 1)
    f = a + b * (xmin + dx * np.arange(n))
    c = 2.0 / 3.0 * f
    c[1] = f[1] / 3.0
    c[n - 2] = f[n - 2] / 3.0
    c[0] = -f[0] / 3.0 - 2.0 / 3.0 * dx * b
    c[n - 1] = -f[n - 1] / 3.0 + 2.0 / 3.0 * dx * b
    return c


def kron_coef(per_dim):
    """coef(nodes(1),...,nodes(ndim)), dimension 1 fastest, from 1-D factors."""
    c = per_dim[0]
    for cd in per_dim[1:]:
        c = np.multiply.outer(cd, c)
    return c.ravel()


def stencil_matvec(S, nodes, c):
    """G c with G in orthant-stencil storage S[node, sum_d delta_d 4^d] (common.cuh)."""
    nd = len(nodes)
    shape = tuple(reversed(nodes))                      # numpy axes: last axis = dimension 1
    C = c.reshape(shape)
    Sg = S.reshape(shape + (4 ** nd,))
    out = np.zeros(shape)
    for delta in itertools.product(range(4), repeat=nd):            # delta[d] for dimension d+1
        sten = sum(dl * 4 ** d for d, dl in enumerate(delta))
        for signs in itertools.product(*[((1,) if dl == 0 else (1, -1)) for dl in delta]):
            # row i, column j = i + signs*delta; node = min(i, j)
            sl_i, sl_j, sl_n = [], [], []
            ok = True
            for d in range(nd):
                n, dl, sg = nodes[d], delta[d], signs[d]
                if dl >= n:
                    ok = False
                    break
                if sg > 0:      # j = i + dl: i in [0, n-dl), node = i
                    si, sj = slice(0, n - dl), slice(dl, n)
                    sn = si
                else:           # j = i - dl: i in [dl, n), node = j
                    si, sj = slice(dl, n), slice(0, n - dl)
                    sn = sj
                sl_i.append(si), sl_j.append(sj), sl_n.append(sn)
            if not ok:
                continue
            ii = tuple(reversed(sl_i))
            jj = tuple(reversed(sl_j))
            nn = tuple(reversed(sl_n))
            out[ii] += Sg[nn + (sten,)] * C[jj]
    return out.ravel()


def band_from_stencil(S, nodes):
    """LAPACK lower band storage ab[k, j] = G[j + k, j] from S (only for the solveh_banded cross-check)."""
    nd = len(nodes)
    ncol = int(np.prod(nodes))
    strides = np.cumprod([1] + list(nodes[:-1]))
    bw = int(3 * strides.sum())
    ab = np.zeros((bw + 1, ncol))
    shape = tuple(reversed(nodes))
    Sg = S.reshape(shape + (4 ** nd,))
    lin = np.arange(ncol).reshape(shape)
    for delta in itertools.product(range(4), repeat=nd):
        sten = sum(dl * 4 ** d for d, dl in enumerate(delta))
        for signs in itertools.product(*[((1,) if dl == 0 else (1, -1)) for dl in delta]):
            off = sum(sg * dl * int(strides[d]) for d, (dl, sg) in enumerate(zip(delta, signs)))
            if off < 0:
                continue                                 # lower triangle: row = col + off, off >= 0
            sl_j, sl_n = [], []
            ok = True
            for d in range(nd):
                n, dl, sg = nodes[d], delta[d], signs[d]
                if dl >= n:
                    ok = False
                    break
                # column j, row i = j + sg*dl; node = min(i, j)
                if sg > 0:
                    sj = slice(0, n - dl)
                    sn = sj
                else:
                    sj = slice(dl, n)
                    sn = slice(0, n - dl)
                sl_j.append(sj), sl_n.append(sn)
            if not ok:
                continue
            jj = tuple(reversed(sl_j))
            nn = tuple(reversed(sl_n))
            cols = lin[jj].ravel()
            ab[off, cols] = Sg[nn + (sten,)].ravel()
    return ab, bw


def test_cfg1_full_parity_with_oracle(oracle):
    """BASELINE configs[0]: 1-D splcw fit of 10,000 noisy sin(x) samples on 50 nodes, splfe at 100,000 points --
    the one config the reference can run; compared with the oracle directly."""
    rng = np.random.default_rng(1)
    n, nodes = 10_000, [50]
    x = rng.random((n, 1)) * 2 * np.pi
    y = np.sin(x[:, 0]) + 0.05 * rng.standard_normal(n)
    w = rng.random(n) + 0.5
    mn, mx = [0.0], [2 * np.pi]
    ref, ie = oracle.initialize(1, x, y, w, mn, mx, nodes, 1.0)
    s = sp.SplpakType(quiet=True)
    got, ierr = s.initialize(1, x, 1, y, w, n, mn, mx, nodes, 1.0)
    assert ie == 0 and ierr == 0
    A, r = oracle.rows(1, x, y, w, mn, mx, nodes, 1.0)
    cond = np.linalg.cond(A.T @ A)
    tol = max(1e-10, 10 * EPS * cond)
    assert np.abs(got - ref).max() / np.abs(ref).max() <= tol, f"cond {cond:.2e}"
    q = rng.random((100_000, 1)) * 2 * np.pi
    v, ierr = sp.eval_batch(1, q, ref, mn, mx, nodes)
    assert ierr == 0
    pick = rng.integers(0, len(q), 4000)
    want, _ = oracle.evaluate_batch(1, q[pick], ref, mn, mx, nodes)
    np.testing.assert_allclose(v[pick], want, rtol=0, atol=50 * EPS * 4 * np.abs(ref).max())
    assert np.abs(v - np.sin(q[:, 0])).max() < 0.05


def test_cfg3_scale_constant_and_gram_consistency():
    """cfg3 size (1e8 weighted points, 24^3 nodes), y = 1, xtrap = 0: coefficients must be the Kronecker product of
    K3, and G c_one = g to accumulation roundoff."""
    ndim, nodes, n = 3, [24, 24, 24], 100_000_000
    x, y, w = synth.points_torch(ndim, n, seed=7)
    y.fill_(1.0)
    torch.cuda.synchronize()        # the handle launches on its own stream: inputs must be complete
    h = sp.FitHandle(ndim, [0] * 3, [1] * 3, nodes, 0.0)
    assert h.add_points_device(x, ndim, y, w, n, True) == 0
    S, g, cnt, totlwt, nrows = h.normal_equations()
    assert nrows == n
    c_one = kron_coef([c1d_linear(24, 0.0, 1.0, 1.0, 0.0)] * 3)
    Gc = stencil_matvec(S, nodes, c_one)
    np.testing.assert_allclose(Gc, g, rtol=0, atol=2e-12 * np.abs(g).max())
    dcoef = torch.zeros(24 ** 3, dtype=torch.float64, device="cuda")
    assert h.compute_device(dcoef) == 0
    got = dcoef.cpu().numpy()
    # backward error of the band Cholesky on the assembled system
    resid = stencil_matvec(S, nodes, got) - g
    assert np.abs(resid).max() <= 1e-11 * np.abs(g).max()
    # forward error: cond(G) at this size is ~1e7..1e8 (15^(2 ndim))
    assert np.abs(got - c_one).max() <= 1e-7
    h.destroy()
    del x, y, w


def test_cfg3_scale_fit_solver_vs_lapack_and_eval_vs_oracle(oracle):
    """cfg3 size, synthetic smooth function + noise, xtrap = 1, streamed in 8 chunks: backward error, agreement with
    LAPACK dpbsv on the SAME assembled system, fit-vs-truth, and splfe of a subsample against the oracle."""
    from scipy.linalg import solveh_banded

    ndim, nodes, n = 3, [24, 24, 24], 100_000_000
    x, y, w = synth.points_torch(ndim, n, seed=42)
    wsum = float(w.sum())
    torch.cuda.synchronize()
    h = sp.FitHandle(ndim, [0] * 3, [1] * 3, nodes, 1.0)
    step = n // 8
    for lo in range(0, n, step):
        assert h.add_points_device(x[lo:lo + step], ndim, y[lo:lo + step], w[lo:lo + step], step, True) == 0
    S, g, cnt, totlwt, nrows = h.normal_equations()
    assert nrows == n and abs(totlwt - wsum) <= 1e-9 * totlwt
    assert abs(cnt.sum() - totlwt) <= 1e-9 * totlwt                 # every in-range point lands on one node
    dcoef = torch.zeros(24 ** 3, dtype=torch.float64, device="cuda")
    assert h.compute_device(dcoef) == 0                            # no node is data sparse at 8,219 points per cell
    got = dcoef.cpu().numpy()
    resid = stencil_matvec(S, nodes, got) - g
    assert np.abs(resid).max() <= 1e-11 * np.abs(g).max()
    ab, bw = band_from_stencil(S, nodes)
    assert bw == 1803
    ref = solveh_banded(ab, g, lower=True)
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max()
    # one-shot on the same data agrees with the streamed fit
    h.reset()
    assert h.add_points_device(x, ndim, y, w, n, True) == 0
    d2 = torch.zeros_like(dcoef)
    assert h.compute_device(d2) == 0
    assert float((d2 - dcoef).abs().max()) <= 1e-7 * np.abs(ref).max()
    h.destroy()
    del x, y, w
    # evaluation at scale: 2e8 queries, subsample against the oracle, and against the truth
    nq = 200_000_000
    q = synth.queries_torch(ndim, nq)
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0] * 3, [1] * 3, nodes, out) == 0
    torch.cuda.synchronize()
    pick = torch.randint(0, nq, (3000,), device="cuda")
    qh, vh = q[pick].cpu().numpy(), out[pick].cpu().numpy()
    want, _ = oracle.evaluate_batch(ndim, qh, got, [0] * 3, [1] * 3, nodes)
    np.testing.assert_allclose(vh, want, rtol=0, atol=50 * EPS * 64 * np.abs(got).max())
    assert np.abs(vh - synth._smooth(qh, np)).max() < 5e-3


def test_cfg2_scale_constraints_and_gradient(oracle):
    """cfg2 size: 2-D splcc of 1e6 points on 64x64 nodes with a data hole (derivative constraints fire), linear f:
    the second-derivative constraint rows vanish on a linear function, so it is reproduced exactly; splde gradient at
    1e7 points."""
    ndim, nodes, n = 2, [64, 64], 1_000_000
    x, _, _ = synth.points_torch(ndim, int(n * 1.08), seed=11, weighted=False)
    keep = ((x - 0.5) ** 2).sum(dim=1) > 0.15 ** 2
    x = x[keep][:n].contiguous()
    assert x.shape[0] == n
    a, b = 0.7, np.array([1.3, -0.8])
    y = a + x @ torch.tensor(b, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    h = sp.FitHandle(ndim, [0, 0], [1, 1], nodes, 1.0)
    assert h.add_points_device(x, ndim, y, None, n, False) == 0
    dcoef = torch.zeros(64 * 64, dtype=torch.float64, device="cuda")
    assert h.compute_device(dcoef) == 0
    S, g, cnt, totlwt, nrows = h.normal_equations()
    assert nrows > n, "derivative-constraint rows were expected to fire around the data hole"
    got = dcoef.cpu().numpy()
    # The constraint rows carry dxin^2 * xtrap * (expected - found weight) ~ 1e6 against O(1) data rows, so G --
    # their squares -- is badly conditioned and the normal-equations solve loses eps*cond(G) (the
    # reference's QR works at sqrt(cond)).  Every tolerance below is scaled by the measured cond(G).
    ev = np.linalg.eigvalsh(dense_from_stencil(S, nodes))
    cond = ev[-1] / ev[0]
    assert ev[0] > 0
    tol = 10 * EPS * cond
    print(f"cfg2: cond(G) = {cond:.3e}, tolerance {tol:.3e}")
    want = kron_coef([c1d_linear(64, 0.0, 1.0, a, b[0]), c1d_linear(64, 0.0, 1.0, 1.0, 0.0)]) + \
        kron_coef([c1d_linear(64, 0.0, 1.0, 1.0, 0.0), c1d_linear(64, 0.0, 1.0, 0.0, b[1])])
    assert np.abs(got - want).max() <= tol * np.abs(want).max()
    # ... and two refinement passes over the same device-resident points recover the reference's accuracy
    for _ in range(2):
        assert h.refine_device(x, ndim, y, None, n, dcoef, weighted=False) == 0
    refined = dcoef.cpu().numpy()
    err_plain = np.abs(got - want).max() / np.abs(want).max()
    err_ref = np.abs(refined - want).max() / np.abs(want).max()
    print(f"cfg2: coefficient error vs the analytic answer: plain {err_plain:.2e}, refined {err_ref:.2e}")
    assert err_ref <= max(1e-9, 100 * EPS * np.sqrt(cond))
    tol = max(1e-9, 100 * EPS * np.sqrt(cond))
    got = refined
    h.destroy()
    nq = 10_000_000
    q = synth.queries_torch(ndim, nq)
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    for nder, want in (([1, 0], b[0]), ([0, 1], b[1])):
        assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0, 0], [1, 1], nodes, out, nderiv=nder) == 0
        torch.cuda.synchronize()
        assert float((out - want).abs().max()) <= tol * 63 * 3        # d/dx of a basis function is <= 3 * dxin
    assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0, 0], [1, 1], nodes, out) == 0
    torch.cuda.synchronize()
    f = a + q @ torch.tensor(b, dtype=torch.float64, device="cuda")
    assert float((out - f).abs().max()) <= tol * 4
    pick = torch.randint(0, nq, (2000,), device="cuda")
    qh = q[pick].cpu().numpy()
    want, _ = oracle.evaluate_batch(ndim, qh, got, [0, 0], [1, 1], nodes, nderiv=[0, 1])
    sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0, 0], [1, 1], nodes, out, nderiv=[0, 1])
    torch.cuda.synchronize()
    np.testing.assert_allclose(out[pick].cpu().numpy(), want, rtol=0, atol=50 * EPS * 16 * 63 * np.abs(got).max())


def test_cfg4_scale_multilinear():
    """cfg4 size: 4-D fit of 1e7 points on 12^4 nodes (20,736 coefficients, half bandwidth 5,655), xtrap = 0,
    multilinear f: coefficients are the Kronecker product of K2, values and the 4-fold mixed derivative are exact."""
    ndim, nodes, n = 4, [12, 12, 12, 12], 10_000_000
    x, _, w = synth.points_torch(ndim, n, seed=5)
    a = np.array([0.6, 1.1, 0.9, 1.4])
    b = np.array([0.5, -0.4, 0.3, -0.2])
    ta, tb = (torch.tensor(v, dtype=torch.float64, device="cuda") for v in (a, b))
    y = (ta + tb * x).prod(dim=1)
    torch.cuda.synchronize()
    h = sp.FitHandle(ndim, [0] * 4, [1] * 4, nodes, 0.0)
    assert h.add_points_device(x, ndim, y, w, n, True) == 0
    dcoef = torch.zeros(12 ** 4, dtype=torch.float64, device="cuda")
    assert h.compute_device(dcoef) == 0
    t = h.timings()
    h.destroy()
    got = dcoef.cpu().numpy()
    want = kron_coef([c1d_linear(12, 0.0, 1.0, a[d], b[d]) for d in range(4)])
    assert np.abs(got - want).max() <= 5e-6 * np.abs(want).max(), t
    nq = 2_000_000
    q = synth.queries_torch(ndim, nq) * 1.4 - 0.2                 # also outside the domain: linear extrapolation
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0] * 4, [1] * 4, nodes, out) == 0
    torch.cuda.synchronize()
    f = (ta + tb * q).prod(dim=1)
    # cond(G) ~ 15^(2*4) = 2.6e9 for dense uniform 4-D data: eps*cond = 6e-7
    assert float((out - f).abs().max()) <= 5e-6
    assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0] * 4, [1] * 4, nodes, out, nderiv=[1, 1, 1, 1]) == 0
    torch.cuda.synchronize()
    assert float((out - float(np.prod(b))).abs().max()) <= 5e-6 * 11 ** 4 * 0.1


def test_more_than_2_31_queries_and_points(oracle):
    """64-bit indexing end to end (the reference's default integers stop at 2^31 - 1): a 1-D fit of
    2.2e9 unweighted points (17 chunks of 2^27) of a linear function -> analytic K2 coefficients, then 2.2e9
    evaluations, checked against the oracle at both ends of the array.  Empty inputs are no-ops."""
    n = 2_200_000_000
    nodes, a, b = [50], 0.3, 1.7
    x = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    for lo in range(0, n, 1 << 28):
        hi = min(n, lo + (1 << 28))
        synth.queries_torch(1, hi - lo, start=lo, out=x[lo:hi])
    y = a + b * x[:, 0]
    torch.cuda.synchronize()
    h = sp.FitHandle(1, [0.0], [1.0], nodes, 0.0)
    assert h.add_points_device(x, 1, y, None, 0, False) == 0            # empty chunk: nothing happens
    assert h.add_points_device(x, 1, y, None, n, False) == 0
    dcoef = torch.zeros(50, dtype=torch.float64, device="cuda")
    assert h.compute_device(dcoef) == 0
    S, g, cnt, totlwt, nrows = h.normal_equations()
    assert nrows == n
    h.destroy()
    got = dcoef.cpu().numpy()
    want = c1d_linear(50, 0.0, 1.0, a, b)
    assert np.abs(got - want).max() <= 1e-9
    del y
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    assert sp.eval_batch_device(1, x, 1, 0, dcoef, [0.0], [1.0], nodes, out) == 0   # nq = 0
    assert sp.eval_batch_device(1, x, 1, n, dcoef, [0.0], [1.0], nodes, out) == 0
    torch.cuda.synchronize()
    for sl in (slice(0, 1000), slice(n - 1000, n), slice((1 << 31) - 500, (1 << 31) + 500)):
        qh = x[sl].cpu().numpy()
        ref, _ = oracle.evaluate_batch(1, qh, got, [0.0], [1.0], nodes)
        np.testing.assert_allclose(out[sl].cpu().numpy(), ref, rtol=0, atol=50 * EPS * 4 * np.abs(got).max())
    assert float((out - (a + b * x[:, 0])).abs().max()) <= 1e-8
