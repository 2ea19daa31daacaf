"""The oracle against an INDEPENDENT THIRD-PARTY spline implementation: scipy.interpolate.BSpline.

No Fortran compiler exists in this image, so the reference binary cannot pin the oracle (oracle/README.md).  What can
be pinned without sharing a single line with the oracle's `bascmp`: the reference's natural spline with linear edge
functions (src/splpak.F90:302-379) is a uniform cubic B-spline series on the grid extended by one phantom node per side,

    s(x) = sum_{j=-1}^{n} a_j C(t - j),   t = (x - xmin)/dx,   C = 3/2 * (cardinal cubic B-spline),
    a_0 = 2 c_0 + 4 c_1, a_1 = 2 c_1, a_j = c_j, a_{n-2} = 2 c_{n-2}, a_{n-1} = 2 c_{n-1} + 4 c_{n-2},
    a_{-1} = 2 a_0 - a_1, a_n = 2 a_{n-1} - a_{n-2},

(the identity the CUDA evaluation kernels use).  Here scipy's B-splines evaluate the right-hand side:
  * splfe / splde (every derivative order, 1-D..3-D) == the scipy series inside the domain;
  * splcw / splcc with xtrap = 0 == numpy.linalg.lstsq on a design matrix built from scipy's basis (weights, 1-D and 2-D).
A wrong formula in the oracle's `bascmp` / `splde` / row loop -- or a wrong identity in the kernels -- fails here.
"""
import numpy as np
import pytest
from scipy.interpolate import BSpline

EPS = np.finfo(float).eps


def ext_matrix(n):
    """(n + 2) x n map c -> a (extended index -1 .. n at rows 0 .. n + 1)."""
    M = np.zeros((n + 2, n))
    for j in range(n):
        M[j + 1, j] = 1.0
    M[1, :] = 0.0
    M[1, 0], M[1, 1] = 2.0, 4.0                      # a_0
    M[2, :] = 0.0
    M[2, 1] = 2.0                                   # a_1
    M[n, :] = 0.0
    M[n, n - 1], M[n, n - 2] = 2.0, 4.0              # a_{n-1}
    M[n - 1, :] = 0.0
    M[n - 1, n - 2] = 2.0                            # a_{n-2}
    M[0, :] = 2.0 * M[1, :] - M[2, :]                # a_{-1}
    M[n + 1, :] = 2.0 * M[n, :] - M[n - 1, :]        # a_n
    return M


def scipy_basis(t, n, nu=0):
    """len(t) x (n + 2) matrix of 3/2 * B3(t - j), j = -1 .. n (nu-th derivative in t), from scipy alone."""
    knots = np.arange(-3.0, n + 3.0)                 # basis i is centred at i - 1
    out = np.empty((len(t), n + 2))
    for i in range(n + 2):
        c = np.zeros(n + 2)
        c[i] = 1.5
        b = BSpline(knots, c, 3, extrapolate=False)
        out[:, i] = b(t, nu=nu) if nu == 0 else b.derivative(nu)(t)
    return np.nan_to_num(out)


@pytest.mark.parametrize("n", [4, 5, 6, 7, 12, 24, 50])
def test_splde_1d_equals_scipy_bspline_series(oracle, n):
    rng = np.random.default_rng(n)
    coef = rng.standard_normal(n)
    xmin, xmax = -0.7, 2.3
    dx = (xmax - xmin) / (n - 1)
    x = np.concatenate([rng.uniform(xmin, xmax, 400), xmin + dx * (np.arange(n - 1) + 0.5), [xmin + 1e-9, xmax - 1e-9]])
    t = (x - xmin) / dx
    a = ext_matrix(n) @ coef
    for nu in (0, 1, 2):
        want = scipy_basis(t, n, nu) @ a / dx ** nu
        got, ie = oracle.evaluate_batch(1, x[:, None], coef, [xmin], [xmax], [n], nderiv=[nu])
        assert ie == 0
        scale = np.abs(scipy_basis(t, n, nu)) @ np.abs(a) / dx ** nu
        assert (np.abs(got - want) <= (64 + 8 * n) * EPS * np.maximum(scale, np.abs(a).max() / dx ** nu)).all(), nu


@pytest.mark.parametrize("nodes", [[5, 7], [6, 4, 9]])
def test_splde_nd_equals_scipy_tensor_series(oracle, nodes):
    ndim = len(nodes)
    rng = np.random.default_rng(10 * ndim)
    coef = rng.standard_normal(int(np.prod(nodes)))
    mn = -rng.random(ndim)
    mx = 1.0 + rng.random(ndim)
    dx = (mx - mn) / (np.array(nodes) - 1)
    x = mn + (mx - mn) * rng.random((300, ndim))
    # extended coefficient tensor: the 1-D map applied along every dimension (coef is dimension-1-fastest)
    A = coef.reshape(nodes[::-1])                                    # axes: (d_N, ..., d_1)
    for d in range(ndim):
        ax = ndim - 1 - d
        A = np.moveaxis(np.tensordot(ext_matrix(nodes[d]), A, axes=([1], [ax])), 0, ax)
    for nd in ([0] * ndim, [1] + [0] * (ndim - 1), [0] * (ndim - 1) + [2], [1] * ndim):
        B = [scipy_basis((x[:, d] - mn[d]) / dx[d], nodes[d], nd[d]) / dx[d] ** nd[d] for d in range(ndim)]
        if ndim == 2:
            want = np.einsum("pj,pi,ji->p", B[1], B[0], A)
        else:
            want = np.einsum("pk,pj,pi,kji->p", B[2], B[1], B[0], A)
        got, ie = oracle.evaluate_batch(ndim, x, coef, mn, mx, nodes, nderiv=nd)
        assert ie == 0
        scale = np.abs(coef).max() * 6.0 ** ndim * np.prod((3.0 / dx) ** np.array(nd))
        assert np.abs(got - want).max() <= 256 * EPS * scale, nd


@pytest.mark.parametrize("weighted", [False, True])
def test_splcw_1d_equals_lstsq_on_scipy_design_matrix(oracle, weighted):
    rng = np.random.default_rng(3)
    n, ndata = 14, 600
    xmin, xmax = 0.5, 3.0
    dx = (xmax - xmin) / (n - 1)
    x = rng.uniform(xmin, xmax, ndata)
    y = np.sin(2 * x) + 0.05 * rng.standard_normal(ndata)
    w = rng.uniform(0.5, 1.5, ndata) if weighted else None
    coef, ie = oracle.initialize(1, x[:, None], y, w, [xmin], [xmax], [n], 0.0)
    assert ie == 0
    D = scipy_basis((x - xmin) / dx, n) @ ext_matrix(n)             # design matrix in the reference's coefficients
    ww = w if weighted else np.ones(ndata)
    want = np.linalg.lstsq(D * ww[:, None], y * ww, rcond=None)[0]
    cond = np.linalg.cond(D * ww[:, None])
    assert np.abs(coef - want).max() <= 50 * EPS * cond * np.abs(want).max()


def test_splcc_2d_equals_lstsq_on_scipy_design_matrix(oracle):
    rng = np.random.default_rng(4)
    nodes, ndata = [6, 5], 900
    mn, mx = np.array([0.0, -1.0]), np.array([1.0, 1.0])
    dx = (mx - mn) / (np.array(nodes) - 1)
    x = mn + (mx - mn) * rng.random((ndata, 2))
    y = np.cos(3 * x[:, 0]) * x[:, 1] + 0.02 * rng.standard_normal(ndata)
    coef, ie = oracle.initialize(2, x, y, None, mn, mx, nodes, 0.0)
    assert ie == 0
    D1 = scipy_basis((x[:, 0] - mn[0]) / dx[0], nodes[0]) @ ext_matrix(nodes[0])
    D2 = scipy_basis((x[:, 1] - mn[1]) / dx[1], nodes[1]) @ ext_matrix(nodes[1])
    D = np.einsum("pj,pi->pji", D2, D1).reshape(ndata, -1)           # column = i + n1 * j (dimension 1 fastest)
    want = np.linalg.lstsq(D, y, rcond=None)[0]
    cond = np.linalg.cond(D)
    assert np.abs(coef - want).max() <= 50 * EPS * cond * np.abs(want).max()
