"""CPU tests: the oracle against the reference's own tests, analytic known answers and an
independent least-squares solve.  No GPU, no product code."""
import json
import os

import numpy as np
import pytest

from util import make_problem

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_linear_test(oracle):
    """/root/reference/test/splpak_test_linear.f90:41-89 (inputs, ierror, errmax <= 1e-1, slope 2 within 1e-12)."""
    g = json.load(open(os.path.join(GOLD, "splpak_test_linear.json")))
    x = np.array(g["xdata"]).reshape(-1, 1)
    coef, ierr = oracle.initialize(1, x, g["ydata"], g["wdata"], g["xmin"], g["xmax"], g["nodes"], g["xtrap"],
                                   nwrk=g["nwrk"])
    assert ierr == 0
    errmax = 0.0
    for xe in g["x_est"]:
        f, ie = oracle.evaluate(1, [xe], coef, g["xmin"], g["xmax"], g["nodes"])
        assert ie == 0
        errmax = max(errmax, abs(f - 2.0 * xe))
    assert errmax <= g["errmax_tol"]
    fl, ie = oracle.evaluate(1, [0.0], coef, g["xmin"], g["xmax"], g["nodes"], nderiv=[1])
    assert ie == 0 and abs(fl - 2.0) <= g["slope_tol"]
    fr, ie = oracle.evaluate(1, [1.0], coef, g["xmin"], g["xmax"], g["nodes"], nderiv=[1])
    assert ie == 0 and abs(fr - 2.0) <= g["slope_tol"]
    # analytic coefficients of a linear function
    np.testing.assert_allclose(27.0 * coef, g["coef_times_27"], atol=1e-12)
    # linear extrapolation outside the grid
    assert abs(oracle.evaluate(1, [1.5], coef, g["xmin"], g["xmax"], g["nodes"])[0] - 3.0) < 1e-12
    assert abs(oracle.evaluate(1, [-0.7], coef, g["xmin"], g["xmax"], g["nodes"])[0] + 1.4) < 1e-12


def test_reference_noisy_test(oracle):
    """/root/reference/test/splpak_test.f90:50-84: tolerance-only (gfortran's random stream is not portable)."""
    rng = np.random.default_rng(42)
    n = 20
    r = (rng.random(n) - 0.5) / 10.0
    x = (np.arange(n) / (n - 1)).reshape(-1, 1)
    f1 = lambda t: 0.5 * (t * np.exp(-t) + np.sin(t))
    y = f1(x[:, 0]) + r
    w = 1.0 - np.abs(r)
    coef, ierr = oracle.initialize(1, x, y, w, [0.0], [1.0], [10], 1.0, nwrk=111)
    assert ierr == 0
    errmax = max(abs(oracle.evaluate(1, [i / 100], coef, [0.0], [1.0], [10])[0] - f1(i / 100)) for i in range(100))
    assert errmax <= 1e-1


def test_constant_function_K3(oracle):
    """f == 1 on 9 nodes -> [-1/3, 1/3, 2/3 x5, 1/3, -1/3]."""
    x = np.linspace(0, 1, 50).reshape(-1, 1)
    coef, ierr = oracle.initialize(1, x, np.ones(50), None, [0.0], [1.0], [9], 0.0)
    assert ierr == 0
    np.testing.assert_allclose(coef, [-1 / 3, 1 / 3] + [2 / 3] * 5 + [1 / 3, -1 / 3], atol=1e-12)


@pytest.mark.parametrize("ndim,nodes", [(2, [6, 7]), (3, [5, 4, 6])])
def test_multilinear_K1(oracle, ndim, nodes):
    """Multilinear functions are reproduced to roundoff, inside and outside the grid (K1)."""
    rng = np.random.default_rng(3)
    x = rng.random((600, ndim))
    a = rng.random(ndim) + 0.5
    b = rng.random(ndim) - 0.5
    f = lambda p: np.prod(a + b * p, axis=-1)
    coef, ierr = oracle.initialize(ndim, x, f(x), None, [0] * ndim, [1] * ndim, nodes, 0.0)
    assert ierr == 0
    q = rng.random((50, ndim)) * 1.6 - 0.3
    v, ie = oracle.evaluate_batch(ndim, q, coef, [0] * ndim, [1] * ndim, nodes)
    assert ie == 0
    np.testing.assert_allclose(v, f(q), rtol=0, atol=5e-12)
    nd = [1] * ndim
    v, _ = oracle.evaluate_batch(ndim, q, coef, [0] * ndim, [1] * ndim, nodes, nderiv=nd)
    np.testing.assert_allclose(v, np.full(len(q), np.prod(b)), atol=1e-10)
    nd = [2] + [0] * (ndim - 1)
    v, _ = oracle.evaluate_batch(ndim, q, coef, [0] * ndim, [1] * ndim, nodes, nderiv=nd)
    np.testing.assert_allclose(v, 0.0, atol=1e-9)


@pytest.mark.parametrize("ndim,nodes,xtrap,hole", [(1, [12], 1.0, False), (2, [6, 7], 0.0, False),
                                                   (2, [7, 6], 1.0, True), (3, [4, 5, 4], 1.0, True)])
def test_against_numpy_lstsq(oracle, ndim, nodes, xtrap, hole):
    """suprls restatement == numpy.linalg.lstsq on the oracle's own rows (independent solver)."""
    x, y, w, mn, mx = make_problem(ndim, nodes, 800, seed=ndim, weighted=True, hole=hole)
    coef, ierr = oracle.initialize(ndim, x, y, w, mn, mx, nodes, xtrap)
    assert ierr == 0
    A, r = oracle.rows(ndim, x, y, w, mn, mx, nodes, xtrap)
    if xtrap != 0 and hole:
        assert A.shape[0] > len(x), "constraint rows were expected to fire"
    ref = np.linalg.lstsq(A, r, rcond=None)[0]
    np.testing.assert_allclose(coef, ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_workspace_size_independence(oracle):
    """The Householder/Givens schedule depends on nn; results agree to roundoff."""
    x, y, w, mn, mx = make_problem(2, [5, 5], 300, seed=7)
    ncol = 25
    base, _ = oracle.initialize(2, x, y, w, mn, mx, [5, 5], 1.0)
    nreq = ((ncol + 5) * ncol + 2) // 2
    for nwrk in (ncol + nreq, ncol + nreq + 1, ncol + nreq + 26, ncol * (ncol + 1), 5 * ncol * ncol):
        c, ierr = oracle.initialize(2, x, y, w, mn, mx, [5, 5], 1.0, nwrk=nwrk)
        assert ierr == 0
        np.testing.assert_allclose(c, base, atol=1e-11 * np.abs(base).max())


def test_error_codes(oracle):
    x, y, w, mn, mx = make_problem(1, [10], 30, seed=1)
    assert oracle.initialize(0, x, y, w, mn, mx, [10], 1.0)[1] == 101
    assert oracle.initialize(1, x, y, w, mn, mx, [3], 1.0)[1] == 102
    assert oracle.initialize(1, x, y, w, [0.0], [0.0], [10], 1.0)[1] == 103
    assert oracle.initialize(1, x, y, w, mn, mx, [10], 1.0, ncf=9)[1] == 104
    assert oracle.initialize(1, x, y, w, mn, mx, [10], 1.0, ndata=0)[1] == 105
    assert oracle.initialize(1, x, y, w, mn, mx, [10], 1.0, nwrk=10)[1] == 106
    assert oracle.initialize(1, x, y, w, mn, mx, [10], 1.0, nwrk=50)[1] == 107        # suprls 32
    assert oracle.initialize(1, x[:5], y[:5], w[:5], mn, mx, [10], 0.0)[1] == 107     # suprls 33
    assert oracle.initialize(1, x, y, np.zeros(30), mn, mx, [10], 0.0)[1] == 107      # all weights zero
    coef = np.ones(10)
    assert oracle.evaluate(0, [0.5], coef, mn, mx, [10])[1] == 101
    assert oracle.evaluate(1, [0.5], coef, mn, mx, [3])[1] == 102
    assert oracle.evaluate(1, [0.5], coef, [1.0], [1.0], [10])[1] == 103
    assert oracle.evaluate(1, [0.5], coef, mn, mx, [10], nderiv=[3])[1] == 104


def test_real32_oracle(oracle, oracle32):
    x, y, w, mn, mx = make_problem(2, [6, 6], 500, seed=5)
    c64, _ = oracle.initialize(2, x, y, w, mn, mx, [6, 6], 1.0)
    c32, ierr = oracle32.initialize(2, x, y, w, mn, mx, [6, 6], 1.0)
    assert ierr == 0
    np.testing.assert_allclose(c32, c64, atol=2e-3 * np.abs(c64).max())
