"""The C oracle (oracle/splpak_oracle.c) against the SECOND, independently written restatement
(oracle/numpy_model.py, transliterated from src/splpak.F90, not from the C).  A transcription error would have to
be made twice, identically, for both to agree:

  * least-squares rows (data rows + derivative-constraint rows, incl. the :899 quirk): BIT-equal
  * coefficients: equal to a few ulp * cond(A) (the two suprls restatements sum in different orders)
  * splde for every derivative order, inside and outside the domain: equal to 4 ulp of sum |c_j phi_j|
  * error codes
Random cases come from hypothesis (ndim 1-4, nodes 4-9, weights incl. zeros, xtrap 0/1, holes, exterior points).
"""
import itertools

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import Oracle
from oracle.numpy_model import SplpakModel

EPS = np.finfo(float).eps


def make_case(seed, ndim, nodes, ndata, weighted, zero_w, hole, exterior):
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-2.0, 1.0, ndim)
    hi = lo + rng.uniform(0.5, 3.0, ndim)
    x = lo + (hi - lo) * rng.random((ndata, ndim))
    if exterior:
        x[: max(1, ndata // 10)] = lo + (hi - lo) * (rng.random((max(1, ndata // 10), ndim)) * 1.6 - 0.3)
    if hole:
        c = 0.5 * (lo + hi)
        keep = np.linalg.norm((x - c) / (hi - lo), axis=1) > 0.28
        x = x[keep]
    # a few points exactly on nodes / domain corners
    x[0] = lo
    x[-1] = hi
    if len(x) > 4:
        x[1] = lo + (hi - lo) * np.round(rng.random(ndim) * (np.array(nodes) - 1)) / (np.array(nodes) - 1)
    y = np.sin(x.sum(axis=1)) + 0.1 * rng.standard_normal(len(x))
    w = None
    if weighted:
        w = rng.uniform(0.5, 1.5, len(x))
        if zero_w:
            w[rng.random(len(x)) < 0.2] = 0.0
            w[0] = max(w[0], 0.5)          # w[0] >= 0 keeps the fit weighted (:796)
    return x, y, w, lo, hi


def dense_rows(rows, ncol):
    A = np.zeros((len(rows), ncol))
    r = np.zeros(len(rows))
    for i, (row, rhs) in enumerate(rows):
        for c, v in row.items():
            A[i, c] = v
        r[i] = rhs
    return A, r


case_strategy = st.tuples(
    st.integers(0, 10**6),                       # seed
    st.integers(1, 4),                           # ndim
    st.lists(st.integers(4, 9), min_size=4, max_size=4),
    st.booleans(), st.booleans(), st.booleans(), st.booleans(),
    st.sampled_from([0.0, 1.0, 0.37]),
)


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(case_strategy)
def test_rows_bit_equal(case):
    seed, ndim, nodes4, weighted, zero_w, hole, exterior, xtrap = case
    nodes = nodes4[:ndim]
    if ndim == 4:
        nodes = [min(n, 5) for n in nodes]
    if ndim == 3:
        nodes = [min(n, 7) for n in nodes]
    ndata = {1: 60, 2: 150, 3: 250, 4: 300}[ndim]
    x, y, w, lo, hi = make_case(seed, ndim, nodes, ndata, weighted, zero_w, hole, exterior)
    A, r = Oracle().rows(ndim, x, y, w, lo, hi, nodes, xtrap)
    rows, ncol = SplpakModel().rows(ndim, x, y, w, lo, hi, nodes, xtrap)
    A2, r2 = dense_rows(rows, ncol)
    assert A.shape == A2.shape, (A.shape, A2.shape)
    assert np.array_equal(A, A2), np.abs(A - A2).max()
    assert np.array_equal(r, r2)


@pytest.mark.parametrize("ndim,nodes,xtrap,weighted,hole", [
    (1, [10], 1.0, True, False),
    (1, [7], 0.0, False, False),
    (2, [6, 5], 1.0, True, True),
    (2, [5, 7], 0.0, True, False),
    (3, [4, 5, 4], 1.0, False, True),
    (4, [4, 4, 4, 4], 1.0, True, False),
])
def test_coefficients_agree(ndim, nodes, xtrap, weighted, hole):
    ncol = int(np.prod(nodes))
    x, y, w, lo, hi = make_case(11 * ndim + len(nodes), ndim, nodes, max(6 * ncol, 200), weighted, True, hole, True)
    o = Oracle()
    c1, ie1 = o.initialize(ndim, x, y, w, lo, hi, nodes, xtrap)
    c2, ie2 = SplpakModel().splcw(ndim, x, y, w if w is not None else [-1.0], lo, hi, nodes, xtrap)
    assert ie1 == 0 and ie2 == 0, (ie1, ie2)
    A, _ = o.rows(ndim, x, y, w, lo, hi, nodes, xtrap)
    cond = np.linalg.cond(A)
    err = np.abs(c1 - c2).max() / np.abs(c1).max()
    assert err <= 50 * EPS * cond, (err, cond)
    # both equal the least-squares solution of the (bit-equal) rows
    r = np.zeros(len(A))
    rows, _ = SplpakModel().rows(ndim, x, y, w, lo, hi, nodes, xtrap)
    for i, (_, rhs) in enumerate(rows):
        r[i] = rhs
    cl = np.linalg.lstsq(A, r, rcond=None)[0]
    assert np.abs(cl - c1).max() / np.abs(cl).max() <= 100 * EPS * cond


@pytest.mark.parametrize("nwrk_extra", [0, 7, 400])
def test_workspace_schedule_variants(nwrk_extra):
    """The Householder/Givens schedule depends on the scratch size (SURVEY A); both restatements must follow it."""
    ndim, nodes = 1, [9]
    ncol = 9
    x, y, w, lo, hi = make_case(5, ndim, nodes, 80, True, False, False, False)
    nwrk = ncol * (ncol + 1) + 1 + nwrk_extra
    c1, ie1 = Oracle().initialize(ndim, x, y, w, lo, hi, nodes, 1.0, nwrk=nwrk)
    c2, ie2 = SplpakModel().splcw(ndim, x, y, w, lo, hi, nodes, 1.0, nwrk=nwrk)
    assert ie1 == 0 and ie2 == 0
    assert np.abs(c1 - c2).max() <= 1e-12 * np.abs(c1).max()


def test_single_row_batches_use_givens():
    """nn just above the minimum leaves room for exactly one new row per reduction once the triangle is full:
    the Givens branch (:1488-1515) of both restatements."""
    ndim, nodes, ncol = 1, [6], 6
    x, y, w, lo, hi = make_case(8, ndim, nodes, 40, True, False, False, False)
    nreq = ((ncol + 5) * ncol + 2) // 2
    nwrk = nreq + 6                      # xtrap = 0: nn = nwrk
    c1, ie1 = Oracle().initialize(ndim, x, y, w, lo, hi, nodes, 0.0, nwrk=nwrk)
    c2, ie2 = SplpakModel().splcw(ndim, x, y, w, lo, hi, nodes, 0.0, nwrk=nwrk)
    assert ie1 == 0 and ie2 == 0
    assert np.abs(c1 - c2).max() <= 1e-12 * np.abs(c1).max()


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(0, 10**6), st.integers(1, 4), st.lists(st.integers(4, 9), min_size=4, max_size=4))
def test_splde_all_orders(seed, ndim, nodes4):
    nodes = nodes4[:ndim]
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-1.0, 1.0, ndim)
    hi = lo + rng.uniform(0.5, 2.0, ndim)
    coef = rng.standard_normal(int(np.prod(nodes)))
    o, m = Oracle(), SplpakModel()
    pts = lo + (hi - lo) * (rng.random((6, ndim)) * 1.5 - 0.25)
    pts[0] = lo
    pts[1] = hi
    orders = list(itertools.product(range(3), repeat=ndim))
    rng.shuffle(orders)
    for nd in orders[:12]:
        for p in pts:
            v1, e1 = o.evaluate(ndim, p, coef, lo, hi, nodes, nderiv=list(nd))
            v2, e2 = m.splde(ndim, p, list(nd), coef, lo, hi, nodes)
            assert e1 == e2 == 0
            scale = np.abs(coef).max() * 4 ** ndim * np.prod([(1.0 / ((hi[d] - lo[d]) / (nodes[d] - 1))) ** nd[d] * 6 for d in range(ndim)])
            assert abs(v1 - v2) <= 8 * EPS * scale, (nd, p, v1, v2)


def test_error_codes_agree():
    o, m = Oracle(), SplpakModel()
    x = np.random.default_rng(0).random((30, 1))
    y = x[:, 0]
    for args, want in [
        ((1, x, y, None, [0.0], [1.0], [3], 1.0), 102),
        ((1, x, y, None, [1.0], [1.0], [5], 1.0), 103),
    ]:
        _, ie = o.initialize(*args)
        _, ie2 = m.splcw(args[0], args[1], args[2], [-1.0], *args[4:])
        assert ie == ie2 == want
    # too few rows -> 107 (suprls error 33)
    _, ie = o.initialize(1, x[:3], y[:3], None, [0.0], [1.0], [6], 0.0)
    _, ie2 = m.splcw(1, x[:3], y[:3], [-1.0], [0.0], [1.0], [6], 0.0)
    assert ie == ie2 == 107
    # nderiv out of range: 104 without returning (:1190-1194)
    coef = np.arange(5.0)
    v1, e1 = o.evaluate(1, [0.3], coef, [0.0], [1.0], [5], nderiv=[3])
    v2, e2 = m.splde(1, [0.3], [3], coef, [0.0], [1.0], [5])
    assert e1 == e2 == 104 and v1 == v2


def test_algorithm_matched_cpu_baseline_is_a_correct_fit():
    """bench.py's optional 'algorithm-matched' CPU row (oracle/cpu_matched.py: sparse normal equations by window +
    LAPACK band Cholesky) must solve the same problem: its coefficients agree with the oracle's suprls."""
    from oracle import cpu_matched

    rng = np.random.default_rng(4)
    for ndim, nodes, n in [(1, [9], 300), (2, [6, 7], 2500), (3, [5, 6, 5], 6000)]:
        x = rng.random((n, ndim)) * 1.1 - 0.05
        y = np.cos(x.sum(axis=1))
        w = rng.uniform(0.5, 1.5, n)
        mn, mx = [0.0] * ndim, [1.0] * ndim
        c, _, _ = cpu_matched.fit(ndim, x, y, w, mn, mx, nodes)
        ref, ie = Oracle().initialize(ndim, x, y, w, mn, mx, nodes, 0.0)
        assert ie == 0
        A, _ = Oracle().rows(ndim, x, y, w, mn, mx, nodes, 0.0)
        assert np.abs(c - ref).max() <= 10 * EPS * np.linalg.cond(A) ** 2 * np.abs(ref).max()
