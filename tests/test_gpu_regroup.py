"""GPU parity of the regrouping evaluation kernel (spl_eval_regroup_kernel: warp-private per-bank-class FIFOs).

The per-query arithmetic is the plain kernel's, only the lane that evaluates a query changes, so the results must
be BIT-identical to the plain kernel for every ordering of the queries -- uniform random (FIFO path), raster order
(coherent batches bypass the FIFOs), mixtures, NaN and exterior points, batch sizes that are not multiples of
anything -- and equal to the oracle within the reordering-roundoff tolerance of tests/test_gpu_eval.py.
"""
import os

import numpy as np
import pytest

import splpak_b200 as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default", "exact"], autouse=True)
def basis_mode(request, monkeypatch):
    """Both forms of the basis (tests/test_gpu_eval.py): the regrouping and the plain kernel must agree bit for bit in each."""
    if request.param == "exact":
        monkeypatch.setenv("SPLPAK_B200_BASIS", "exact")
    else:
        monkeypatch.delenv("SPLPAK_B200_BASIS", raising=False)
    return request.param


def _eval(mode, ndim, q, coef, mn, mx, nodes, nderiv=None, real32=False):
    old = os.environ.get("SPLPAK_B200_EVAL")
    os.environ["SPLPAK_B200_EVAL"] = mode
    try:
        out, ierr = sp.eval_batch(ndim, q, coef, mn, mx, nodes, nderiv=nderiv, real32=real32)
    finally:
        if old is None:
            del os.environ["SPLPAK_B200_EVAL"]
        else:
            os.environ["SPLPAK_B200_EVAL"] = old
    assert ierr == 0
    return out


def _orderings(rng, ndim, nq, mn, mx):
    u = rng.random((nq, ndim)) * 1.3 - 0.15
    rand = mn + u * (mx - mn)
    # raster order of a regular grid (dimension 1 fastest), the csagrid use case
    m = max(2, int(round(nq ** (1.0 / ndim))))
    ax = [np.linspace(mn[d], mx[d], m) for d in range(ndim)]
    g = np.stack(np.meshgrid(*ax[::-1], indexing="ij")[::-1], axis=-1).reshape(-1, ndim)[:nq]
    mixed = rand.copy()
    blk = 4096
    for s in range(0, min(len(g), nq) - blk, 3 * blk):      # coherent runs inside a random stream
        mixed[s:s + blk] = g[s:s + blk]
    special = rand.copy()
    special[5] = np.nan
    special[77, 0] = np.nan
    special[100] = mn
    special[101] = mx
    special[200:232] = mn + 0.5 * (mx - mn)                  # one full warp of identical queries
    return {"random": rand, "raster": g, "mixed": mixed, "special": special}


@pytest.mark.parametrize("ndim,nodes,nq", [
    (2, [64, 64], 700_001), (2, [9, 33], 300_000),
    (3, [24, 24, 24], 1_000_003), (3, [5, 4, 6], 280_000), (3, [17, 30, 8], 400_000),
    (4, [12, 12, 12, 12], 500_000), (4, [4, 5, 4, 6], 270_000),
])
def test_regroup_bit_identical_to_plain(ndim, nodes, nq):
    rng = np.random.default_rng(100 * ndim + nodes[0])
    coef = rng.standard_normal(int(np.prod(nodes)))
    mn = -rng.random(ndim)
    mx = 1.0 + rng.random(ndim)
    for name, q in _orderings(rng, ndim, nq, mn, mx).items():
        a = _eval("plain", ndim, q, coef, mn, mx, nodes)
        b = _eval("regroup", ndim, q, coef, mn, mx, nodes)
        assert np.array_equal(a, b, equal_nan=True), (name, np.nanmax(np.abs(a - b)), int((a != b).sum()))


@pytest.mark.parametrize("nq", [1, 31, 33, 255, 257, 4097, 70_000])
def test_regroup_small_and_ragged_batches(nq):
    rng = np.random.default_rng(nq)
    nodes = [7, 6, 9]
    coef = rng.standard_normal(int(np.prod(nodes)))
    mn, mx = np.zeros(3), np.ones(3)
    q = rng.random((nq, 3)) * 1.2 - 0.1
    a = _eval("plain", 3, q, coef, mn, mx, nodes)
    b = _eval("regroup", 3, q, coef, mn, mx, nodes)       # forced below the size threshold
    assert np.array_equal(a, b)


def test_regroup_derivatives_bit_identical():
    rng = np.random.default_rng(3)
    nodes = [10, 11, 9]
    coef = rng.standard_normal(int(np.prod(nodes)))
    mn, mx = np.zeros(3), np.full(3, 2.0)
    q = rng.random((300_000, 3)) * 2.4 - 0.2
    for nd in ([1, 0, 0], [0, 2, 1], [2, 2, 2]):
        a = _eval("plain", 3, q, coef, mn, mx, nodes, nderiv=nd)
        b = _eval("regroup", 3, q, coef, mn, mx, nodes, nderiv=nd)
        assert np.array_equal(a, b), nd


def test_regroup_matches_oracle(oracle):
    rng = np.random.default_rng(9)
    nodes = [24, 24, 24]
    coef = rng.standard_normal(24 ** 3)
    mn, mx = np.zeros(3), np.ones(3)
    q = rng.random((600_000, 3)) * 1.2 - 0.1
    got = _eval("regroup", 3, q, coef, mn, mx, nodes)
    idx = rng.choice(len(q), 3000, replace=False)
    ref, _ = oracle.evaluate_batch(3, q[idx], coef, mn, mx, nodes)
    bound, _ = oracle.evaluate_batch(3, q[idx], np.abs(coef), mn, mx, nodes)
    eps = np.finfo(float).eps
    assert (np.abs(got[idx] - ref) <= np.maximum(64 * eps * np.abs(coef).max() * 216, (128 + 4 * 72) * eps * np.abs(bound))).all()


def test_default_dispatch_probes_the_query_order():
    """Without the override, large 3-D batches launch the order probe, the table padding and BOTH evaluation kernels
    (the probe's flag makes one of them exit at once): scattered queries -> regrouping kernel, raster order -> plain."""
    rng = np.random.default_rng(1)
    nodes = [24, 24, 24]
    coef = rng.standard_normal(24 ** 3)
    q = rng.random((1 << 19, 3))
    os.environ.pop("SPLPAK_B200_EVAL", None)
    n0 = sp.total_launches()
    out, ierr = sp.eval_batch(3, q, coef, [0, 0, 0], [1, 1, 1], nodes)
    assert ierr == 0 and sp.total_launches() - n0 == 4
    assert np.array_equal(out, _eval("plain", 3, q, coef, [0, 0, 0], [1, 1, 1], nodes))
    ax = np.linspace(0.0, 1.0, 81)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij")[::-1], axis=-1).reshape(-1, 3)      # 531,441 raster-ordered points
    out, ierr = sp.eval_batch(3, g, coef, [0, 0, 0], [1, 1, 1], nodes)
    assert ierr == 0
    assert np.array_equal(out, _eval("plain", 3, g, coef, [0, 0, 0], [1, 1, 1], nodes))


@pytest.mark.parametrize("ndim,nodes,nq", [
    (2, [64, 64], 400_001), (3, [24, 24, 24], 600_003), (3, [5, 4, 6], 280_000),
    (4, [12, 12, 12, 12], 300_000), (4, [4, 5, 4, 6], 270_000),
])
def test_real32_regroup_bit_identical_to_plain_and_matches_oracle(oracle32, basis_mode, ndim, nodes, nq):
    """REAL32 library, uniform form: the 32-class float regrouping kernel (spl_eval_regroup_f32_kernel, one FIFO column
    per lane) against the plain float kernel -- bit-identical for every ordering -- and against the float oracle."""
    if basis_mode == "exact":
        pytest.skip("the float regrouping kernel exists in the uniform form only")
    rng = np.random.default_rng(7 * ndim + nodes[0])
    coef = rng.standard_normal(int(np.prod(nodes))).astype(np.float32)
    mn = np.zeros(ndim)
    mx = np.ones(ndim)
    for name, q in _orderings(rng, ndim, nq, mn, mx).items():
        q = q.astype(np.float32)
        a = _eval("plain", ndim, q, coef, mn, mx, nodes, real32=True)
        b = _eval("regroup", ndim, q, coef, mn, mx, nodes, real32=True)
        assert a.dtype == np.float32
        assert np.array_equal(a, b, equal_nan=True), (name, np.nanmax(np.abs(a - b)), int((a != b).sum()))
        if name == "random":
            idx = rng.choice(len(q), 2000, replace=False)
            ref, _ = oracle32.evaluate_batch(ndim, q[idx], coef, mn, mx, nodes)
            bound, _ = oracle32.evaluate_batch(ndim, q[idx], np.abs(coef), mn, mx, nodes)
            eps32 = float(np.finfo(np.float32).eps)
            tol = np.maximum(8 * eps32 * np.abs(coef).max(),
                             (2 * 4 ** ndim + 8 + 4 * int(np.sum(nodes))) * eps32 * np.abs(bound))
            assert (np.abs(b[idx].astype(np.float64) - ref.astype(np.float64)) <= tol).all()
            d = sp.eval_batch(ndim, q, coef, mn, mx, nodes, real32=True)[0]       # default dispatch (order probe)
            assert np.array_equal(d, a, equal_nan=True)
