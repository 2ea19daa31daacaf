"""GPU parity: batched splfe/splde (CUDA, through the C ABI) against the oracle's scalar loop."""
import numpy as np
import pytest

import splpak_b200 as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["default", "exact"], autouse=True)
def basis_mode(request, monkeypatch):
    """Every test runs twice: with the default dispatch (large batches whose extended table fits in shared memory use
    the uniform phantom-node form of the basis, csrc/basis.cuh) and with SPLPAK_B200_BASIS=exact (bascmp's node-by-node
    formulas, bit-identical 1-D values, everywhere)."""
    if request.param == "exact":
        monkeypatch.setenv("SPLPAK_B200_BASIS", "exact")
    else:
        monkeypatch.delenv("SPLPAK_B200_BASIS", raising=False)
    return request.param


CASES = [
    (1, [10]), (1, [4]), (1, [50]),
    (2, [6, 7]), (2, [4, 4]), (2, [64, 64]),
    (3, [5, 4, 6]), (3, [24, 24, 24]),
    (4, [4, 5, 4, 6]), (4, [12, 12, 12, 12]),
]


def _queries(rng, ndim, nq, mn, mx):
    q = rng.random((nq, ndim)) * 1.5 - 0.25          # inside and outside the grid
    q = mn + q * (mx - mn)
    # exact node / boundary hits
    q[0] = mn
    q[1] = mx
    q[2] = mn + (mx - mn) * 0.5
    return q


def _tol(coef, ndim, oracle=None, q=None, mn=None, mx=None, nodes=None):
    """Pure reordering roundoff: |delta| <= ~50 eps * sum_j |c_j Phi_j(x)|.  The sum is
    evaluated per query with the oracle on |coef| (value basis functions are non-negative, and they
    grow linearly outside the grid, so the bound must be per point, not global).  The uniform form of the basis
    (default for large batches) does not reproduce the reference's own rounding of the node positions (eps * node
    index in u, src/splpak.F90:246), hence the term 4 eps sum(nodes): it covers both forms."""
    eps = np.finfo(float).eps
    floor = 64 * eps * np.abs(coef).max() * 6.0 ** ndim
    if oracle is None:
        return floor
    bound, _ = oracle.evaluate_batch(ndim, q, np.abs(coef), mn, mx, nodes)
    return np.maximum(floor, (128 + 4 * int(np.sum(nodes))) * eps * np.abs(bound))


@pytest.mark.parametrize("ndim,nodes", CASES)
def test_splfe_matches_oracle(oracle, ndim, nodes):
    rng = np.random.default_rng(11 * ndim + nodes[0])
    ncol = int(np.prod(nodes))
    coef = rng.standard_normal(ncol)
    mn = -rng.random(ndim)
    mx = 1.0 + rng.random(ndim)
    q = _queries(rng, ndim, 4000, mn, mx)
    ref, ie = oracle.evaluate_batch(ndim, q, coef, mn, mx, nodes)
    got, ierr = sp.eval_batch(ndim, q, coef, mn, mx, nodes)
    assert ie == 0 and ierr == 0
    assert (np.abs(got - ref) <= _tol(coef, ndim, oracle, q, mn, mx, nodes)).all(), np.abs(got - ref).max()


@pytest.mark.parametrize("ndim,nodes", [(1, [10]), (2, [6, 7]), (3, [5, 4, 6]), (3, [24, 24, 24]), (4, [4, 5, 4, 6])])
def test_splde_all_derivative_orders(oracle, ndim, nodes):
    rng = np.random.default_rng(5 + ndim)
    ncol = int(np.prod(nodes))
    coef = rng.standard_normal(ncol)
    mn = np.zeros(ndim)
    mx = np.full(ndim, 2.0)
    q = _queries(rng, ndim, 1500, mn, mx)
    dxin = (np.array(nodes) - 1) / (mx - mn)
    for code in range(3 ** ndim):
        nd = [(code // 3 ** d) % 3 for d in range(ndim)]
        ref, _ = oracle.evaluate_batch(ndim, q, coef, mn, mx, nodes, nderiv=nd)
        got, ierr = sp.eval_batch(ndim, q, coef, mn, mx, nodes, nderiv=nd)
        assert ierr == 0
        scale = np.prod((3.0 * dxin) ** np.array(nd))     # |d^k Phi| <= (3 dxin)^k * O(value bound)
        tol = _tol(coef, ndim, oracle, q, mn, mx, nodes) * scale
        assert (np.abs(got - ref) <= tol).all(), (nd, np.abs(got - ref).max())


def test_scalar_entry_points_and_generic(oracle):
    """splfe / splde one point per call, and the splpak_type generic resolution."""
    rng = np.random.default_rng(2)
    nodes = [7, 5]
    coef = rng.standard_normal(35)
    s = sp.SplpakType(quiet=True)
    for _ in range(5):
        x = rng.random(2) * 1.4 - 0.2
        ref, _ = oracle.evaluate(2, x, coef, [0, 0], [1, 1], nodes)
        got, ierr = s.evaluate(2, x, coef, [0, 0], [1, 1], nodes)
        assert ierr == 0 and abs(got - ref) <= _tol(coef, 2)
        ref, _ = oracle.evaluate(2, x, coef, [0, 0], [1, 1], nodes, nderiv=[1, 2])
        got, ierr = s.evaluate(2, x, [1, 2], coef, [0, 0], [1, 1], nodes)
        assert ierr == 0 and abs(got - ref) <= _tol(coef, 2) * 6 * 16


def test_nderiv_out_of_range_sets_104_but_still_evaluates(oracle):
    """:1190-1194 sets 104 and does NOT return; the value is whatever bascmp's select-case gives."""
    rng = np.random.default_rng(8)
    coef = rng.standard_normal(12)
    q = rng.random((64, 1)) * 1.2 - 0.1
    for nd in ([3], [-1], [4]):
        ref, ie = oracle.evaluate_batch(1, q, coef, [0.0], [1.0], [12], nderiv=nd)
        got, ierr = sp.eval_batch(1, q, coef, [0.0], [1.0], [12], nderiv=nd)
        assert ie == 104 and ierr == 104
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9 * max(1.0, np.abs(ref).max()))


def test_large_table_goes_through_global_path(oracle):
    """A coefficient table larger than shared memory (2-D 200x200 = 320 KB) uses the L2 gather path."""
    rng = np.random.default_rng(4)
    nodes = [200, 200]
    coef = rng.standard_normal(40000)
    q = rng.random((3000, 2)) * 1.2 - 0.1
    ref, _ = oracle.evaluate_batch(2, q, coef, [0, 0], [1, 1], nodes)
    got, ierr = sp.eval_batch(2, q, coef, [0, 0], [1, 1], nodes)
    assert ierr == 0
    assert (np.abs(got - ref) <= _tol(coef, 2, oracle, q, [0, 0], [1, 1], nodes)).all()


def test_odd_ncol_and_strided_x(oracle):
    """ncol odd (bulk copy needs padding) and l1x > ndim."""
    rng = np.random.default_rng(6)
    nodes = [5, 7]
    coef = rng.standard_normal(35)
    q = np.zeros((500, 4))
    q[:, :2] = rng.random((500, 2))
    q[:, 2:] = 99.0
    ref, _ = oracle.evaluate_batch(2, q, coef, [0, 0], [1, 1], nodes)
    got, ierr = sp.eval_batch(2, q, coef, [0, 0], [1, 1], nodes)
    assert ierr == 0
    np.testing.assert_allclose(got, ref, rtol=0, atol=_tol(coef, 2))


def test_chunked_host_path_many_queries(oracle):
    """More queries than one staging chunk (2^22): exercises the double-buffered H2D/D2H pipeline."""
    rng = np.random.default_rng(9)
    nodes = [24, 24, 24]
    coef = rng.standard_normal(24 ** 3)
    nq = (1 << 22) * 2 + 12345
    q = rng.random((nq, 3))
    got, ierr = sp.eval_batch(3, q, coef, [0, 0, 0], [1, 1, 1], nodes)
    assert ierr == 0
    pick = rng.integers(0, nq, 3000)
    pick[:3] = [0, (1 << 22), nq - 1]
    ref, _ = oracle.evaluate_batch(3, q[pick], coef, [0, 0, 0], [1, 1, 1], nodes)
    np.testing.assert_allclose(got[pick], ref, rtol=0, atol=_tol(coef, 3))


def test_real32_library(oracle32):
    rng = np.random.default_rng(10)
    nodes = [8, 9]
    coef = rng.standard_normal(72).astype(np.float32)
    q = (rng.random((2000, 2)) * 1.2 - 0.1).astype(np.float32)
    ref, _ = oracle32.evaluate_batch(2, q, coef, [0, 0], [1, 1], nodes)
    got, ierr = sp.eval_batch(2, q, coef, [0, 0], [1, 1], nodes, real32=True)
    assert ierr == 0 and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-4 * np.abs(coef).max())


@pytest.mark.parametrize("ndim,nodes", [(1, [50]), (2, [8, 9]), (2, [64, 64]), (3, [24, 24, 24]), (3, [5, 4, 6]), (4, [5, 4, 6, 5])])
def test_real32_splfe_runs_in_working_precision(oracle32, ndim, nodes, basis_mode):
    """REAL32 library, splfe: float arithmetic like the reference built with -DREAL32 (src/splpak.F90:33-34).
    SPLPAK_B200_BASIS=exact: the 1-D basis values are bit-identical to the float oracle's (unfused float operations); only
    the summation order of the 4^ndim terms differs, so the tolerance is k * eps32 * sum |c_j phi_j| -- not round 1's fixed
    2e-4.  Default: the uniform form with float weights and float contraction; it does not repeat the float oracle's
    rounding of the node positions (eps32 * node index in u), hence the extra 4 eps32 sum(nodes) (see _tol)."""
    import os

    rng = np.random.default_rng(20 + ndim)
    coef = rng.standard_normal(int(np.prod(nodes))).astype(np.float32)
    mn, mx = [0.0] * ndim, [1.0] * ndim
    q = (rng.random((3000, ndim)) * 1.3 - 0.15).astype(np.float32)
    q[0] = 0.0
    q[1] = 1.0
    ref, _ = oracle32.evaluate_batch(ndim, q, coef, mn, mx, nodes)
    bound, _ = oracle32.evaluate_batch(ndim, q, np.abs(coef), mn, mx, nodes)
    got, ierr = sp.eval_batch(ndim, q, coef, mn, mx, nodes, real32=True)
    assert ierr == 0 and got.dtype == np.float32
    eps32 = float(np.finfo(np.float32).eps)
    # worst-case rounding of two different summation orders of 4^ndim float terms (sequential in the oracle, nested here)
    knod = 0 if basis_mode == "exact" else 4 * int(np.sum(nodes))
    tol = np.maximum(8 * eps32 * np.abs(coef).max(), (2 * 4 ** ndim + 8 + knod) * eps32 * np.abs(bound))
    assert (np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= tol).all(), np.abs(got - ref).max()
    # the float64-internal path of the same library (float I/O) agrees to float rounding of the result
    os.environ["SPLPAK_B200_R32"] = "f64"
    try:
        got64, ierr = sp.eval_batch(ndim, q, coef, mn, mx, nodes, real32=True)
    finally:
        del os.environ["SPLPAK_B200_R32"]
    assert ierr == 0
    # (the float64-internal path takes the uniform form of the basis for these batch sizes: it does not repeat the float
    # rounding of the node positions, eps32 * node index in u, hence the extra 4 eps32 sum(nodes) -- see _tol)
    tol64 = tol + 4 * int(np.sum(nodes)) * eps32 * np.abs(bound) + eps32 * np.abs(got64)
    assert (np.abs(got.astype(np.float64) - got64.astype(np.float64)) <= tol64).all()
    # large batch (dynamic scheduler, shared-memory table) == small batch results for the same points
    big = np.tile(q, (200, 1))
    gotb, ierr = sp.eval_batch(ndim, big, coef, mn, mx, nodes, real32=True)
    assert ierr == 0
    if ndim < 4:
        assert np.array_equal(gotb[: len(q)], got) and np.array_equal(gotb[-len(q):], got)
    else:
        # 4-D, >= 2^18 scattered queries: the order probe hands the batch to the float64 regrouping kernel (float I/O)
        assert (np.abs(gotb[: len(q)].astype(np.float64) - ref.astype(np.float64)) <= tol).all()
        assert np.array_equal(gotb[: len(q)], gotb[-len(q):])


@pytest.mark.parametrize("ndim,nodes,naxis,nderiv", [
    (1, [9], [37], None),
    (1, [12], [50], [2]),
    (2, [8, 6], [23, 17], None),
    (2, [7, 9], [15, 31], [1, 0]),
    (3, [6, 5, 7], [13, 9, 11], None),
    (3, [8, 8, 8], [10, 12, 9], [0, 1, 2]),
    (4, [5, 4, 6, 5], [7, 5, 6, 4], None),
    (4, [5, 5, 5, 5], [4, 6, 5, 7], [1, 0, 0, 1]),
])
def test_eval_grid_matches_pointwise_and_oracle(oracle, ndim, nodes, naxis, nderiv):
    """splpak_b200_eval_grid (separable mode products) == splfe/splde at every grid point: same 1-D basis values,
    different summation order only (tolerance 50 eps sum|c Phi| as for the point-wise path)."""
    rng = np.random.default_rng(100 + ndim)
    coef = rng.standard_normal(int(np.prod(nodes)))
    mn = [-0.5 + 0.1 * d for d in range(ndim)]
    mx = [1.0 + 0.2 * d for d in range(ndim)]
    axes = [np.sort(rng.random(n) * (mx[d] - mn[d]) * 1.4 + mn[d] - 0.2 * (mx[d] - mn[d])) for d, n in enumerate(naxis)]
    got, ierr = sp.eval_grid(ndim, axes, coef, mn, mx, nodes, nderiv=nderiv)
    assert ierr == 0 and got.shape == tuple(reversed(naxis))
    mesh = np.meshgrid(*axes, indexing="ij")                       # mesh[d][i1, ..., iN]
    pts = np.stack([m.transpose(*reversed(range(ndim))).ravel() for m in mesh], axis=1)   # dimension 1 fastest
    want, ierr = sp.eval_batch(ndim, pts, coef, mn, mx, nodes, nderiv=nderiv)
    assert ierr == 0
    scale = np.abs(coef).max() * 4 ** ndim * 6 ** ndim
    if nderiv is not None:
        for d in range(ndim):
            scale *= ((nodes[d] - 1) / (mx[d] - mn[d])) ** nderiv[d]
    tol = 50 * np.finfo(float).eps * scale
    np.testing.assert_allclose(got.ravel(), want, rtol=0, atol=tol)
    pick = rng.integers(0, len(pts), 300)
    ref, _ = oracle.evaluate_batch(ndim, pts[pick], coef, mn, mx, nodes, nderiv=nderiv)
    np.testing.assert_allclose(got.ravel()[pick], ref, rtol=0, atol=tol)


def test_eval_grid_large_device_and_slabs(oracle):
    """Device variant on a 600 x 500 x 400 grid (1.2e8 points) and the host variant's slab loop (> 32M outputs)."""
    import torch

    rng = np.random.default_rng(7)
    nodes = [24, 24, 24]
    coef = rng.standard_normal(24 ** 3)
    naxis = [600, 500, 400]
    axes = [np.linspace(-0.05, 1.05, n) for n in naxis]
    d_axes = torch.tensor(np.concatenate(axes), device="cuda")
    d_coef = torch.tensor(coef, device="cuda")
    d_out = torch.empty(int(np.prod(naxis)), dtype=torch.float64, device="cuda")
    assert sp.eval_grid_device(3, d_axes, naxis, d_coef, [0, 0, 0], [1, 1, 1], nodes, d_out) == 0
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().reshape(400, 500, 600)
    idx = rng.integers(0, [600, 500, 400], size=(2000, 3))
    pts = np.stack([axes[d][idx[:, d]] for d in range(3)], axis=1)
    ref, _ = oracle.evaluate_batch(3, pts, coef, [0, 0, 0], [1, 1, 1], nodes)
    tol = 50 * np.finfo(float).eps * 64 * 216 * np.abs(coef).max()
    np.testing.assert_allclose(out[idx[:, 2], idx[:, 1], idx[:, 0]], ref, rtol=0, atol=tol)
    host, ierr = sp.eval_grid(3, axes, coef, [0, 0, 0], [1, 1, 1], nodes)
    assert ierr == 0
    np.testing.assert_array_equal(host, out)
