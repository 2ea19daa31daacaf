"""Calibration of the pivot-ratio bound (FitHandle.condition_estimate) against the true cond(G), and the error of the
plain Cholesky solve against the oracle -- to choose the thresholds of the automatic refinement / orthogonal fallback."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import splpak_b200 as sp
from oracle import Oracle
from util import make_problem
o = Oracle()
rng = np.random.default_rng(0)
print("ndim nodes xtrap kind | cond(A) cond(G) est est/cond(G) | err_plain err_refined")
for ndim, nodes, n in [(1, [12], 300), (2, [8, 8], 2000), (2, [12, 12], 6000), (3, [6, 6, 6], 6000), (3, [8, 7, 9], 20000)]:
    for kind in ("uniform", "hole", "clustered"):
        for xtrap in (0.0, 1.0):
            x, y, w, mn, mx = make_problem(ndim, nodes, n, seed=9, hole=(kind == "hole"))
            if kind == "clustered":
                x = mn + (mx - mn) * (0.5 + 0.5 * np.tanh(4 * (x - 0.5) / (mx - mn)))   # squeeze towards the centre... keeps coverage thin at the edges
                x = np.clip(x, mn, mx)
            ref, ie = o.initialize(ndim, x, y, w, mn, mx, nodes, xtrap)
            if ie != 0:
                continue
            A, _ = o.rows(ndim, x, y, w, mn, mx, nodes, xtrap)
            ca = np.linalg.cond(A)
            h = sp.FitHandle(ndim, mn, mx, nodes, xtrap, solver="cholesky")
            h.add_points(x, y, w)
            c0, ierr = h.compute()
            est = h.condition_estimate()
            e0 = np.abs(c0 - ref).max() / np.abs(ref).max() if ierr == 0 else float('nan')
            e1 = float('nan')
            if ierr == 0:
                c1, ie1 = h.refine(x, y, w, steps=2)
                e1 = np.abs(c1 - ref).max() / np.abs(ref).max() if ie1 == 0 else float('nan')
            h.destroy()
            print(f"{ndim} {nodes} {xtrap} {kind:9s} | {ca:.2e} {ca*ca:.2e} {est:.2e} {est/(ca*ca):.2e} | ierr {ierr} {e0:.2e} {e1:.2e}", flush=True)
# scalar-call latency (the reference's usage: evaluate once per point in a loop, test/splpak_test.f90:72-80)
nodes = [24, 24, 24]
coef = rng.standard_normal(24 ** 3)
s = sp.SplpakType(quiet=True)
pts = rng.random((2000, 3))
for _ in range(50):
    s.evaluate(3, pts[0], coef, [0, 0, 0], [1, 1, 1], nodes)
t0 = time.perf_counter()
for p in pts:
    s.evaluate(3, p, coef, [0, 0, 0], [1, 1, 1], nodes)
t1 = time.perf_counter()
print(f"scalar splfe through the C ABI (python ctypes caller, 24^3 table): {1e6 * (t1 - t0) / len(pts):.1f} us per call")
t0 = time.perf_counter()
for p in pts[:500]:
    s.evaluate(3, p, coef + 1e-9 * p[0], [0, 0, 0], [1, 1, 1], nodes)          # a new table every call: upload each time
t1 = time.perf_counter()
print(f"   ... with a different coefficient table per call: {1e6 * (t1 - t0) / 500:.1f} us per call")
t0 = time.perf_counter()
for p in pts:
    o.evaluate(3, p, coef, [0, 0, 0], [1, 1, 1], nodes)
t1 = time.perf_counter()
print(f"oracle (CPU restatement) scalar splfe: {1e6 * (t1 - t0) / len(pts):.1f} us per call")
