"""Coefficient error against the oracle (the reference's QR numerics) with and without CSNE refinement,
on fits where derivative-constraint rows fire."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import splpak_b200 as sp
from oracle import Oracle
from util import make_problem
o = Oracle()
for ndim, nodes, n, seed in ((1, [30], 400, 1), (2, [12, 10], 3000, 2), (2, [20, 20], 20000, 3), (3, [8, 7, 8], 20000, 4), (3, [10, 10, 10], 60000, 5)):
    x, y, w, mn, mx = make_problem(ndim, nodes, n, seed=seed, weighted=True, hole=True)
    ref, ie = o.initialize(ndim, x, y, w, mn, mx, nodes, 1.0)
    A, r = o.rows(ndim, x, y, w, mn, mx, nodes, 1.0)
    condA = np.linalg.cond(A)
    h = sp.FitHandle(ndim, mn, mx, nodes, 1.0)
    h.add_points(x, y, w)
    c0, ierr = h.compute()
    fired = h.constraints_fired()
    errs = [np.abs(c0 - ref).max() / np.abs(ref).max()]
    for step in range(3):
        c, ierr = h.refine(x, y, w)
        errs.append(np.abs(c - ref).max() / np.abs(ref).max())
    h.destroy()
    one, ierr = sp.splcw(ndim, x, x.shape[1], y, w, len(x), mn, mx, nodes, 1.0, quiet=True)
    print(ndim, nodes, "rows", A.shape, "fired", fired, "cond(A) %.2e cond^2 %.2e" % (condA, condA ** 2),
          "err: plain %.2e, refine x1 %.2e x2 %.2e x3 %.2e; splcw %.2e" % (*errs, np.abs(one - ref).max() / np.abs(ref).max()), flush=True)
