import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import splpak_b200 as sp
from oracle import Oracle
from util import make_problem
o = Oracle()
ndim, nodes, n, xtrap = 2, [10, 10], 1500, 1.0
x, y, w, mn, mx = make_problem(ndim, nodes, n, seed=3, hole=True)
ref, ie = o.initialize(ndim, x, y, w, mn, mx, nodes, xtrap)
for trial in range(3):
    h = sp.FitHandle(ndim, mn, mx, nodes, xtrap, solver="orthogonal")
    h.add_points(x, y, w)
    Rw0 = h.orthogonal_factor(0)
    c, ierr = h.compute()
    Rw = h.orthogonal_factor(0); Rb = h.orthogonal_factor(1)
    h.destroy()
    print('trial', trial, 'ierr', ierr, 'Rw0 nan', np.isnan(Rw0).sum(), 'Rw nan', np.isnan(Rw).sum(), 'windows with nan', np.flatnonzero(np.isnan(Rw).any(axis=(1, 2)))[:20],
          'Rb nan rows', np.flatnonzero(np.isnan(Rb).any(axis=1))[:20], 'coef nan', np.isnan(c).sum(), flush=True)
    if np.isnan(Rw).any():
        wbad = np.flatnonzero(np.isnan(Rw).any(axis=(1, 2)))[0]
        print('window', wbad, 'before constraints:\n', np.array2string(Rw0[wbad][:6, :8], precision=3))
        print('after:\n', np.array2string(Rw[wbad][:6, :8], precision=3))
    else:
        A, r = o.rows(ndim, x, y, w, mn, mx, nodes, xtrap)
        G = A.T @ A
        # R^T R == G ?
        ncol = len(c); bw = Rb.shape[1] - 2
        R = np.zeros((ncol, ncol))
        for i in range(ncol):
            k = min(bw, ncol - 1 - i)
            R[i, i:i + k + 1] = Rb[i, :k + 1]
        print('R^T R vs G', np.abs(R.T @ R - G).max() / np.abs(G).max(), 'err', np.abs(c - ref).max() / np.abs(ref).max())
