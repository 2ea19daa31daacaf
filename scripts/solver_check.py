"""Solver-only check at growing sizes: GPU compute() vs scipy.linalg.solveh_banded on the GPU's own
normal equations (xtrap = 0 so compute() adds nothing to G)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg as sl
import splpak_b200 as sp
from splpak_b200 import synth

def band_from_stencil(S, nodes):
    nodes = [int(v) for v in nodes]; ndim = len(nodes); n = int(np.prod(nodes))
    strides = np.cumprod([1] + nodes[:-1])
    bw = int(min(3 * strides.sum(), n - 1))
    ab = np.zeros((bw + 1, n))
    idx = np.arange(n); multi = []
    k = idx.copy()
    for d in range(ndim):
        multi.append(k % nodes[d]); k //= nodes[d]
    multi = np.stack(multi, 1)
    # enumerate stencil offsets with linear offset >= 0
    import itertools
    for delta in itertools.product(range(-3, 4), repeat=ndim):
        off = int(sum(dd * st for dd, st in zip(delta, strides)))
        if off < 0 or off > bw: continue
        # column j, row i = j + off, i_d = j_d + delta_d
        jd = multi; idd = multi + np.array(delta)
        ok = ((idd >= 0) & (idd < np.array(nodes))).all(1)
        mn = np.minimum(jd, idd); node = (mn * strides).sum(1)
        sten = int(sum(abs(dd) * 4 ** d for d, dd in enumerate(delta)))
        ab[off, idx[ok]] += S[node[ok], sten]
    return ab, bw

import sys
cases=[(3,[7,7,7],100000),(3,[8,8,8],100000),(3,[9,9,9],100000),(3,[10,10,10],200000),(3,[10,10,10],20000),(2,[30,30],100000),(1,[200],50000),(1,[1000],50000),(3,[16,16,16],600000)]
if len(sys.argv)>1: cases=[(3,[10,10,10],20000)]
for ndim, nodes, n in cases:
    x,y,w = synth.points_numpy(ndim, n)
    h = sp.FitHandle(ndim,[0]*ndim,[1]*ndim,nodes,0.0)
    h.add_points(x,y,w)
    S,g,cnt,tot,nrows = h.normal_equations()
    coef, ierr = h.compute()
    ab, bw = band_from_stencil(S, nodes)
    ref = sl.solveh_banded(ab, g, lower=True)
    print(ndim, nodes, "ncol", len(g), "bw", bw, "ierr", ierr, "rel err", np.abs(coef-ref).max()/np.abs(ref).max(), flush=True)
    h.destroy()
