import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import splpak_b200 as sp
from oracle import Oracle
o = Oracle()
def ev(mode, ndim, q, coef, nodes):
    os.environ["SPLPAK_B200_EVAL"] = mode
    out, ie = sp.eval_batch(ndim, q, coef, [0.0]*ndim, [1.0]*ndim, nodes)
    return out
rng = np.random.default_rng(0)
for ndim, nodes, nq in [(3, [5, 4, 6], 40), (3, [24, 24, 24], 5000), (2, [9, 33], 5000), (4, [4, 5, 4, 6], 3000), (3,[16,16,16],5000), (3,[8,8,8],5000)]:
    ncol = int(np.prod(nodes))
    coef = rng.standard_normal(ncol)
    q = rng.random((nq, ndim))
    a = ev("plain", ndim, q, coef, nodes); b = ev("regroup", ndim, q, coef, nodes)
    ref, _ = o.evaluate_batch(ndim, q, coef, [0.0]*ndim, [1.0]*ndim, nodes)
    print(ndim, nodes, nq, "plain-vs-oracle", np.abs(a - ref).max(), "regroup-vs-oracle", np.abs(b - ref).max(),
          "permutation?", np.allclose(np.sort(a), np.sort(b)), "n differ", int((a != b).sum()))
    if ndim == 3 and nq == 40:
        print(np.c_[a, b, ref][:12])
    # linear coef: result should reproduce a multilinear function -> tells whether the table or the weights are off
    idx = np.arange(ncol)
    coef1 = np.ones(ncol)
    a1 = ev("plain", ndim, q, coef1, nodes); b1 = ev("regroup", ndim, q, coef1, nodes)
    print("   ones-table: plain", a1[:4], "regroup", b1[:4])
    # delta tables: which coefficient does each kernel read?
    if ncol <= 200:
        bad = 0
        for j in range(ncol):
            cj = np.zeros(ncol); cj[j] = 1.0
            aj = ev("plain", ndim, q, cj, nodes); bj = ev("regroup", ndim, q, cj, nodes)
            if not np.array_equal(aj, bj):
                bad += 1
                if bad <= 5:
                    print("   delta", j, "differs at", int((aj != bj).sum()), "queries; first", aj[aj != bj][:3], bj[aj != bj][:3])
        print("   delta tables differing:", bad, "of", ncol)
