"""Timing of the orthogonal (Householder) fit path against the Cholesky path at cfg2 / cfg3 sizes (device-resident)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth
for ndim, nodes, npts in [(2, [64, 64], 1_000_000), (3, [24, 24, 24], 10_000_000)]:
    x, y, w = synth.points_torch(ndim, npts, start=0, seed=42, device="cuda")
    ncol = int(np.prod(nodes))
    res = {}
    for solver in ("cholesky", "orthogonal"):
        h = sp.FitHandle(ndim, [0.0] * ndim, [1.0] * ndim, nodes, 1.0, solver=solver)
        d = torch.zeros(ncol, dtype=torch.float64, device="cuda")
        ts = []
        for rep in range(3 if solver == "cholesky" else 2):
            h.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
            h.add_points_device(x, ndim, y, w, npts, True); h.synchronize(); t1 = time.perf_counter()
            ie = h.compute_device(d); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
        res[solver] = d.cpu().numpy()
        h.destroy()
        print(f"ndim {ndim} nodes {nodes} npts {npts:.0e} {solver:10s} ierr {ie} add_points {ts[-1][0]:9.2f} ms  compute {ts[-1][1]:9.2f} ms", flush=True)
    print("   max |c_orth - c_chol| / max|c| =", np.abs(res['orthogonal'] - res['cholesky']).max() / np.abs(res['cholesky']).max(), flush=True)
    del x, y, w
