"""A/B of the two 3-D assembly paths (cell moments vs direct orthant-stencil accumulation):
normal equations and coefficients on a small case with exterior points, then cfg3 timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth

def run(mode, nodes, x, y, w, xtrap=1.0):
    os.environ["SPLPAK_B200_ASSEMBLY"] = mode
    h = sp.FitHandle(3, [0] * 3, [1] * 3, nodes, xtrap)
    rc = h.add_points(x, y, w, True)
    S, g, cnt, tot, rows = h.normal_equations()
    coef, ierr = h.compute()
    h.destroy() if hasattr(h, "destroy") else None
    return rc, S, g, coef, ierr

rng = np.random.default_rng(5)
for nodes, n, lo, hi in (([8, 7, 9], 200_000, -0.15, 1.15), ([5, 4, 6], 50_000, -1.0, 2.0), ([24] * 3, 2_000_000, 0.0, 1.0)):
    x = rng.uniform(lo, hi, (n, 3))
    x[:8] = [[0, 0, 0], [1, 1, 1], [0, 1, 0.5], [1, 0, 0.25], [0.5, 0.5, 0.5], [1, 1, 0], [0, 0, 1], [0.999999, 1e-9, 1]]
    y = np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1]) + x[:, 2] ** 2
    w = rng.uniform(0.5, 2.0, n)
    w[::17] = 0.0
    a = run("direct", nodes, x, y, w)
    b = run("moments", nodes, x, y, w)
    sS = np.abs(a[1]).max(); sg = np.abs(a[2]).max()
    print("nodes", nodes, "rc", a[0], b[0], "ierr", a[4], b[4],
          "S rel diff %.3e" % (np.abs(a[1] - b[1]).max() / sS), "g rel diff %.3e" % (np.abs(a[2] - b[2]).max() / sg),
          "coef rel diff %.3e" % (np.abs(a[3] - b[3]).max() / np.abs(a[3]).max()), flush=True)

if len(sys.argv) > 1:
    n = int(float(sys.argv[1]))
    nodes = [24] * 3
    x, y, w = synth.points_torch(3, n)
    dcoef = torch.zeros(24 ** 3, dtype=torch.float64, device="cuda")
    ref = None
    for mode in ("direct", "moments"):
        os.environ["SPLPAK_B200_ASSEMBLY"] = mode
        h = sp.FitHandle(3, [0] * 3, [1] * 3, nodes, 1.0)
        torch.cuda.synchronize()
        for rep in range(4):
            h.reset()
            rc = h.add_points_device(x, 3, y, w, n, True)
            ierr = h.compute_device(dcoef)
            torch.cuda.synchronize()
            t = h.timings()
            print(mode, "rc", rc, "ierr", ierr, {k: round(v, 3) for k, v in t.items()}, "sum %.2f" % sum(t.values()), flush=True)
        c = dcoef.cpu().numpy().copy()
        if ref is None: ref = c
        else: print("coef rel diff %.3e" % (np.abs(c - ref).max() / np.abs(ref).max()))
