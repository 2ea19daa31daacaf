"""Does the default mode actually differ from run to run where the deterministic mode does not?  (sanity check of
tests/test_gpu_fit.py::test_deterministic_mode_is_bit_reproducible) + the cost of the mode at cfg3 / cfg4 scale."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth
from util import make_problem

def run(ndim, nodes, x, y, w, mn, mx):
    h = sp.FitHandle(ndim, mn, mx, nodes, 1.0)
    assert h.add_points(x, y, w) == 0
    S, g, *_ = h.normal_equations()
    c, ierr = h.compute()
    h.destroy()
    return S.copy(), g.copy(), c.copy()

for ndim, nodes, n in ((1, [300], 200_000), (2, [30, 25], 300_000), (3, [9, 8, 10], 300_000), (4, [6, 5, 6, 5], 300_000)):
    x, y, w, mn, mx = make_problem(ndim, nodes, n, seed=70 + ndim, weighted=True, hole=True, outside=0.05)
    for det in ("0", "1"):
        os.environ["SPLPAK_B200_DETERMINISTIC"] = det
        r = [run(ndim, nodes, x, y, w, mn, mx) for _ in range(4)]
        dS = max(np.abs(q[0] - r[0][0]).max() for q in r[1:]) / np.abs(r[0][0]).max()
        dc = max(np.abs(q[2] - r[0][2]).max() for q in r[1:]) / np.abs(r[0][2]).max()
        print(f"{ndim}-D {nodes} det={det}: max run-to-run difference S {dS:.2e} (relative to max|S|), coef {dc:.2e}", flush=True)

def timed(name, ndim, nodes, n):
    x, y, w = synth.points_torch(ndim, n, seed=42, weighted=True)
    dcoef = torch.zeros(int(np.prod(nodes)), dtype=torch.float64, device="cuda")
    for det in ("0", "1"):
        os.environ["SPLPAK_B200_DETERMINISTIC"] = det
        h = sp.FitHandle(ndim, [0.0] * ndim, [1.0] * ndim, nodes, 1.0)
        st = torch.cuda.ExternalStream(h.stream())
        best = 1e30
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h.reset()
            e0.record(st)
            assert h.add_points_device(x, ndim, y, w, n, True) == 0
            assert h.compute_device(dcoef) == 0
            e1.record(st)
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        print(f"{name} det={det}: fit {best:.2f} ms", {k: round(v, 2) for k, v in h.timings().items()}, flush=True)
        h.destroy()

timed("cfg3", 3, [24, 24, 24], 100_000_000)
timed("cfg4", 4, [12] * 4, 10_000_000)
timed("cfg2", 2, [64, 64], 1_000_000)
