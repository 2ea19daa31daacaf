"""Device-resident timings of BASELINE.json's four fit configs on one B200 (CUDA events around the whole fit /
evaluation call sequence, best of 3 after one warm-up), with the per-stage split of the fit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth

def run(name, ndim, nodes, n, nq, weighted, xtrap, hole=False, nderiv=None):
    x, y, w = synth.points_torch(ndim, int(n * (1.1 if hole else 1.0)), seed=42, weighted=weighted)
    if hole:
        keep = ((x - 0.5) ** 2).sum(dim=1) > 0.15 ** 2
        x, y = x[keep][:n].contiguous(), y[keep][:n].contiguous()
        if w is not None:
            w = w[keep][:n].contiguous()
    ncol = int(np.prod(nodes))
    dcoef = torch.zeros(ncol, dtype=torch.float64, device="cuda")
    q = synth.queries_torch(ndim, nq)
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    h = sp.FitHandle(ndim, [0.0] * ndim, [1.0] * ndim, nodes, xtrap)
    st = torch.cuda.ExternalStream(h.stream())
    best_fit, best_ref, best_ev, stages, fired = 1e30, 1e30, 1e30, None, False
    for rep in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        h.reset()
        e[0].record(st)
        assert h.add_points_device(x, ndim, y, w, n, weighted) == 0
        assert h.compute_device(dcoef) == 0
        e[1].record(st)
        fired = h.constraints_fired()
        if fired:
            assert h.refine_device(x, ndim, y, w, n, dcoef, weighted=weighted) == 0
        e[2].record(st)
        assert sp.eval_batch_device(ndim, q, ndim, nq, dcoef, [0.0] * ndim, [1.0] * ndim, nodes, out, nderiv=nderiv, stream=st) == 0
        e[3].record(st)
        torch.cuda.synchronize()
        if rep:
            best_fit = min(best_fit, e[0].elapsed_time(e[1]))
            best_ref = min(best_ref, e[1].elapsed_time(e[2]))
            best_ev = min(best_ev, e[2].elapsed_time(e[3]))
            stages = h.timings()
    h.destroy()
    print(f"| {name} | {ndim} | {'x'.join(map(str, nodes))} | {n:.0e} | {best_fit:.2f} | {n / best_fit / 1e6:.3f} | "
          f"{'%.2f' % best_ref if fired else '-'} | {nq:.0e} | {best_ev:.3f} | {nq / best_ev / 1e6:.2f} | "
          + ", ".join(f"{k} {v:.2f}" for k, v in stages.items()) + " |", flush=True)

print("| config | ndim | nodes | points | fit ms | Gpoints/s | refinement step ms | queries | eval ms | Gq/s | fit stages (ms, incl. refinement pass) |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---|")
only = sys.argv[1:]          # e.g. "cfg4": just that row
_run = run
def run(name, *a, **k):
    if not only or any(name.startswith(o) for o in only):
        _run(name, *a, **k)
run("cfg1 splcw", 1, [50], 10_000, 100_000, True, 1.0)
run("cfg2 splcc + hole, splde d/dx", 2, [64, 64], 1_000_000, 10_000_000, False, 1.0, hole=True, nderiv=[1, 0])
run("cfg3 splcw", 3, [24, 24, 24], 100_000_000, 1_000_000_000, True, 1.0)
run("cfg4 splcw", 4, [12, 12, 12, 12], 10_000_000, 100_000_000, True, 1.0)
