"""BASELINE configs[4]: batched evaluation sweep -- splfe/splde, 1-D..4-D, real64 and real32, uniform-random and
raster-ordered queries, 1e6..1e9 queries on one B200.  Prints a markdown table (device-resident, CUDA events,
best of 4 after one warm-up).   usage: python scripts/eval_sweep.py [max_nq]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth

HBM = 6534.5
try:
    HBM = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
GRIDS = {1: [50], 2: [64, 64], 3: [24, 24, 24], 4: [12, 12, 12, 12]}
max_nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000


def run(ndim, nq, real32, raster, nderiv):
    nodes = GRIDS[ndim]
    dt = torch.float32 if real32 else torch.float64
    ncol = 1
    for n in nodes:
        ncol *= n
    coef = torch.randn(ncol, dtype=dt, device="cuda")
    q = synth.queries_torch(ndim, nq, raster=raster, dtype=dt)
    out = torch.empty(nq, dtype=dt, device="cuda")
    torch.cuda.synchronize()
    best = 1e30
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ierr = sp.eval_batch_device(ndim, q, ndim, nq, coef, [0.0] * ndim, [1.0] * ndim, nodes, out, nderiv=nderiv,
                                    stream=torch.cuda.current_stream(), real32=real32)
        e1.record()
        torch.cuda.synchronize()
        assert ierr == 0
        if rep:
            best = min(best, e0.elapsed_time(e1))
    s = 4 if real32 else 8
    gbs = nq * (ndim + 1) * s / best / 1e6
    return best, nq / best / 1e6, gbs


print("| ndim | nodes | real | entry | order | queries | ms | Gq/s | algorithmic GB/s | of HBM (measured %.0f GB/s) |" % HBM)
print("|---|---|---|---|---|---:|---:|---:|---:|---:|")
for ndim in (1, 2, 3, 4):
    for real32 in (False, True):
        for nq in (1_000_000, 100_000_000, 1_000_000_000):
            if nq > max_nq or nq * (ndim + 1) * (4 if real32 else 8) > 120e9:
                continue
            for raster in (False, True):
                for nderiv in (None, [1] + [0] * (ndim - 1)):
                    if nderiv is not None and (nq != 100_000_000 or raster):
                        continue
                    ms, gq, gbs = run(ndim, nq, real32, raster, nderiv)
                    print(f"| {ndim} | {'x'.join(map(str, GRIDS[ndim]))} | {'real32' if real32 else 'real64'} | "
                          f"{'splfe' if nderiv is None else 'splde d/dx1'} | {'raster' if raster else 'random'} | {nq:.0e} | "
                          f"{ms:.3f} | {gq:.2f} | {gbs:.0f} | {100 * gbs / HBM:.1f} % |", flush=True)
