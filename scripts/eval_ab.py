"""A/B timing of the evaluation kernels: plain vs regrouping, random and raster order, 2-D/3-D/4-D.
usage: python scripts/eval_ab.py [nq] [reps]   (device-resident, CUDA events, best and median of reps)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth
GRIDS = {1: [50], 2: [64, 64], 3: [24, 24, 24], 4: [12, 12, 12, 12]}
nq0 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dims = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [3, 2, 4]
for ndim in dims:
    nodes = GRIDS[ndim]
    nq = nq0 if ndim < 4 else min(nq0, 500_000_000)
    ncol = 1
    for n in nodes:
        ncol *= n
    coef = torch.randn(ncol, dtype=torch.float64, device="cuda")
    for raster in (False, True):
        q = synth.queries_torch(ndim, nq, raster=raster)
        out = torch.empty(nq, dtype=torch.float64, device="cuda")
        ref = None
        for mode in ("plain", "regroup"):
            os.environ["SPLPAK_B200_EVAL"] = mode
            ts = []
            for rep in range(reps + 1):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                ierr = sp.eval_batch_device(ndim, q, ndim, nq, coef, [0.0] * ndim, [1.0] * ndim, nodes, out,
                                            stream=torch.cuda.current_stream())
                e1.record(); torch.cuda.synchronize()
                if rep:
                    ts.append(e0.elapsed_time(e1))
            chk = out[:: max(1, nq // 100000)].clone()
            same = True if ref is None else bool(torch.equal(chk, ref))
            ref = chk if ref is None else ref
            print(f"ndim {ndim} nq {nq:.0e} {'raster' if raster else 'random'} {mode:8s} ierr {ierr} "
                  f"best {min(ts):8.3f} ms median {statistics.median(ts):8.3f} ms  {nq / min(ts) / 1e6:7.2f} Gq/s  "
                  f"{nq * (ndim + 1) * 8 / min(ts) / 1e6:6.0f} GB/s  same_as_plain {same}", flush=True)
        del q, out
os.environ.pop("SPLPAK_B200_EVAL", None)
