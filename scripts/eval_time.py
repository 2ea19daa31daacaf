"""Evaluation-only timing at cfg3 scale: random and raster order, a few repetitions.
usage: python scripts/eval_time.py [nq] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth
ndim, nodes = 3, [24, 24, 24]
nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
coef = torch.randn(24 ** 3, dtype=torch.float64, device="cuda")
for raster in (False, True):
    q = synth.queries_torch(ndim, nq, raster=raster)
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ts = []
    for rep in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ierr = sp.eval_batch_device(ndim, q, 3, nq, coef, [0] * 3, [1] * 3, nodes, out, stream=torch.cuda.current_stream())
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    print("eval nq", nq, "raster", raster, "ierr", ierr, "ms", ts, "best Gq/s %.2f" % (nq / min(ts) / 1e6),
          "GB/s %.0f" % (nq * 32 / min(ts) / 1e6), "checksum", float(out[::max(1, nq // 1000)].sum()), flush=True)
    del q, out
