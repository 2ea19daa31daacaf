"""A/B of the fit's device-resident assembly with L2-resident chunks (SPLPAK_B200_L2CHUNK) at cfg3."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth
npts = 100_000_000
x, y, w = synth.points_torch(3, npts, start=0, seed=42, device="cuda")
d = torch.zeros(24 ** 3, dtype=torch.float64, device="cuda")
ref = None
for chunk in (0, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
    if chunk: os.environ["SPLPAK_B200_L2CHUNK"] = str(chunk)
    else: os.environ.pop("SPLPAK_B200_L2CHUNK", None)
    h = sp.FitHandle(3, [0.] * 3, [1.] * 3, [24] * 3, 1.0)
    st = torch.cuda.ExternalStream(h.stream())
    ts, ta = [], []
    with torch.cuda.stream(st):
        for rep in range(6):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            h.reset(); e0.record(st)
            h.add_points_device(x, 3, y, w, npts, True); e1.record(st)
            ie = h.compute_device(d); e2.record(st); torch.cuda.synchronize()
            if rep: ts.append(e0.elapsed_time(e2)); ta.append(e0.elapsed_time(e1))
    c = d.cpu().numpy().copy()
    if ref is None: ref = c
    print(f"L2CHUNK {chunk:9d} ierr {ie} fit best {min(ts):7.3f} ms median {statistics.median(ts):7.3f}  assembly best {min(ta):7.3f} ms   max|dc|/max|c| {np.abs(c - ref).max() / np.abs(ref).max():.2e}", flush=True)
    h.destroy()
