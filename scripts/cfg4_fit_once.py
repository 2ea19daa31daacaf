"""One cfg4-shaped fit (12^4 nodes, half bandwidth 5,655; few points: the solve does not depend on them) -- the command
the ncu launch list / captures of the kernel-per-phase factor loop are taken from."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth
ndim, nodes = 4, [12] * 4
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x, y, w = synth.points_torch(ndim, n, seed=42, weighted=True)
h = sp.FitHandle(ndim, [0.0] * ndim, [1.0] * ndim, nodes, 1.0)
dcoef = torch.zeros(12 ** 4, dtype=torch.float64, device="cuda")
for rep in range(reps):
    h.reset()
    assert h.add_points_device(x, ndim, y, w, n, True) == 0
    ierr = h.compute_device(dcoef)
    torch.cuda.synchronize()
    print("ierr", ierr, {k: round(v, 3) for k, v in h.timings().items()}, flush=True)
