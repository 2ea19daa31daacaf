"""Why does the eval kernel time vary?  Repeats the same launch and prints every duration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import splpak_b200 as sp
from splpak_b200 import synth
ndim, nodes = 3, [24,24,24]
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
coef = torch.randn(24**3, dtype=torch.float64, device="cuda")
q = synth.queries_torch(ndim, nq)
out = torch.empty(nq, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
def run(tag, stream=None, n=8):
    ts=[]
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        s = stream or torch.cuda.current_stream()
        e0.record(s)
        sp.eval_batch_device(ndim,q,3,nq,coef,[0]*3,[1]*3,nodes,out,stream=s)
        e1.record(s); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1),2))
    print(tag, ts, flush=True)
run("default stream")
h = sp.FitHandle(ndim,[0]*3,[1]*3,nodes,1.0)
ext = torch.cuda.ExternalStream(h.stream())
run("handle stream", ext)
# after a fit (band storage etc. allocated)
x,y,w = synth.points_torch(ndim, 100_000_000)
dcoef = torch.zeros(24**3, dtype=torch.float64, device="cuda")
h.add_points_device(x,3,y,w,100_000_000,True); print("fit ierr", h.compute_device(dcoef))
run("after fit, default stream")
run("after fit, handle stream", ext)
coef = dcoef
run("fitted coef, default stream")
q2 = synth.queries_torch(ndim, nq, raster=True)
q = q2
run("raster")
