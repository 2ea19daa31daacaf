"""cfg3-scale stage timings (device resident): fit stages + eval, a few repetitions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth
ndim, nodes = 3, [24,24,24]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
x,y,w = synth.points_torch(ndim,n)
h = sp.FitHandle(ndim,[0]*3,[1]*3,nodes,1.0)
dcoef = torch.zeros(24**3, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(4):
    h.reset()
    t0=time.time()
    rc = h.add_points_device(x,3,y,w,n,True)
    ierr = h.compute_device(dcoef)
    torch.cuda.synchronize()
    t1=time.time()
    t = h.timings()
    print("fit n",n,"rc",rc,"ierr",ierr,"wall ms %.2f"%((t1-t0)*1e3), {k: round(v,3) for k,v in t.items()}, "sum %.2f"%sum(t.values()), flush=True)
q = synth.queries_torch(ndim, 1000)
out = torch.zeros(1000, dtype=torch.float64, device="cuda")
sp.eval_batch_device(ndim,q,3,1000,dcoef,[0]*3,[1]*3,nodes,out)
torch.cuda.synchronize()
print("fit-vs-truth max abs", float((out-synth._smooth(q, torch)).abs().max()))
del x,y,w
for raster in (False, True):
    q = synth.queries_torch(ndim,nq,raster=raster)
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ts=[]
    for rep in range(8):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        ierr = sp.eval_batch_device(ndim,q,3,nq,dcoef,[0]*3,[1]*3,nodes,out,stream=torch.cuda.current_stream())
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1),2))
    print("eval nq",nq,"raster",raster,"ierr",ierr,"ms",ts,"best Gq/s %.2f"%(nq/min(ts)/1e6),"GB/s %.0f"%(nq*32/min(ts)/1e6), flush=True)
    del q,out
