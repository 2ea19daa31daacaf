"""First-contact GPU diagnostics: peaks, small parity numbers per stage, cfg3-scale timings.
Writes gpurun_out/diag.log.  Not a test; the parity gates live in tests/."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import splpak_b200 as sp
from splpak_b200 import synth
from oracle import Oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import make_problem, dense_from_stencil

def section(name):
    print("\n=== " + name, flush=True)

o = Oracle()
section("device")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count)
try:
    section("peaks")
    print("DFMA TF/s, DMMA TF/s, copy GB/s:", sp.measure_peaks())
except Exception:
    traceback.print_exc()

section("eval parity")
for ndim, nodes in [(1,[10]),(2,[6,7]),(3,[5,4,6]),(3,[24,24,24]),(4,[4,5,4,6]),(4,[12,12,12,12])]:
    try:
        rng = np.random.default_rng(ndim)
        coef = rng.standard_normal(int(np.prod(nodes)))
        q = rng.random((3000, ndim))*1.5-0.25
        ref,_ = o.evaluate_batch(ndim,q,coef,[0]*ndim,[1]*ndim,nodes)
        got,ierr = sp.eval_batch(ndim,q,coef,[0]*ndim,[1]*ndim,nodes)
        print(ndim,nodes,"ierr",ierr,"maxerr",np.abs(got-ref).max(), "maxref", np.abs(ref).max(), flush=True)
    except Exception:
        traceback.print_exc()

section("fit parity")
for ndim,nodes,nd,xtrap,hole in [(1,[10],200,1.0,False),(2,[6,7],2000,1.0,False),(2,[9,8],3000,1.0,True),(3,[5,4,6],4000,1.0,False),(3,[6,6,6],5000,1.0,True),(4,[4,5,4,4],6000,1.0,False)]:
    try:
        x,y,w,mn,mx = make_problem(ndim,nodes,nd,seed=ndim,hole=hole)
        h = sp.FitHandle(ndim,mn,mx,nodes,xtrap)
        rc = h.add_points(x,y,w)
        S,g,cnt,tot,nrows = h.normal_equations()
        A,r = o.rows(ndim,x,y,w,mn,mx,nodes,0.0)
        G = dense_from_stencil(S,nodes)
        Gref = A.T@A; gref = A.T@r
        print(ndim,nodes,"add rc",rc,"G err",np.abs(G-Gref).max()/np.abs(Gref).max(),"g err",np.abs(g-gref).max()/np.abs(gref).max(),"nrows",nrows,A.shape[0],"tot",tot,np.sum(w), flush=True)
        coef,ierr = h.compute()
        ref,ie = o.initialize(ndim,x,y,w,mn,mx,nodes,xtrap)
        A2,r2 = o.rows(ndim,x,y,w,mn,mx,nodes,xtrap)
        print("   compute ierr",ierr,ie,"coef relerr",np.abs(coef-ref).max()/np.abs(ref).max(),"rows",A2.shape[0],"cond",np.linalg.cond(A2.T@A2), h.timings(), flush=True)
        h.destroy()
    except Exception:
        traceback.print_exc()

section("cfg3-scale timings (device resident)")
try:
    ndim, nodes = 3, [24,24,24]
    for n in (10_000_000, 100_000_000):
        x,y,w = synth.points_torch(ndim,n)
        torch.cuda.synchronize()
        h = sp.FitHandle(ndim,[0]*3,[1]*3,nodes,1.0)
        dcoef = torch.zeros(24**3, dtype=torch.float64, device="cuda")
        for rep in range(3):
            h.reset()
            t0=time.time()
            rc = h.add_points_device(x,3,y,w,n,True)
            ierr = h.compute_device(dcoef)
            torch.cuda.synchronize()
            t1=time.time()
            print("n",n,"rep",rep,"rc",rc,"ierr",ierr,"wall ms",(t1-t0)*1e3,h.timings(),"launches",h.launch_count(), flush=True)
        # quality: compare fitted values with the noise-free function at a few points
        q = synth.queries_torch(ndim, 1000)
        out = torch.zeros(1000, dtype=torch.float64, device="cuda")
        sp.eval_batch_device(ndim,q,3,1000,dcoef,[0]*3,[1]*3,nodes,out)
        torch.cuda.synchronize()
        truth = synth._smooth(q, torch)
        print("   fit-vs-truth max abs", float((out-truth).abs().max()))
        h.destroy()
        del x,y,w
    section("eval timings")
    for nq in (100_000_000, 1_000_000_000):
        for raster in (False, True):
            q = synth.queries_torch(ndim,nq,raster=raster)
            out = torch.empty(nq, dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            for rep in range(3):
                e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
                e0.record()
                ierr = sp.eval_batch_device(ndim,q,3,nq,dcoef,[0]*3,[1]*3,nodes,out,stream=torch.cuda.current_stream())
                e1.record(); torch.cuda.synchronize()
                ms=e0.elapsed_time(e1)
                print("nq",nq,"raster",raster,"ierr",ierr,"ms",ms,"Gq/s",nq/ms/1e6,"GB/s",nq*32/ms/1e6, flush=True)
            del q,out
except Exception:
    traceback.print_exc()
