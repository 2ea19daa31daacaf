"""Clock/power trace while repeating the same eval launch (random and raster order)."""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splpak_b200 as sp
from splpak_b200 import synth
ndim, nodes = 3, [24,24,24]
nq = 1_000_000_000
coef = torch.randn(24**3, dtype=torch.float64, device="cuda")
out = torch.empty(nq, dtype=torch.float64, device="cuda")
lines=[]
p = subprocess.Popen(["nvidia-smi","--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown,pstate","--format=csv,noheader","-lms","20","-i","0"],stdout=subprocess.PIPE,text=True)
def rd():
    for l in p.stdout: lines.append((time.time(), l.strip()))
threading.Thread(target=rd,daemon=True).start()
def run(tag, q, stream, n=12):
    for i in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        t0=time.time()
        e0.record(stream)
        sp.eval_batch_device(ndim,q,3,nq,coef,[0]*3,[1]*3,nodes,out,stream=stream)
        e1.record(stream); torch.cuda.synchronize()
        t1=time.time()
        smp=[l for (t,l) in lines if t0<=t<=t1]
        print(tag, i, round(e0.elapsed_time(e1),1), "ms |", " || ".join(s.split(",",1)[1] for s in smp[:6]), flush=True)
q = synth.queries_torch(ndim, nq)
torch.cuda.synchronize()
run("random/default", q, torch.cuda.current_stream())
s2 = torch.cuda.Stream()
run("random/torch-stream", q, s2)
h = sp.FitHandle(ndim,[0]*3,[1]*3,nodes,1.0)
run("random/handle-stream", q, torch.cuda.ExternalStream(h.stream()))
q = synth.queries_torch(ndim, nq, raster=True)
torch.cuda.synchronize()
run("raster/default", q, torch.cuda.current_stream())
p.terminate()
