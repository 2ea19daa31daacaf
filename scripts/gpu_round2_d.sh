#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_regroup.py tests/test_gpu_eval.py -x -q -m gpu 2>&1 | tail -4
python - <<'PY'
import os, sys, statistics, torch
sys.path.insert(0, '.')
import splpak_b200 as sp
from splpak_b200 import synth
nodes=[24,24,24]; nq=1_000_000_000
coef=torch.randn(24**3,dtype=torch.float64,device='cuda')
os.environ.pop('SPLPAK_B200_EVAL',None)
for raster in (False, True):
    q=synth.queries_torch(3,nq,raster=raster); out=torch.empty(nq,dtype=torch.float64,device='cuda'); ts=[]
    for rep in range(5):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
        sp.eval_batch_device(3,q,3,nq,coef,[0.]*3,[1.]*3,nodes,out,stream=torch.cuda.current_stream()); e1.record(); torch.cuda.synchronize()
        if rep: ts.append(e0.elapsed_time(e1))
    print('auto dispatch raster',raster,'best %.3f ms'%min(ts),'median %.3f'%statistics.median(ts),flush=True)
    del q,out
PY
