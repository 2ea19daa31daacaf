"""Repeat the host-array evaluation call (pinned buffers) and print per-call wall times."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
from splpak_b200 import synth
nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
lib = sp.load()
q = synth.queries_torch(3, nq)
hq = torch.empty((nq, 3), dtype=torch.float64, pin_memory=True); hq.copy_(q); torch.cuda.synchronize(); del q
hout = torch.empty(nq, dtype=torch.float64, pin_memory=True)
coef = np.random.default_rng(0).standard_normal(24 ** 3)
mn = (C.c_double * 3)(0, 0, 0); mx = (C.c_double * 3)(1, 1, 1); no = (C.c_int * 3)(24, 24, 24)
ie = C.c_int(0)
for rep in range(8):
    t0 = time.perf_counter()
    lib.splpak_b200_eval(3, C.c_void_p(hq.data_ptr()), 3, nq, None, C.c_void_p(coef.ctypes.data), mn, mx, no,
                         C.c_void_p(hout.data_ptr()), C.byref(ie))
    t1 = time.perf_counter()
    print(f"rep {rep}: {1e3 * (t1 - t0):.1f} ms ierr {ie.value}", flush=True)
# plain copies for reference
d = torch.empty((nq, 3), dtype=torch.float64, device="cuda"); o = torch.empty(nq, dtype=torch.float64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hq, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    hout.copy_(o, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"H2D 2.4 GB {1e3 * (t1 - t0):.1f} ms = {2.4 / (t1 - t0):.1f} GB/s; D2H 0.8 GB {1e3 * (t2 - t1):.1f} ms = {0.8 / (t2 - t1):.1f} GB/s")
