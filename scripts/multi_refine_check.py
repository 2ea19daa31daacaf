"""2-GPU C-ABI flow of tests/test_gpu_multi.py, printing the coefficient error after the solve and after every
refinement step (several repetitions, default and deterministic mode; one GPU when only one is visible)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import splpak_b200 as sp
from oracle import Oracle
from util import make_problem

lib = sp.load()
o = Oracle()
ndim, nodes = 2, [9, 8]
x, y, w, mn, mx = make_problem(ndim, nodes, 20000, seed=71, hole=True)
ref, ie = o.initialize(ndim, x, y, w, mn, mx, nodes, 1.0)
scale = np.abs(ref).max()
ngpu = min(2, torch.cuda.device_count())
for det in ("0", "1"):
    os.environ["SPLPAK_B200_DETERMINISTIC"] = det
    for rep in range(4):
        devs = (C.c_int * 2)(0, 1)
        comms = (C.c_void_p * 2)()
        if ngpu == 2:
            assert lib.splpak_b200_comm_init_all(2, devs, comms) == 0
        handles, coefs, errs = [], [], []
        bounds = np.linspace(0, len(x), ngpu + 1).astype(int)
        for r in range(ngpu):
            torch.cuda.set_device(r)
            h = sp.FitHandle(ndim, mn, mx, nodes, 1.0)
            lo, hi = bounds[r], bounds[r + 1]
            assert h.add_points(x[lo:hi], y[lo:hi], w[lo:hi], weighted=True) == 0
            handles.append(h)
        if ngpu == 2:
            lib.splpak_b200_comm_group_start()
            for r in range(2):
                torch.cuda.set_device(r)
                assert lib.splpak_b200_fit_allreduce(handles[r].h, comms[r]) == 0
            lib.splpak_b200_comm_group_end()
        for r in range(ngpu):
            torch.cuda.set_device(r)
            c, ierr = handles[r].compute()
            assert ierr == 0
            coefs.append(c)
        errs.append(max(np.abs(c - ref).max() for c in coefs) / scale)
        for step in range(3):
            for r in range(ngpu):
                torch.cuda.set_device(r)
                lo, hi = bounds[r], bounds[r + 1]
                assert lib.splpak_b200_fit_refine_begin(handles[r].h) == 0
                xa, ya, wa = (np.ascontiguousarray(a[lo:hi]) for a in (x, y, w))
                assert lib.splpak_b200_fit_refine_add_points(handles[r].h, C.c_void_p(xa.ctypes.data), ndim,
                                                             C.c_void_p(ya.ctypes.data), C.c_void_p(wa.ctypes.data), 1, int(hi - lo)) == 0
            if ngpu == 2:
                lib.splpak_b200_comm_group_start()
                for r in range(2):
                    torch.cuda.set_device(r)
                    assert lib.splpak_b200_fit_allreduce_rhs(handles[r].h, comms[r]) == 0
                lib.splpak_b200_comm_group_end()
            for r in range(ngpu):
                torch.cuda.set_device(r)
                ierr = C.c_int(0)
                lib.splpak_b200_fit_refine_compute(handles[r].h, C.c_void_p(coefs[r].ctypes.data), len(coefs[r]), C.byref(ierr))
                assert ierr.value == 0
            errs.append(max(np.abs(c - ref).max() for c in coefs) / scale)
        rep_diff = np.abs(coefs[0] - coefs[-1]).max() / scale
        print(f"gpus {ngpu} det={det} rep {rep}: error vs oracle: solve {errs[0]:.2e}, refine " + ", ".join(f"{e:.2e}" for e in errs[1:]) +
              f"; replicas differ by {rep_diff:.2e}", flush=True)
        for r in range(ngpu):
            torch.cuda.set_device(r)
            handles[r].destroy()
            if ngpu == 2:
                lib.splpak_b200_comm_destroy(comms[r])
        torch.cuda.set_device(0)
