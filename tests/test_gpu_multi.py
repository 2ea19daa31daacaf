"""Multi-GPU through the C ABI alone (no torch.distributed): the path a Fortran / C host takes.

One process drives two devices: splpak_b200_comm_init_all creates the communicators (ncclCommInitAll through the
lazily loaded libnccl.so.2), every device assembles its shard, splpak_b200_fit_allreduce sums the partial normal
equations inside an NCCL group, every device solves.  Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise.
"""
import ctypes as C

import numpy as np
import pytest

import splpak_b200 as sp
from util import make_problem

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.parametrize("deterministic", [False, True])
def test_c_abi_allreduce_two_devices(oracle, monkeypatch, deterministic):
    import torch

    if deterministic:
        # SPLPAK_B200_DETERMINISTIC=1: every rank's partial buffer is order-independent, the all-reduce hands every rank the
        # same sums and the constraint rows go through integer limbs -> the replicated solves agree BITWISE
        monkeypatch.setenv("SPLPAK_B200_DETERMINISTIC", "1")
    lib = sp.load()
    ndim, nodes = 2, [9, 8]
    x, y, w, mn, mx = make_problem(ndim, nodes, 20000, seed=71, hole=True)
    ref, ie = oracle.initialize(ndim, x, y, w, mn, mx, nodes, 1.0)
    assert ie == 0
    devs = (C.c_int * 2)(0, 1)
    comms = (C.c_void_p * 2)()
    assert lib.splpak_b200_comm_init_all(2, devs, comms) == 0
    handles, coefs = [], []
    half = len(x) // 2
    for r in range(2):
        torch.cuda.set_device(r)
        h = sp.FitHandle(ndim, mn, mx, nodes, 1.0)
        assert h.ierror == 0
        lo, hi = (0, half) if r == 0 else (half, len(x))
        assert h.add_points(x[lo:hi], y[lo:hi], w[lo:hi], weighted=True) == 0
        handles.append(h)
    assert lib.splpak_b200_comm_group_start() == 0
    for r in range(2):
        torch.cuda.set_device(r)
        assert lib.splpak_b200_fit_allreduce(handles[r].h, comms[r]) == 0
    assert lib.splpak_b200_comm_group_end() == 0
    for r in range(2):
        torch.cuda.set_device(r)
        c, ierr = handles[r].compute()
        assert ierr == 0
        coefs.append(c)
    # refinement across the two devices: all-reduce of the right-hand side between the residual pass and the solve
    for step in range(2):
        for r in range(2):
            torch.cuda.set_device(r)
            lo, hi = (0, half) if r == 0 else (half, len(x))
            assert lib.splpak_b200_fit_refine_begin(handles[r].h) == 0
            xa, ya, wa = (np.ascontiguousarray(a[lo:hi]) for a in (x, y, w))
            assert lib.splpak_b200_fit_refine_add_points(handles[r].h, C.c_void_p(xa.ctypes.data), ndim,
                                                         C.c_void_p(ya.ctypes.data), C.c_void_p(wa.ctypes.data), 1,
                                                         hi - lo) == 0
        assert lib.splpak_b200_comm_group_start() == 0
        for r in range(2):
            torch.cuda.set_device(r)
            assert lib.splpak_b200_fit_allreduce_rhs(handles[r].h, comms[r]) == 0
        assert lib.splpak_b200_comm_group_end() == 0
        for r in range(2):
            torch.cuda.set_device(r)
            ierr = C.c_int(0)
            lib.splpak_b200_fit_refine_compute(handles[r].h, C.c_void_p(coefs[r].ctypes.data), len(coefs[r]), C.byref(ierr))
            assert ierr.value == 0
    scale = np.abs(ref).max()
    # The constraint rows are added through integer limbs in every mode, so both ranks factor the SAME system and hold the
    # SAME coefficients bit for bit (round 2; before, the replicas differed by ~eps cond(G), each rank formed the residual
    # of its shard with its own coefficients, and the refinement stalled at that difference: 2.5e-10 .. 5e-10 here,
    # scripts/multi_refine_check.py).  cond(A) = 2e4: the refined coefficients are at eps cond(A) of the oracle's QR.
    assert np.array_equal(coefs[0], coefs[1]), np.abs(coefs[0] - coefs[1]).max() / scale
    for r in range(2):
        assert np.abs(coefs[r] - ref).max() <= 1e-11 * scale, np.abs(coefs[r] - ref).max() / scale
    for r in range(2):
        torch.cuda.set_device(r)
        handles[r].destroy()
        assert lib.splpak_b200_comm_destroy(comms[r]) == 0
    torch.cuda.set_device(0)
