"""ctypes binding of the C oracle (oracle/splpak_oracle.c) -- TEST INFRASTRUCTURE ONLY.

The oracle restates /root/reference/src/splpak.F90 (bascmp :206-389, splcw :512-1060,
splcc :421-446, splde :1089-1240, splfe :1258-1275, suprls :1375-1695) in C.
Nothing here is used by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_lib_path(real32: bool = False) -> str:
    return os.path.join(_HERE, "liboracle_splpak_r32.so" if real32 else "liboracle_splpak.so")


def build_oracle(force: bool = False) -> None:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "splpak_oracle.c")
    stale = any(
        (not os.path.exists(p)) or os.path.getmtime(p) < os.path.getmtime(src)
        for p in (oracle_lib_path(False), oracle_lib_path(True))
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    # the reference itself: only when a Fortran compiler and /root/reference exist (never on the GPU box)
    if os.path.exists("/root/reference/src/splpak.F90"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=False, capture_output=True)


class _State(C.Structure):
    _MAXDIM = 8


def _make_state(real):
    class State(C.Structure):
        _fields_ = [
            ("mdim", C.c_int),
            ("dx", real * 8),
            ("dxin", real * 8),
            ("ib", C.c_int * 8),
            ("ibmn", C.c_int * 8),
            ("ibmx", C.c_int * 8),
            ("ilast", C.c_longlong),
            ("isav", C.c_longlong),
            ("iold", C.c_longlong),
            ("np1", C.c_longlong),
            ("l", C.c_longlong),
            ("il1", C.c_longlong),
            ("k", C.c_longlong),
            ("k1", C.c_longlong),
            ("errsum", real),
            ("quiet", C.c_int),
            ("nmsg", C.c_longlong),
            ("last_suprls_ier", C.c_int),
            ("row_sink", C.c_void_p),
            ("row_sink_ctx", C.c_void_p),
        ]

    return State


class Oracle:
    """Mirror of the reference's splpak_type on the CPU (initialize/evaluate), plus helpers."""

    def __init__(self, real32: bool = False, quiet: bool = True):
        build_oracle()
        self.real32 = real32
        self.dtype = np.float32 if real32 else np.float64
        self._real = C.c_float if real32 else C.c_double
        self.lib = C.CDLL(oracle_lib_path(real32))
        assert self.lib.oracle_sizeof_real() == np.dtype(self.dtype).itemsize
        self._State = _make_state(self._real)
        self.state = self._State()
        self.lib.oracle_init(C.byref(self.state))
        self.state.quiet = 1 if quiet else 0
        rp = C.POINTER(self._real)
        ip = C.POINTER(C.c_int)
        L = self.lib
        L.oracle_splcw.argtypes = [C.c_void_p, C.c_int, rp, C.c_int, rp, rp, C.c_longlong, rp, rp, ip,
                                   self._real, rp, C.c_longlong, rp, C.c_longlong, ip]
        L.oracle_splcw.restype = None
        L.oracle_splcc.argtypes = [C.c_void_p, C.c_int, rp, C.c_int, rp, C.c_longlong, rp, rp, ip,
                                   self._real, rp, C.c_longlong, rp, C.c_longlong, ip]
        L.oracle_splcc.restype = None
        L.oracle_splde.argtypes = [C.c_void_p, C.c_int, rp, ip, rp, rp, rp, ip, ip]
        L.oracle_splde.restype = self._real
        L.oracle_splfe.argtypes = [C.c_void_p, C.c_int, rp, rp, rp, rp, ip, ip]
        L.oracle_splfe.restype = self._real
        L.oracle_eval_batch.argtypes = [C.c_int, rp, C.c_int, C.c_longlong, ip, rp, rp, rp, ip, rp]
        L.oracle_eval_batch.restype = C.c_int
        L.oracle_rows.argtypes = [C.c_int, rp, C.c_int, rp, rp, C.c_longlong, rp, rp, ip, self._real,
                                  rp, rp, C.c_longlong]
        L.oracle_rows.restype = C.c_longlong
        L.oracle_suprls_steady_sample.argtypes = [C.c_int, rp, C.c_int, rp, rp, C.c_longlong, rp, rp, ip]
        L.oracle_suprls_steady_sample.restype = C.c_double
        L.oracle_bascmp.argtypes = [C.c_void_p, rp, ip, rp, ip, C.POINTER(C.c_longlong), rp]
        L.oracle_bascmp.restype = None

    # -- helpers ---------------------------------------------------------------------------
    def _r(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        return a, a.ctypes.data_as(C.POINTER(self._real))

    @staticmethod
    def _i(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a, a.ctypes.data_as(C.POINTER(C.c_int))

    @staticmethod
    def _xdata(xdata, ndim):
        """Accept (ndata, l1xdat) C-order == Fortran xdata(l1xdat, ndata)."""
        x = np.asarray(xdata)
        if x.ndim == 1:
            x = x.reshape(-1, 1)
        return x

    # -- reference API ----------------------------------------------------------------------
    def initialize(self, ndim, xdata, ydata, wdata, xmin, xmax, nodes, xtrap, nwrk=None, ncf=None,
                   ndata=None):
        """splcw (wdata given) or splcc (wdata None).  Returns (coef, ierror).

        xdata: array (ndata, l1xdat) in C order, i.e. the memory image of Fortran xdata(l1xdat, ndata).
        """
        x = self._xdata(xdata, ndim)
        l1xdat = x.shape[1]
        nd = x.shape[0] if ndata is None else ndata
        ncol = int(np.prod(np.asarray(nodes[:max(ndim, 0)], dtype=np.int64))) if ndim >= 1 else 1
        if ncf is None:
            ncf = ncol
        if nwrk is None:
            nwrk = ncol * (ncol + 1) + 1
        xa, xp = self._r(x)
        ya, yp = self._r(ydata)
        mn, mnp = self._r(xmin)
        mx, mxp = self._r(xmax)
        no, nop = self._i(nodes)
        coef = np.zeros(max(int(ncf), 1), dtype=self.dtype)
        work = np.zeros(max(int(nwrk), 1), dtype=self.dtype)
        ierr = C.c_int(0)
        cp = coef.ctypes.data_as(C.POINTER(self._real))
        wp = work.ctypes.data_as(C.POINTER(self._real))
        if wdata is None:
            self.lib.oracle_splcc(C.byref(self.state), ndim, xp, l1xdat, yp, nd, mnp, mxp, nop,
                                  self._real(xtrap), cp, ncf, wp, nwrk, C.byref(ierr))
        else:
            wa, wpp = self._r(wdata)
            self.lib.oracle_splcw(C.byref(self.state), ndim, xp, l1xdat, yp, wpp, nd, mnp, mxp, nop,
                                  self._real(xtrap), cp, ncf, wp, nwrk, C.byref(ierr))
        return coef, ierr.value

    def evaluate(self, ndim, x, coef, xmin, xmax, nodes, nderiv=None):
        """splfe (nderiv None) or splde at one point.  Returns (value, ierror)."""
        xa, xp = self._r(np.atleast_1d(x))
        ca, cp = self._r(coef)
        mn, mnp = self._r(xmin)
        mx, mxp = self._r(xmax)
        no, nop = self._i(nodes)
        ierr = C.c_int(0)
        if nderiv is None:
            v = self.lib.oracle_splfe(C.byref(self.state), ndim, xp, cp, mnp, mxp, nop, C.byref(ierr))
        else:
            nd, ndp = self._i(nderiv)
            v = self.lib.oracle_splde(C.byref(self.state), ndim, xp, ndp, cp, mnp, mxp, nop, C.byref(ierr))
        return float(v), ierr.value

    def evaluate_batch(self, ndim, x, coef, xmin, xmax, nodes, nderiv=None):
        """Loop of scalar splfe/splde calls over x (nq, l1x).  Returns (values, worst ierror)."""
        x = self._xdata(x, ndim)
        xa, xp = self._r(x)
        ca, cp = self._r(coef)
        mn, mnp = self._r(xmin)
        mx, mxp = self._r(xmax)
        no, nop = self._i(nodes)
        out = np.zeros(x.shape[0], dtype=self.dtype)
        ndp = None
        if nderiv is not None:
            nd, ndp = self._i(nderiv)
        ierr = self.lib.oracle_eval_batch(ndim, xp, x.shape[1], x.shape[0], ndp, cp, mnp, mxp, nop,
                                          out.ctypes.data_as(C.POINTER(self._real)))
        return out, ierr

    def rows(self, ndim, xdata, ydata, wdata, xmin, xmax, nodes, xtrap):
        """The least-squares rows splcw would hand to suprls: (A, r) dense."""
        x = self._xdata(xdata, ndim)
        ncol = int(np.prod(np.asarray(nodes[:ndim], dtype=np.int64)))
        npairs = ndim * (ndim + 1) // 2
        maxrows = x.shape[0] + ncol * npairs
        A = np.zeros((maxrows, ncol), dtype=self.dtype)
        r = np.zeros(maxrows, dtype=self.dtype)
        xa, xp = self._r(x)
        ya, yp = self._r(ydata)
        w = np.array([-1.0]) if wdata is None else wdata
        wa, wpp = self._r(w)
        mn, mnp = self._r(xmin)
        mx, mxp = self._r(xmax)
        no, nop = self._i(nodes)
        n = self.lib.oracle_rows(ndim, xp, x.shape[1], yp, wpp, x.shape[0], mnp, mxp, nop,
                                 self._real(xtrap), A.ctypes.data_as(C.POINTER(self._real)),
                                 r.ctypes.data_as(C.POINTER(self._real)), maxrows)
        if n < 0:
            raise RuntimeError(f"oracle rows: ierror {-n}")
        return A[:n], r[:n]

    def bascmp(self, ib, x, nderiv, xmin, xmax, nodes):
        """One bascmp call with the node index vector ib (0-based node indices as in the reference)."""
        ndim = len(ib)
        st = self.state
        st.mdim = ndim
        for d in range(ndim):
            dx = (xmax[d] - xmin[d]) / (nodes[d] - 1)
            st.dx[d] = dx
            st.dxin[d] = self.dtype(1.0) / self.dtype(dx)
            st.ib[d] = int(ib[d])
        xa, xp = self._r(x)
        nd, ndp = self._i(nderiv)
        mn, mnp = self._r(xmin)
        no, nop = self._i(nodes)
        icol = C.c_longlong(0)
        basm = self._real(0)
        self.lib.oracle_bascmp(C.byref(st), xp, ndp, mnp, nop, C.byref(icol), C.byref(basm))
        return icol.value, float(basm.value)

    def suprls_steady_sample(self, ndim, xdata, ydata, wdata, m, xmin, xmax, nodes):
        """Seconds for m data rows entering a full triangle (CPU-baseline sample; see bench.py)."""
        x = self._xdata(xdata, ndim)
        xa, xp = self._r(x)
        ya, yp = self._r(ydata)
        w = np.array([-1.0]) if wdata is None else wdata
        wa, wpp = self._r(w)
        mn, mnp = self._r(xmin)
        mx, mxp = self._r(xmax)
        no, nop = self._i(nodes)
        return float(self.lib.oracle_suprls_steady_sample(ndim, xp, x.shape[1], yp, wpp, m, mnp, mxp, nop))
