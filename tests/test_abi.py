"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and returns the reference's argument-error codes without needing a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import splpak_b200 as sp
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "splpak_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(splpak_b200_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("real32", [False, True])
def test_library_exports_every_declared_symbol(real32):
    lib = C.CDLL(sp.lib_path(real32))
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/splpak_b200.h but not exported"
    assert sorted(sp.SYMBOLS) == declared
    assert lib.splpak_b200_sizeof_real() == (4 if real32 else 8)


def test_library_has_no_oracle_or_torch_dependency():
    out = subprocess.run(["ldd", sp.lib_path(False)], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", sp.lib_path(False)], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


def test_sass_is_blackwell_native():
    """The built library carries sm_100a SASS with the FP64 tensor MMA, the bulk async copy and f64 reductions."""
    out = subprocess.run(["cuobjdump", "-sass", sp.lib_path(False)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("DMMA.8x8x4", "UBLKCP", "REDG.E.ADD.F64"):
        assert mnemonic in out, mnemonic


def test_fit_argument_errors_in_reference_order():
    """Codes 101..106 and their order (src/splpak.F90:718-781); these never touch the device."""
    x = np.linspace(0, 1, 30).reshape(-1, 1)
    y = 2 * x[:, 0]
    w = np.ones(30)
    a = dict(quiet=True)
    assert sp.splcw(0, x, 1, y, w, 30, [0.0], [1.0], [10], 1.0, **a)[1] == 101
    assert sp.splcw(5, np.zeros((30, 5)), 5, y, w, 30, [0.0] * 5, [1.0] * 5, [4] * 5, 1.0, **a)[1] == 101
    assert sp.splcw(1, x, 1, y, w, 30, [0.0], [1.0], [3], 1.0, **a)[1] == 102
    assert sp.splcw(1, x, 1, y, w, 30, [0.5], [0.5], [10], 1.0, **a)[1] == 103
    assert sp.splcw(2, np.zeros((30, 2)), 2, y, w, 30, [0.0, 1.0], [1.0, 1.0], [3, 10], 1.0, **a)[1] == 102
    assert sp.splcw(1, x, 1, y, w, 30, [0.0], [1.0], [10], 1.0, ncf=9, **a)[1] == 104
    assert sp.splcw(1, x, 1, y, w, 0, [0.0], [1.0], [10], 1.0, **a)[1] == 105
    assert sp.splcw(1, x, 1, y, w, 30, [0.0], [1.0], [10], 1.0, nwrk=10, **a)[1] == 106
    assert sp.splcw(1, x, 1, y, w, 30, [0.0], [1.0], [10], 0.0, nwrk=0, **a)[1] == 106
    # 104 is checked before 105, 105 before 106
    assert sp.splcw(1, x, 1, y, w, 0, [0.0], [1.0], [10], 1.0, ncf=9, nwrk=1, **a)[1] == 104
    assert sp.splcw(1, x, 1, y, w, 0, [0.0], [1.0], [10], 1.0, nwrk=1, **a)[1] == 105
    assert sp.splcc(1, x, 1, y, 30, [0.0], [1.0], [3], 1.0, **a)[1] == 102


def test_eval_argument_errors():
    coef = np.ones(10)
    a = dict(quiet=True)
    assert sp.splfe(0, [0.5], coef, [0.0], [1.0], [10], **a)[1] == 101
    assert sp.splfe(1, [0.5], coef, [0.0], [1.0], [3], **a)[1] == 102
    assert sp.splfe(1, [0.5], coef, [1.0], [1.0], [10], **a)[1] == 103
    assert sp.splde(1, [0.5], [0], coef, [1.0], [1.0], [10], **a)[1] == 103


def test_cfaerr_text_matches_reference_format(capsys):
    """' IERR=' I5 then the message (src/splpak.F90:404-405), messages from :720-779."""
    assert sp.cfaerr_text(102, False) == " IERR=  102\n splcc or splcw - NODES(IDIM) is less than 4 for some IDIM\n"
    assert sp.cfaerr_text(104, True).startswith(" IERR=  104\n splde - NDERIV(IDIM) IS less than 0")
    s = sp.SplpakType()
    x = np.linspace(0, 1, 30).reshape(-1, 1)
    coef, ierr = s.initialize(1, x, 1, x[:, 0], np.ones(30), 30, [0.0], [1.0], [3], 1.0)
    assert ierr == 102
    assert capsys.readouterr().out == sp.cfaerr_text(102, False)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    """The product path fails loudly (201) instead of computing on the CPU."""
    x = np.linspace(0, 1, 30).reshape(-1, 1)
    coef, ierr = sp.splcw(1, x, 1, 2 * x[:, 0], np.ones(30), 30, [0.0], [1.0], [10], 1.0, quiet=True)
    assert ierr == 201 and not coef.any()
    v, ierr = sp.eval_batch(1, x, np.ones(10), [0.0], [1.0], [10])
    assert ierr == 201 and not v.any()
    h = sp.FitHandle(1, [0.0], [1.0], [10], 1.0)
    assert h.ierror == 201


def test_product_package_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "splpak_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.lower() or f == "basis.cuh", f"{f} mentions the oracle"


def test_fortran_shim_binds_only_exported_symbols():
    """Every bind(C, name=...) of fortran/splpak_module.F90 (which cannot be compiled here: no Fortran compiler)
    names a symbol the header declares and the library exports, and the shim keeps the reference's public names."""
    import re

    from splpak_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "fortran", "splpak_module.F90")).read()
    names = set(re.findall(r"bind\(C,\s*name='([A-Za-z0-9_]+)'\)", src))
    assert names, "no bind(C) interfaces found"
    assert names <= set(_lib.SYMBOLS), names - set(_lib.SYMBOLS)
    for must in ("splpak_b200_splcw", "splpak_b200_splcc", "splpak_b200_splde", "splpak_b200_splfe"):
        assert must in names
    # the reference's public surface (src/splpak.F90:43-45, :117-119)
    for text in ("module splpak_module", "type,public :: splpak_type", "integer,parameter,public :: splpak_wp",
                 "generic,public   :: initialize => splcc, splcw", "generic,public   :: evaluate   => splfe, splde",
                 "procedure,public :: destroy"):
        assert text in src, text
