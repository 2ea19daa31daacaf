"""Pin of the oracle against the COMPILED reference, when one can be built.

`make -C oracle ref` compiles /root/reference/src/splpak.F90 + oracle/ref_golden.f90 (a driver that prints the
reference's own coefficients and splde values) into oracle/_ref/ref_golden when a Fortran compiler exists.  This
image has none (gfortran, flang, nvfortran, ifx, lfortran, f951 all absent), so here the test SKIPS with that
reason; on a machine with a compiler it checks the C oracle and the numpy model against the real module and
rewrites tests/golden/reference_cases.json from the binary's output.
"""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import Oracle
from oracle.numpy_model import SplpakModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_golden")
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_cases.json")


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for ndim, nodes, ndata, weighted, xtrap in [(1, [10], 20, True, 1.0), (2, [5, 6], 200, False, 1.0),
                                                (3, [4, 4, 5], 500, True, 0.0), (2, [6, 6], 120, True, 1.0)]:
        x = rng.random((ndata, ndim))
        if ndim == 2 and weighted:
            x = x[np.linalg.norm(x - 0.5, axis=1) > 0.3]          # hole: constraint rows fire
        y = np.cos(x.sum(axis=1))
        w = rng.uniform(0.5, 1.5, len(x)) if weighted else None
        q = rng.random((5, ndim)) * 1.4 - 0.2
        nd = rng.integers(0, 3, (5, ndim))
        out.append(dict(ndim=ndim, nodes=nodes, x=x, y=y, w=w, xtrap=xtrap, q=q, nd=nd))
    return out


def _fmt(case):
    nd, x, y, w = case["ndim"], case["x"], case["y"], case["w"]
    lines = [f"{nd} {len(x)} {len(case['q'])} {1 if w is not None else 0} {case['xtrap']!r}",
             " ".join(str(n) for n in case["nodes"]), " ".join(["0.0"] * nd), " ".join(["1.0"] * nd)]
    for i in range(len(x)):
        row = [repr(float(v)) for v in x[i]] + [repr(float(y[i]))]
        if w is not None:
            row.append(repr(float(w[i])))
        lines.append(" ".join(row))
    for p, d in zip(case["q"], case["nd"]):
        lines.append(" ".join([repr(float(v)) for v in p] + [str(int(v)) for v in d]))
    return "\n".join(lines) + "\n"


def test_oracle_against_compiled_reference():
    if not os.path.exists(BIN):
        fc = next((c for c in ("gfortran", "flang-new", "flang", "nvfortran", "ifx") if shutil.which(c)), None)
        if fc is None or not os.path.exists("/root/reference/src/splpak.F90"):
            pytest.skip("no Fortran compiler in this image (gfortran/flang/nvfortran/ifx absent) or no /root/reference: "
                        "oracle/_ref cannot be built; the oracle is double-restated instead (tests/test_oracle_double.py)")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    cases = _cases()
    res = subprocess.run([BIN], input="".join(_fmt(c) for c in cases), capture_output=True, text=True, check=True)
    tok = res.stdout.split()
    pos = 0
    golden = []
    o = Oracle()
    for c in cases:
        assert tok[pos] == "ierror"
        ierr = int(tok[pos + 1])
        assert tok[pos + 2] == "coef"
        ncol = int(tok[pos + 3])
        coef = np.array([float(t) for t in tok[pos + 4: pos + 4 + ncol]])
        pos += 4 + ncol
        assert tok[pos] == "eval"
        nq = int(tok[pos + 1])
        vals = np.array([float(t) for t in tok[pos + 2: pos + 2 + 2 * nq: 2]])
        pos += 2 + 2 * nq
        mn, mx = [0.0] * c["ndim"], [1.0] * c["ndim"]
        c1, ie1 = o.initialize(c["ndim"], c["x"], c["y"], c["w"], mn, mx, c["nodes"], c["xtrap"])
        assert ie1 == ierr
        A, _ = o.rows(c["ndim"], c["x"], c["y"], c["w"], mn, mx, c["nodes"], c["xtrap"])
        tol = 50 * np.finfo(float).eps * np.linalg.cond(A)
        assert np.abs(c1 - coef).max() <= tol * np.abs(coef).max()
        m = SplpakModel()
        for p, d, v in zip(c["q"], c["nd"], vals):
            v1, _ = o.evaluate(c["ndim"], p, coef, mn, mx, c["nodes"], nderiv=list(d))
            v2, _ = m.splde(c["ndim"], p, list(d), coef, mn, mx, c["nodes"])
            assert v1 == v and v2 == v            # same arithmetic, unfused: bit-equal
        golden.append(dict(ndim=c["ndim"], nodes=c["nodes"], xtrap=c["xtrap"], coef=coef.tolist(), values=vals.tolist()))
    json.dump(golden, open(GOLDEN, "w"))
