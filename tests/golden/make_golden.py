"""Writes tests/golden/splpak_test_linear.json: the inputs and asserted outputs of the reference's own
deterministic test, /root/reference/test/splpak_test_linear.f90 (:14-20 sizes, :41-48 data, :60-89
assertions), plus the analytic coefficients of SURVEY 8c K2 (27*coef for f = 2x on 10 nodes).
The reference itself cannot be run here (no Fortran compiler), so the expected values are the test's
own asserted quantities and closed-form answers, not captured program output."""
import json
import os

nxdata, nodes = 20, [10]
ncol = 10
g = {
    "source": "test/splpak_test_linear.f90",
    "ndim": 1,
    "nodes": nodes,
    "nwrk": ncol * (ncol + 1) + 1,
    "xtrap": 1.0,
    "xmin": [0.0],
    "xmax": [1.0],
    "xdata": [(i) / (nxdata - 1) for i in range(nxdata)],
    "wdata": [1.0] * nxdata,
    "x_est": [i / 100 for i in range(100)],
    "errmax_tol": 1e-1,
    "slope_tol": 1e-12,
    "coef_times_27": [-4, 2, 8, 12, 16, 20, 24, 28, 16, -14],
}
g["ydata"] = [2.0 * x for x in g["xdata"]]
json.dump(g, open(os.path.join(os.path.dirname(__file__), "splpak_test_linear.json"), "w"), indent=1)
