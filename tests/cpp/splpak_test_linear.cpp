// C++ rendition of the reference's test/splpak_test_linear.f90 (program splpak_test_linesr) and
// test/splpak_test.f90, statement for statement, against the splpak_type mirror
// (splpak_b200/csrc/splpak_type.hpp).  The pyplot calls (:91-102) are plotting only and are dropped.
// Exit code 0 = every `error stop` condition of the reference tests is avoided.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../splpak_b200/csrc/splpak_type.hpp"

using splpak::wp;

static wp f1_linear(wp x) { return 2.0 * x; }                                   // :108-112
static wp f1_smooth(wp x) { return 0.5 * (x * std::exp(-x) + std::sin(x)); }    // splpak_test.f90:106-110

static int run(bool linear) {
    const int ndim = 1, nxdata = 20;
    const int nodes[1] = {10};
    const int ncol = 10, nwrk = ncol * (ncol + 1) + 1, ncf = ncol, nxdata_est = 100;   // :14-20
    std::vector<wp> xdata(ndim * nxdata), ydata(nxdata), wdata(nxdata), work(nwrk), coef(ncf);
    wp xmin[1] = {0.0}, xmax[1] = {1.0}, x[1];
    const wp xtrap = 1.0;
    int ierror = 0;
    splpak::splpak_type solver;
    unsigned long long s = 42;   // the reference seeds gfortran's generator with 42 (not portable)
    for (int i = 1; i <= nxdata; ++i) {
        wp r = 0.0;
        if (!linear) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            r = ((double)(s >> 11) / 9007199254740992.0 - 0.5) / 10.0;
        }
        wdata[i - 1] = 1.0 - std::fabs(r);
        xdata[i - 1] = (wp)(i - 1) / (wp)(nxdata - 1);
        ydata[i - 1] = (linear ? f1_linear(xdata[i - 1]) : f1_smooth(xdata[i - 1])) + r;
    }
    solver.initialize(1, xdata.data(), 1, ydata.data(), wdata.data(), nxdata, xmin, xmax, nodes, xtrap,
                      coef.data(), ncf, work.data(), nwrk, ierror);
    std::printf(" splcw ierror = %d\n", ierror);
    if (ierror != 0) { std::puts("error calling splcw"); return 1; }
    wp errmax = 0.0;
    for (int i = 1; i <= nxdata_est; ++i) {
        x[0] = (wp)(i - 1) / nxdata_est;
        const wp f = solver.evaluate(ndim, x, coef.data(), xmin, xmax, nodes, ierror);
        if (ierror != 0) { std::puts("error calling splfe"); return 1; }
        const wp tru = linear ? f1_linear(x[0]) : f1_smooth(x[0]);
        errmax = std::fmax(std::fabs(tru - f), errmax);
    }
    std::printf(" splfe errmax%s = %g\n", linear ? " [linear]" : "", (double)errmax);
    if (std::fabs(errmax) > 1.0e-1) { std::puts("errmax too large"); return 1; }
    if (linear) {
        const int nd1[1] = {1};
        wp x0[1] = {0.0}, x1[1] = {1.0};
        const wp fleft = solver.evaluate(ndim, x0, nd1, coef.data(), xmin, xmax, nodes, ierror);
        if (ierror != 0) { std::puts("error calling splde"); return 1; }
        std::printf(" splde errmax left [linear] = %g\n", (double)(fleft - 2.0));
        if (std::fabs(fleft - 2.0) > 1.0e-12) { std::puts("errmax too large"); return 1; }
        const wp fright = solver.evaluate(ndim, x1, nd1, coef.data(), xmin, xmax, nodes, ierror);
        if (ierror != 0) { std::puts("error calling splde"); return 1; }
        std::printf(" splde errmax right [linear] = %g\n", (double)(fright - 2.0));
        if (std::fabs(fright - 2.0) > 1.0e-12) { std::puts("errmax too large"); return 1; }
        const double c27[10] = {-4, 2, 8, 12, 16, 20, 24, 28, 16, -14};
        for (int k = 0; k < 10; ++k)
            if (std::fabs(27.0 * coef[k] - c27[k]) > 1e-10) { std::printf("coef[%d] off\n", k); return 1; }
    }
    solver.destroy();
    return 0;
}

int main() {
    if (run(true)) return 1;
    if (run(false)) return 1;
    std::puts("ok");
    return 0;
}
