// host_ops.hpp -- plain C++ restatement of the C-ABI operations the Newton-Krylov-multigrid host logic uses
// (TEST INFRASTRUCTURE ONLY: never compiled into libp4b200.so).
//
// p4pdes_b200/csrc/nk_solver.hpp is a template over an `Ops` type.  The product instantiates it with DeviceOps (CUDA
// kernels).  This header provides HostOps -- each method a few loops following the reference formula it cites -- so
// that oracle/native/nk_host_main.cpp can run the SAME solver logic on a machine without a GPU and
// tests/test_native_nk_cpu.py can compare it with the Python oracle (oracle/minimal_solver_oracle.py).
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

struct HostOps {
    int err = 0;
    long long allocs = 0, frees = 0;
    int error() const { return err; }
    double *alloc(size_t n) { allocs++; return (double *)malloc(sizeof(double) * std::max<size_t>(n, 1)); }
    void release(double *p) { if (p) { frees++; free(p); } }
    void to_host(const double *s, double *d, size_t n) { memcpy(d, s, sizeof(double) * n); }
    void from_host(const double *s, double *d, size_t n) { memcpy(d, s, sizeof(double) * n); }

    // [PETSc] Vec operations
    double dot(size_t n, const double *x, const double *y) { double s = 0; for (size_t i = 0; i < n; i++) s += x[i] * y[i]; return s; }
    double norm2(size_t n, const double *x) { return sqrt(dot(n, x, x)); }
    double norminf(size_t n, const double *x) { double m = 0; for (size_t i = 0; i < n; i++) m = std::max(m, fabs(x[i])); return m; }
    void axpy(size_t n, double a, const double *x, double *y) { for (size_t i = 0; i < n; i++) y[i] += a * x[i]; }
    void aypx(size_t n, double a, const double *x, double *y) { for (size_t i = 0; i < n; i++) y[i] = x[i] + a * y[i]; }
    void axpby(size_t n, double a, const double *x, double b, const double *y, double *out) {
        for (size_t i = 0; i < n; i++) out[i] = (x ? a * x[i] : 0.0) + (y ? b * y[i] : 0.0);
    }
    void copy(size_t n, const double *x, double *y) { memmove(y, x, sizeof(double) * n); }
    void set(size_t n, double a, double *y) { for (size_t i = 0; i < n; i++) y[i] = a; }

    // c/ch7/minimal.c:27-42  g_bdry_tent / g_bdry_catenoid at every node of the unit square
    void minimal_sample(int mx, int my, int problem, double H, double c, double *g) {
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const double x = i * (1.0 / (mx - 1)), y = j * (1.0 / (my - 1));
                g[j * mx + i] = problem == 0 ? (x < 1.0e-8 ? 2.0 * H * (y < 0.5 ? y : 1.0 - y) : 0.0)
                                             : c * cosh(x / c) * sin(acos((y / c) / cosh(x / c)));
            }
    }
    // c/ch6/poissonfunctions.c:260-346  InitialState(ZEROS, gonboundary)
    void initial_state2d(int mx, int my, const double *g, double *u) {
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const bool bd = i == 0 || j == 0 || i == mx - 1 || j == my - 1;
                u[j * mx + i] = bd ? g[j * mx + i] : 0.0;
            }
    }
    // c/ch7/minimal.c:210-282  FormFunctionLocal
    void minimal_function(int mx, int my, double q, const double *u, const double *g, double *FF) {
        const double hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1), hxhy = hx / hy, hyhx = hy / hx;
        auto DD = [q](double w) { return pow(1.0 + w, q); };                           // minimal.c:46-48
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const int n = j * mx + i;
                if (j == 0 || i == 0 || i == mx - 1 || j == my - 1) { FF[n] = u[n] - g[n]; continue; }
                auto val = [&](int ii, int jj) {
                    const int m = jj * mx + ii;
                    return (ii == 0 || jj == 0 || ii == mx - 1 || jj == my - 1) ? g[m] : u[m];
                };
                const double uc = u[n], ue = val(i + 1, j), uw = val(i - 1, j), un = val(i, j + 1), us = val(i, j - 1);
                const double une = val(i + 1, j + 1), unw = val(i - 1, j + 1), use = val(i + 1, j - 1), usw = val(i - 1, j - 1);
                double dux, duy;
                dux = (ue - uc) / hx;  duy = (un + une - us - use) / (4.0 * hy);
                const double De = DD(dux * dux + duy * duy);
                dux = (uc - uw) / hx;  duy = (unw + un - usw - us) / (4.0 * hy);
                const double Dw = DD(dux * dux + duy * duy);
                dux = (ue + une - uw - unw) / (4.0 * hx);  duy = (un - uc) / hy;
                const double Dn = DD(dux * dux + duy * duy);
                dux = (ue + use - uw - usw) / (4.0 * hx);  duy = (uc - us) / hy;
                const double Ds = DD(dux * dux + duy * duy);
                FF[n] = -hyhx * (De * (ue - uc) - Dw * (uc - uw)) - hxhy * (Dn * (un - uc) - Ds * (uc - us));
            }
    }
    // [PETSc] MatFDColoringApply ("ds"), 9 colours of the DMDA BOX stencil; vals in the stencil9 layout
    static double fd_dx(double x) {
        const double eps = 1.4901161193847656e-08, umin = 1.0e-6;
        double dx = x;
        if (fabs(dx) < umin) dx = (dx < 0.0 ? -1.0 : 1.0) * umin;
        return dx * eps;
    }
    void minimal_jacobian_fd(int mx, int my, double q, const double *u, const double *g, const double *F0, double *vals) {
        const int N = mx * my;
        std::vector<double> up(N), Fp(N);
        memset(vals, 0, sizeof(double) * 9 * (size_t)N);
        for (int cj = 0; cj < 3; cj++)
            for (int ci = 0; ci < 3; ci++) {
                for (int n = 0; n < N; n++) {
                    const int j = n / mx, i = n - j * mx;
                    up[n] = (i % 3 == ci && j % 3 == cj) ? u[n] + fd_dx(u[n]) : u[n];
                }
                minimal_function(mx, my, q, up.data(), g, Fp.data());
                for (int n = 0; n < N; n++) {
                    const int j = n / mx, i = n - j * mx;
                    int di = ci - i % 3, dj = cj - j % 3;
                    if (di > 1) di -= 3;
                    if (di < -1) di += 3;
                    if (dj > 1) dj -= 3;
                    if (dj < -1) dj += 3;
                    const int ii = i + di, jj = j + dj;
                    if (ii < 0 || ii >= mx || jj < 0 || jj >= my) continue;
                    vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n] = (Fp[n] - F0[n]) * (1.0 / fd_dx(u[jj * mx + ii]));
                }
            }
    }
    // [PETSc] MatMult and the fused Chebyshev/Jacobi step on the stencil9 matrix
    double row_apply(int mx, int my, const double *vals, const double *u, int n, double *diag) const {
        const int N = mx * my, j = n / mx, i = n - j * mx;
        double Au = 0.0;
        for (int dj = -1; dj <= 1; dj++)
            for (int di = -1; di <= 1; di++) {
                const int s = 3 * (dj + 1) + (di + 1);
                const double a = vals[(size_t)s * N + n];
                if (s == 4 && diag) *diag = a;
                const int ii = i + di, jj = j + dj;
                if (ii >= 0 && ii < mx && jj >= 0 && jj < my) Au += a * u[n + dj * mx + di];
            }
        return Au;
    }
    void stencil9_apply(int mx, int my, const double *vals, const double *x, double *y) {
        for (int n = 0; n < mx * my; n++) y[n] = row_apply(mx, my, vals, x, n, nullptr);
    }
    void stencil9_lin(int mx, int my, const double *vals, const double *u, const double *b, const double *pm1, double ca,
                      double cb, double cg, int jacobi, double *out) {
        const int N = mx * my;
        std::vector<double> o(N);
        for (int n = 0; n < N; n++) {
            double diag = 1.0;
            double r = (b ? b[n] : 0.0) - row_apply(mx, my, vals, u, n, &diag);
            if (jacobi) r /= diag;
            o[n] = cb * u[n] + cg * r + (pm1 ? ca * pm1[n] : 0.0);
        }
        memcpy(out, o.data(), sizeof(double) * N);
    }
    double stencil9_gershgorin(int mx, int my, const double *vals, double *) {
        const int N = mx * my;
        double best = 0.0;
        for (int n = 0; n < N; n++) {
            const int j = n / mx, i = n - j * mx;
            double s = 0.0;
            for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                    const int ii = i + di, jj = j + dj;
                    if (ii >= 0 && ii < mx && jj >= 0 && jj < my) s += fabs(vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n]);
                }
            best = std::max(best, s / fabs(vals[(size_t)4 * N + n]));
        }
        return best;
    }
    // [PETSc] DMDA Q1 interpolation (ratio 2), its transpose, and injection
    void inject2d(int cmx, int cmy, const double *uf, double *uc) {
        const int fmx = 2 * cmx - 1;
        for (int J = 0; J < cmy; J++)
            for (int I = 0; I < cmx; I++) uc[J * cmx + I] = uf[(2 * J) * fmx + 2 * I];
    }
    void restrict2d(int fmx, int fmy, const double *rf, double *bc) {
        const int cmx = (fmx - 1) / 2 + 1, cmy = (fmy - 1) / 2 + 1;
        for (int J = 0; J < cmy; J++)
            for (int I = 0; I < cmx; I++) {
                double s = 0.0;
                for (int dj = -1; dj <= 1; dj++)
                    for (int di = -1; di <= 1; di++) {
                        const int i = 2 * I + di, j = 2 * J + dj;
                        if (i < 0 || i >= fmx || j < 0 || j >= fmy) continue;
                        s += (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * rf[j * fmx + i];
                    }
                bc[J * cmx + I] = s;
            }
    }
    void prolong_add2d(int fmx, int fmy, const double *xc, double *xf) {
        const int cmx = (fmx - 1) / 2 + 1;
        for (int j = 0; j < fmy; j++)
            for (int i = 0; i < fmx; i++) {
                const int I0 = i / 2, J0 = j / 2, oi = i & 1, oj = j & 1;
                double s = 0.0;
                for (int dj = 0; dj <= oj; dj++)
                    for (int di = 0; di <= oi; di++) s += xc[(J0 + dj) * cmx + (I0 + di)];
                xf[j * fmx + i] += s * (oi ? 0.5 : 1.0) * (oj ? 0.5 : 1.0);
            }
    }
    void dense_matvec(int n, const double *Ainv, const double *b, double *x) {
        std::vector<double> o(n);
        for (int r = 0; r < n; r++) {
            double s = 0.0;
            for (int c = 0; c < n; c++) s += Ainv[(size_t)r * n + c] * b[c];
            o[r] = s;
        }
        memcpy(x, o.data(), sizeof(double) * n);
    }
};
