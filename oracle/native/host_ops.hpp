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

#include "ts_solver.hpp"      // p4b::nk::PatternOpts (the options struct the pattern operations take)

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
    bool cgs = false;                                           // -gmres_cgs of the harness (p4b_tune("gmres_cgs") in the library)
    bool gmres_cgs() const { return cgs; }
    void mdot(size_t n, int k, const double *const *X, const double *y, double *res) { for (int i = 0; i < k; i++) res[i] = dot(n, X[i], y); }
    double norminf(size_t n, const double *x) { double m = 0; for (size_t i = 0; i < n; i++) m = std::max(m, fabs(x[i])); return m; }
    void axpy(size_t n, double a, const double *x, double *y) { for (size_t i = 0; i < n; i++) y[i] += a * x[i]; }
    void aypx(size_t n, double a, const double *x, double *y) { for (size_t i = 0; i < n; i++) y[i] = x[i] + a * y[i]; }
    void axpby(size_t n, double a, const double *x, double b, const double *y, double *out) {
        for (size_t i = 0; i < n; i++) out[i] = (x ? a * x[i] : 0.0) + (y ? b * y[i] : 0.0);
    }
    void copy(size_t n, const double *x, double *y) { memmove(y, x, sizeof(double) * n); }
    void set(size_t n, double a, double *y) { for (size_t i = 0; i < n; i++) y[i] = a; }
    void band_inverse(int n, int bw, const std::vector<double> &B, double *Ainv) {
        std::vector<double> inv;
        p4b::nk::band_inverse_host(B, n, bw, &inv);
        memcpy(Ainv, inv.data(), sizeof(double) * inv.size());
    }
    void user_monitor(int, int, int, double, int, const double *) {}
    bool verify_converged(int, int, const double *, const double *) { return true; }
    void set_linearisation(const double *) {}
    void set_time(double) {}
    // TS step monitor: the stand-in's own registration (p4b_standin.cpp) or none (nk_host_main.cpp)
    static int (*&ts_fn())(void *, int, double, const double *, size_t) { static int (*f)(void *, int, double, const double *, size_t) = nullptr; return f; }
    static void *&ts_user() { static void *u = nullptr; return u; }
    void ts_step(int k, double t, const double *Y, size_t n) {
        if (ts_fn() && !err && ts_fn()(ts_user(), k, t, Y, n)) err = 66;
    }
    static double &ts_h() { static double h = 0.0; return h; }
    void set_step_size(double h) { ts_h() = h; }

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
    // [PETSc] MatFDColoringApply, default differencing "wp": 9 colours of the DMDA BOX stencil, ONE step for every
    // column, h = sqrt(DBL_EPSILON) sqrt(1 + ||u||_2) (pinned by minimal.test1, oracle/minimal_solver_oracle.py);
    // vals in the stencil9 layout
    double fd_step(size_t n, const double *u) { return 1.4901161193847656e-08 * sqrt(1.0 + norm2(n, u)); }
    // the rows c/ch6/poissonfunctions.c:152-193 inserts on the unit square with cx = cy = 1 (what c/ch7/minimal.c:142-145
    // registers as its Jacobian), stencil9 layout
    void poisson_stencil9(int mx, int my, double *vals) {
        const int N = mx * my;
        const double hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1), scx = hy / hx, scy = hx / hy;
        memset(vals, 0, sizeof(double) * 9 * (size_t)N);
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const int n = j * mx + i;
                vals[(size_t)4 * N + n] = 2.0 * (scx + scy);
                if (i == 0 || i == mx - 1 || j == 0 || j == my - 1) continue;
                if (i - 1 > 0) vals[(size_t)3 * N + n] = -scx;
                if (i + 1 < mx - 1) vals[(size_t)5 * N + n] = -scx;
                if (j - 1 > 0) vals[(size_t)1 * N + n] = -scy;
                if (j + 1 < my - 1) vals[(size_t)7 * N + n] = -scy;
            }
    }
    void minimal_jacobian_fd(int mx, int my, double q, const double *u, const double *g, const double *F0, double *vals) {
        const int N = mx * my;
        std::vector<double> up(N), Fp(N);
        const double h = fd_step((size_t)N, u), vscale = 1.0 / h;
        memset(vals, 0, sizeof(double) * 9 * (size_t)N);
        for (int cj = 0; cj < 3; cj++)
            for (int ci = 0; ci < 3; ci++) {
                for (int n = 0; n < N; n++) {
                    const int j = n / mx, i = n - j * mx;
                    up[n] = (i % 3 == ci && j % 3 == cj) ? u[n] + h : u[n];
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
                    vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n] = (Fp[n] - F0[n]) * vscale;
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
    // ---- c/ch5/pattern.c ----------------------------------------------------------------------------------
    typedef p4b::nk::PatternOpts PO;
    // [PETSc] TSErrorWeightedNorm2 (sum of squares; the caller divides by n and takes the root)
    double wrms2(size_t n, const double *x, const double *y, double atol, double rtol) {
        double s = 0.0;
        for (size_t i = 0; i < n; i++) {
            const double e = (x[i] - y[i]) / (atol + rtol * std::max(fabs(x[i]), fabs(y[i])));
            s += e * e;
        }
        return s;
    }
    // pattern.c:146-179  InitialState without noise
    void pattern_initial_state(int mx, int my, double L, double *Y) {
        const double PI = 3.14159265358979323846264338327950288, ledge = (L - 0.5) / 2.0, redge = L - ledge;
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const double x = i * (L / mx), y = j * (L / my);
                double v = 0.0;
                if (x >= ledge && x <= redge && y >= ledge && y <= redge) {
                    const double sx = sin(4.0 * PI * x), sy = sin(4.0 * PI * y);
                    v = 0.5 * sx * sx * sy * sy;
                }
                Y[2 * (j * mx + i)] = 1.0 - 2.0 * v;
                Y[2 * (j * mx + i) + 1] = v;
            }
    }
    static void lap9(int m, const double *Y, int i, int j, double *lu, double *lv) {
        auto at = [&](int ii, int jj, int c) { return Y[2 * (((jj + m) % m) * m + (ii + m) % m) + c]; };
        double l[2];
        for (int c = 0; c < 2; c++)
            l[c] = at(i - 1, j + 1, c) + 4.0 * at(i, j + 1, c) + at(i + 1, j + 1, c) + 4.0 * at(i - 1, j, c) - 20.0 * at(i, j, c) +
                   4.0 * at(i + 1, j, c) + at(i - 1, j - 1, c) + 4.0 * at(i, j - 1, c) + at(i + 1, j - 1, c);
        *lu = l[0];
        *lv = l[1];
    }
    // pattern.c:242-267  FormIFunctionLocal
    void pattern_ifunction(int m, const PO &o, const double *Y, const double *Ydot, double *F) {
        const double h = o.L / m, Cu = o.Du / (6.0 * h * h), Cv = o.Dv / (6.0 * h * h);
        std::vector<double> out((size_t)2 * m * m);
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) {
                double lu, lv;
                lap9(m, Y, i, j, &lu, &lv);
                const int k = 2 * (j * m + i);
                out[k] = Ydot[k] - Cu * lu;
                out[k + 1] = Ydot[k + 1] - Cv * lv;
            }
        memcpy(F, out.data(), sizeof(double) * out.size());
    }
    // pattern.c:185-199  FormRHSFunctionLocal
    void pattern_rhsfunction(int m, const PO &o, const double *Y, double *G) {
        for (int k = 0; k < m * m; k++) {
            const double u = Y[2 * k], v = Y[2 * k + 1], uv2 = u * v * v;
            G[2 * k] = -uv2 + o.phi * (1.0 - u);
            G[2 * k + 1] = uv2 - (o.phi + o.kappa) * v;
        }
    }
    // J X = shift X - C L9 X - G'(Y) X   (pattern.c:274-318 minus :202-236; Y == nullptr drops G')
    void jac_rows(int m, const PO &o, double shift, const double *Y, const double *X, int i, int j, double *Ju, double *Jv, double *du,
                  double *dv, double *g01, double *g10) const {
        const double h = o.L / m, Cu = o.Du / (6.0 * h * h), Cv = o.Dv / (6.0 * h * h);
        const int k = 2 * (j * m + i);
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
        if (Y) {
            const double u = Y[k], v = Y[k + 1];
            a00 = -v * v - o.phi; a01 = -2.0 * u * v; a10 = v * v; a11 = 2.0 * u * v - (o.phi + o.kappa);
        }
        *du = shift + 20.0 * Cu - a00;
        *dv = shift + 20.0 * Cv - a11;
        *g01 = a01;
        *g10 = a10;
        if (X) {
            double lu, lv;
            lap9(m, X, i, j, &lu, &lv);
            *Ju = shift * X[k] - Cu * lu - (a00 * X[k] + a01 * X[k + 1]);
            *Jv = shift * X[k + 1] - Cv * lv - (a10 * X[k] + a11 * X[k + 1]);
        }
    }
    void pattern_jac_apply(int m, const PO &o, double shift, const double *Y, const double *X, double *out) {
        std::vector<double> r((size_t)2 * m * m);
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) {
                double du, dv, g01, g10;
                jac_rows(m, o, shift, Y, X, i, j, &r[2 * (j * m + i)], &r[2 * (j * m + i) + 1], &du, &dv, &g01, &g10);
            }
        memcpy(out, r.data(), sizeof(double) * r.size());
    }
    void pattern_jac_lin(int m, const PO &o, double shift, const double *Y, const double *X, const double *b, const double *pm1,
                         double ca, double cb, double cg, int jacobi, double *out) {
        std::vector<double> r((size_t)2 * m * m);
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) {
                const int k = 2 * (j * m + i);
                double Ju = 0.0, Jv = 0.0, du, dv, g01, g10;
                jac_rows(m, o, shift, Y, X, i, j, &Ju, &Jv, &du, &dv, &g01, &g10);
                double ru = (b ? b[k] : 0.0) - Ju, rv = (b ? b[k + 1] : 0.0) - Jv;
                if (jacobi) { ru /= du; rv /= dv; }
                r[k] = cb * X[k] + cg * ru + (pm1 ? ca * pm1[k] : 0.0);
                r[k + 1] = cb * X[k + 1] + cg * rv + (pm1 ? ca * pm1[k + 1] : 0.0);
            }
        memcpy(out, r.data(), sizeof(double) * r.size());
    }
    double pattern_jac_gershgorin(int m, const PO &o, double shift, const double *Y, double *) {
        const double h = o.L / m, Cu = o.Du / (6.0 * h * h), Cv = o.Dv / (6.0 * h * h);
        double best = 0.0;
        for (int j = 0; j < m; j++)
            for (int i = 0; i < m; i++) {
                double Ju = 0.0, Jv = 0.0, du, dv, g01, g10;
                jac_rows(m, o, shift, Y, nullptr, i, j, &Ju, &Jv, &du, &dv, &g01, &g10);
                best = std::max(best, (fabs(du) + 20.0 * Cu + fabs(g01)) / fabs(du));
                best = std::max(best, (fabs(dv) + 20.0 * Cv + fabs(g10)) / fabs(dv));
            }
        return best;
    }
    // [PETSc] periodic DMDA Q1 interpolation (ratio 2), its transpose, injection; (u, v) interleaved
    void pattern_restrict(int Mx, int My, const double *rf, double *bc) {
        const int fx = 2 * Mx, fy = 2 * My;
        for (int J = 0; J < My; J++)
            for (int I = 0; I < Mx; I++)
                for (int c = 0; c < 2; c++) {
                    double s = 0.0;
                    for (int dj = -1; dj <= 1; dj++)
                        for (int di = -1; di <= 1; di++)
                            s += (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * rf[2 * (((2 * J + dj + fy) % fy) * fx + (2 * I + di + fx) % fx) + c];
                    bc[2 * (J * Mx + I) + c] = s;
                }
    }
    void pattern_prolong_add(int Mx, int My, const double *xc, double *xf) {
        const int fx = 2 * Mx, fy = 2 * My;
        for (int j = 0; j < fy; j++)
            for (int i = 0; i < fx; i++)
                for (int c = 0; c < 2; c++) {
                    const int I0 = i / 2, J0 = j / 2, I1 = (i & 1) ? (I0 + 1) % Mx : I0, J1 = (j & 1) ? (J0 + 1) % My : J0;
                    xf[2 * (j * fx + i) + c] += 0.25 * ((xc[2 * (J0 * Mx + I0) + c] + xc[2 * (J0 * Mx + I1) + c]) +
                                                        (xc[2 * (J1 * Mx + I0) + c] + xc[2 * (J1 * Mx + I1) + c]));
                }
    }
    void pattern_inject(int Mx, int My, const double *yf, double *yc) {
        const int fx = 2 * Mx;
        for (int J = 0; J < My; J++)
            for (int I = 0; I < Mx; I++)
                for (int c = 0; c < 2; c++) yc[2 * (J * Mx + I) + c] = yf[2 * ((2 * J) * fx + 2 * I) + c];
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

// The residual supplied as a host callback (the FormFunctionLocal contract; include/p4b200.h p4b_residual2d_fn): the CPU
// counterpart of CallbackOps in p4pdes_b200/csrc/nk_device.cu.
struct HostCallbackOps : HostOps {
    int (*fn)(void *user, int mx, int my, const double *u, double *F) = nullptr;
    void *user = nullptr;
    int (*mon)(void *user, int mx, int my, int its, double fnorm, int tablevel, const double *u) = nullptr;
    long long callbacks = 0;
    void user_monitor(int mx, int my, int its, double fnorm, int tablevel, const double *u) {
        if (!mon || err) return;
        std::vector<double> uu(u, u + (size_t)mx * my);
        if (mon(user, mx, my, its, fnorm, tablevel, uu.data())) err = 66;
    }
    void minimal_sample(int, int, int, double, double, double *) {}
    void minimal_function(int mx, int my, double, const double *u, const double *, double *F) {
        callbacks++;
        std::vector<double> uu(u, u + (size_t)mx * my), ff((size_t)mx * my);
        if (fn(user, mx, my, uu.data(), ff.data()) && !err) err = 65;
        memcpy(F, ff.data(), sizeof(double) * ff.size());
    }
    void minimal_jacobian_fd(int mx, int my, double q, const double *u, const double *g, const double *F0, double *vals) {
        const int N = mx * my;
        std::vector<double> up(N), Fp(N);
        const double h = fd_step((size_t)N, u), vscale = 1.0 / h;
        memset(vals, 0, sizeof(double) * 9 * (size_t)N);
        for (int cj = 0; cj < 3; cj++)
            for (int ci = 0; ci < 3; ci++) {
                for (int n = 0; n < N; n++) {
                    const int j = n / mx, i = n - j * mx;
                    up[n] = (i % 3 == ci && j % 3 == cj) ? u[n] + h : u[n];
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
                    vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n] = (Fp[n] - F0[n]) * vscale;
                }
            }
    }
};

// F(t, Y, Ydot) and G(t, Y) supplied as host callbacks, no Jacobian (include/p4b200.h p4b_ts2d_solve): the CPU counterpart of
// CallbackPatternOps in p4pdes_b200/csrc/nk_device.cu -- the stage operator is the differenced residual ([PETSc] MatMFFD "wp").
struct HostCallbackPatternOps : HostOps {
    int (*ifn)(void *user, int m, double t, const double *Y, const double *Ydot, double *F) = nullptr;
    int (*gfn)(void *user, int m, double t, const double *Y, double *G) = nullptr;
    void *user = nullptr;
    const double *lin = nullptr, *r0_for = nullptr;
    double r0_shift = 0.0;
    bool r0_rhs = false;
    std::vector<double> wY, wD, wF, wR0;
    long long callbacks = 0;
    size_t nfield = 0;                     // != 0: a field of that many doubles (p4b_ts_solve_callbacks), m is 0
    size_t fsize(int m) const { return nfield ? nfield : (size_t)2 * m * m; }
    void pattern_ifunction(int m, const PO &, const double *Y, const double *Ydot, double *F) {
        const size_t n = fsize(m);
        std::vector<double> y(Y, Y + n), d(Ydot, Ydot + n), f(n);
        callbacks++;
        if (ifn(user, m, tcur, y.data(), d.data(), f.data()) && !err) err = 65;
        memcpy(F, f.data(), sizeof(double) * n);
    }
    void pattern_rhsfunction(int m, const PO &, const double *Y, double *G) {
        const size_t n = fsize(m);
        std::vector<double> y(Y, Y + n), g(n);
        callbacks++;
        if (gfn(user, m, tcur, y.data(), g.data()) && !err) err = 65;
        memcpy(G, g.data(), sizeof(double) * n);
    }
    void set_linearisation(const double *Y) { lin = Y; r0_for = nullptr; }
    double tcur = 0.0;
    void set_time(double t) { tcur = t; }
    void resid(int m, const PO &o, double shift, bool rhs, const double *W, double *out) {
        const size_t n = fsize(m);
        wD.resize(n);
        for (size_t i = 0; i < n; i++) wD[i] = shift * W[i];
        pattern_ifunction(m, o, W, wD.data(), out);
        if (rhs) { wF.resize(n); pattern_rhsfunction(m, o, W, wF.data()); axpy(n, -1.0, wF.data(), out); }
    }
    void pattern_jac_apply(int m, const PO &o, double shift, const double *Y, const double *X, double *out) {
        const size_t n = fsize(m);
        const bool rhs = Y != nullptr;
        if (!lin) { if (!err) err = 68; return; }
        if (r0_for != lin || r0_shift != shift || r0_rhs != rhs) {
            wR0.resize(n);
            resid(m, o, shift, rhs, lin, wR0.data());
            r0_for = lin; r0_shift = shift; r0_rhs = rhs;
        }
        const double xn = norm2(n, X);
        if (xn == 0.0) { set(n, 0.0, out); return; }
        const double h = 1.4901161193847656e-08 * sqrt(1.0 + norm2(n, lin)) / xn;
        wY.resize(n);
        axpby(n, 1.0, lin, h, X, wY.data());
        resid(m, o, shift, rhs, wY.data(), out);
        axpby(n, 1.0 / h, out, -1.0 / h, wR0.data(), out);
    }
};

// c/ch5/heat.c with plain loops: the CPU counterpart of HeatOps in p4pdes_b200/csrc/nk_device.cu (heat.c:141-163: G; :166-208:
// the rows of dG/du, applied without being stored)
struct HostHeatOps : HostOps {
    int mx = 0, my = 0;
    double D0 = 1.0;
    void heat(int mode, double shift, const double *u, double *out) {
        const double hx = 1.0 / (mx - 1), hy = 1.0 / my, PI = 3.14159265358979323846;
        for (int j = 0; j < my; j++)
            for (int i = 0; i < mx; i++) {
                const int n = j * mx + i, js = j == 0 ? my - 1 : j - 1, jn = j == my - 1 ? 0 : j + 1;
                const double x = hx * i, y = hy * j, c = u[n];
                double ul = i == 0 ? u[n + 1] : u[n - 1];
                if (i == 0 && mode == 0) ul += 2.0 * hx * sin(6.0 * PI * y);
                const double ur = i == mx - 1 ? u[n - 1] : u[n + 1];
                const double lap = (ul - 2.0 * c + ur) / (hx * hx) + (u[js * mx + i] - 2.0 * c + u[jn * mx + i]) / (hy * hy);
                out[n] = mode == 0 ? D0 * lap + 3.0 * exp(-25.0 * (x - 0.6) * (x - 0.6)) * sin(2.0 * PI * y) : shift * c - D0 * lap;
            }
    }
    void pattern_ifunction(int, const PO &, const double *, const double *Ydot, double *F) { copy((size_t)mx * my, Ydot, F); }
    void pattern_rhsfunction(int, const PO &, const double *Y, double *G) { heat(0, 0.0, Y, G); }
    void pattern_jac_apply(int, const PO &, double shift, const double *Y, const double *X, double *out) {
        if (Y) heat(1, shift, X, out);
        else axpby((size_t)mx * my, shift, X, 0.0, nullptr, out);
    }
};
