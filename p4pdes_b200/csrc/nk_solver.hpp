// nk_solver.hpp -- host logic of the Newton-Krylov-multigrid solve around c/ch7/minimal.c, in C++ over an abstract
// set of vector/kernel operations (`Ops`).
//
// What PETSc does inside SNESSolve for `./minimal -snes_fd_color -pc_type mg [-snes_grid_sequence k]`
// (c/ch8/cluster.sh:70, c/ch7/makefile:15-25), statement for statement the logic of
// p4pdes_b200/minimal.py, which is pinned on the goldens through oracle/minimal_solver_oracle.py:
//   Newton + cubic backtracking line search   [PETSc] SNESSolve_NEWTONLS, SNESLineSearchApply_BT, SNESConvergedDefault
//   GMRES(restart), left preconditioned / CG  [PETSc] KSPGMRES / KSPCG
//   V cycle on assembled level Jacobians      [PETSc] PCMG: Chebyshev + Jacobi smoothing, R = P^T, PCLU on the base grid
//   level Jacobians                           [PETSc] -snes_fd_color on every level at the injected iterate
//   grid sequencing                           [PETSc] -snes_grid_sequence: DMRefine + Q1 interpolation of the iterate
//
// The library instantiates it with DeviceOps (nk_device.cu: every operation is a CUDA kernel of this repo, vectors
// live in HBM).  oracle/native/ instantiates the SAME template with plain C++ loops to check this file's control flow
// on a machine without a GPU -- test infrastructure, never linked into libp4b200.so.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

namespace p4b {
namespace nk {

enum { KSP_GMRES = 0, KSP_CG = 1 };
enum { PC_NONE = 0, PC_MG = 1 };
enum { SNES_CONVERGED_FNORM_ABS = 2, SNES_CONVERGED_FNORM_RELATIVE = 3, SNES_CONVERGED_SNORM_RELATIVE = 4,
       SNES_DIVERGED_MAX_IT = -5, SNES_DIVERGED_LINE_SEARCH = -6, SNES_DIVERGED_FNORM_NAN = -4 };
constexpr int MAX_STAGES = 16;
constexpr int MAX_NEWTON = 64;

struct MinimalOpts {
    int problem;                 // 0 tent, 1 catenoid (minimal.c:50-52)
    double q, catenoid_c, tent_H;
    int exact_init;
    int grid_x, grid_y, refine, grid_sequence;
    int ksp_type;                // KSP_GMRES | KSP_CG
    double ksp_rtol;
    int ksp_max_it, gmres_restart;
    int pc_type;                 // PC_NONE | PC_MG
    int mg_levels, smooth_its;
    double snes_rtol, snes_stol, snes_atol;
    int snes_max_it;
    int snes_monitor;            // 0 off, 1 full precision, 2 short
    int snes_converged_reason, ksp_converged_reason;
    int mf_operator;             // -snes_mf_operator: J v by differencing the residual ([PETSc] MatMFFD "wp"); the assembled
                                 // (FD-coloured) Jacobian is the preconditioner's matrix only
    int jacobian;                // which matrix is assembled: 0 the FD-coloured Jacobian of the residual (-snes_fd_color),
                                 // 1 the one minimal.c REGISTERS, Poisson2DJacobianLocal (minimal.c:142-145, "ONLY
                                 // APPROXIMATE"): Newton's matrix when mf_operator == 0 ([PETSc] without -snes_fd_color),
                                 // the preconditioner's under -snes_mf_operator ([PETSc]'s choice there)
};

inline void default_opts(MinimalOpts *o) {
    memset(o, 0, sizeof *o);
    o->problem = 1; o->q = -0.5; o->catenoid_c = 1.1; o->tent_H = 1.0;
    o->grid_x = o->grid_y = 3;
    o->ksp_type = KSP_GMRES; o->ksp_rtol = 1.0e-5; o->ksp_max_it = 10000; o->gmres_restart = 30;
    o->pc_type = PC_MG; o->smooth_its = 2;
    o->snes_rtol = 1.0e-8; o->snes_stol = 1.0e-8; o->snes_atol = 1.0e-50; o->snes_max_it = 50;
}

struct StageResult {
    int mx, my, its, reason, nksp;
    int ksp_its[MAX_NEWTON];
    double lambda[MAX_NEWTON];
    double fnorm[MAX_NEWTON + 1];
};

struct MinimalResult {
    int mx, my, nstages;
    StageResult stage[MAX_STAGES];
    double errinf;               // |u - uexact|_inf for the catenoid with q = -1/2, else -1
    int error;                   // 0, or the first error an operation reported
    char errmsg[256];
};

typedef void (*LineFn)(const char *line, void *ctx);

struct Printer {
    LineFn fn;
    void *ctx;
    void out(const char *fmt, ...) const {
        if (!fn) return;
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        fn(buf, ctx);
    }
};

inline const char *snes_reason_name(int r) {
    switch (r) {
        case SNES_CONVERGED_FNORM_ABS: return "CONVERGED_FNORM_ABS";
        case SNES_CONVERGED_FNORM_RELATIVE: return "CONVERGED_FNORM_RELATIVE";
        case SNES_CONVERGED_SNORM_RELATIVE: return "CONVERGED_SNORM_RELATIVE";
        case SNES_DIVERGED_MAX_IT: return "DIVERGED_MAX_IT";
        case SNES_DIVERGED_LINE_SEARCH: return "DIVERGED_LINE_SEARCH";
        case SNES_DIVERGED_FNORM_NAN: return "DIVERGED_FNORM_NAN";
    }
    return "?";
}

// how -snes_monitor_short prints norms ([PETSc] SNESMonitorDefaultShort): %g above 1e-9, %5.3e down to 1e-11, then
// "< 1.e-11"  (c/ch7/output/minimal.test1:6 "1.772e-10")
inline std::string g6(double v) {
    char a[64];
    if (v > 1.0e-9) snprintf(a, sizeof a, "%g", v);
    else if (v > 1.0e-11) snprintf(a, sizeof a, "%5.3e", v);
    else snprintf(a, sizeof a, "< 1.e-11");
    return a;
}

// ---------------------------------------------------------------------------------------------------------
// dense inverse of a (small) stencil9 matrix on the host: banded LU without pivoting (the Jacobians here are
// diagonally dominant M-matrices plus finite-difference noise), then n banded solves.  [PETSc] PCLU on the base grid.
// vals: 9 planes of n = mx*my doubles (host copy); Ainv: n*n doubles, row-major.
// ---------------------------------------------------------------------------------------------------------
// Banded LU (no pivoting: the base-grid Jacobian is an M-matrix up to the differencing noise) of the 9-point matrix in
// stencil9 layout.  Band storage B[r*W + (c - r + bw)], bw = mx + 1, W = 2 bw + 1; L unit lower.  O(n bw^2): about a
// millisecond on the host for the 33 x 33 base grid of c/ch8/cluster.sh:70.
inline int stencil9_band_lu(const double *vals, int mx, int my, std::vector<double> *Bout) {
    const int n = mx * my, bw = mx + 1, W = 2 * bw + 1;
    std::vector<double> &B = *Bout;
    B.assign((size_t)n * W, 0.0);
    for (int j = 0; j < my; j++)
        for (int i = 0; i < mx; i++) {
            const int r = j * mx + i;
            for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                    const int ii = i + di, jj = j + dj;
                    if (ii < 0 || ii >= mx || jj < 0 || jj >= my) continue;
                    const int c = jj * mx + ii;
                    B[(size_t)r * W + (c - r + bw)] = vals[(size_t)(3 * (dj + 1) + (di + 1)) * n + r];
                }
        }
    for (int k = 0; k < n; k++) {                              // LU in place (L unit lower)
        const double piv = B[(size_t)k * W + bw];
        if (piv == 0.0 || piv != piv) return 1;
        const int rmax = std::min(n - 1, k + bw);
        for (int r = k + 1; r <= rmax; r++) {
            double &l = B[(size_t)r * W + (k - r + bw)];
            if (l == 0.0) continue;
            l /= piv;
            const int cmax = std::min(n - 1, k + bw);
            for (int c = k + 1; c <= cmax; c++) B[(size_t)r * W + (c - r + bw)] -= l * B[(size_t)k * W + (c - k + bw)];
        }
    }
    return 0;
}
// A^-1 (dense, row-major) from the band factors: n independent column solves.  This is the host form (the CPU
// instantiation of the solver templates); the library does the same on the device, one thread per column
// (assembled.cu band_inverse_kernel) -- on the host it took ~50 ms per Newton step at n = 1089 and was, 30 Newton steps
// over, most of the 2049^2 cluster run.
inline void band_inverse_host(const std::vector<double> &B, int n, int bw, std::vector<double> *Ainv) {
    const int W = 2 * bw + 1;
    Ainv->assign((size_t)n * n, 0.0);
    std::vector<double> x(n);
    for (int col = 0; col < n; col++) {
        std::fill(x.begin(), x.end(), 0.0);
        x[col] = 1.0;
        for (int r = col + 1; r < n; r++) {                    // forward: L y = e_col (y_r = 0 for r < col)
            double s = 0.0;
            const int c0 = std::max(col, r - bw);
            for (int c = c0; c < r; c++) s += B[(size_t)r * W + (c - r + bw)] * x[c];
            x[r] -= s;
        }
        for (int r = n - 1; r >= 0; r--) {                     // backward: U x = y
            double s = x[r];
            const int c1 = std::min(n - 1, r + bw);
            for (int c = r + 1; c <= c1; c++) s -= B[(size_t)r * W + (c - r + bw)] * x[c];
            x[r] = s / B[(size_t)r * W + bw];
        }
        for (int r = 0; r < n; r++) (*Ainv)[(size_t)r * n + col] = x[r];
    }
}
// both steps on the host (kept for callers that want the inverse there)
inline int stencil9_inverse(const double *vals, int mx, int my, std::vector<double> *Ainv) {
    std::vector<double> B;
    if (stencil9_band_lu(vals, mx, my, &B)) return 1;
    band_inverse_host(B, mx * my, mx + 1, Ainv);
    return 0;
}
// base-grid inverse into Ops memory: band LU on the host, the n column solves wherever the operations run
template <class Ops>
inline int stencil9_inverse_ops(Ops *ops, const double *vals_dev, int mx, int my, double *Ainv_dev) {
    const size_t n = (size_t)mx * my;
    std::vector<double> hv(9 * n), B;
    ops->to_host(vals_dev, hv.data(), 9 * n);
    if (stencil9_band_lu(hv.data(), mx, my, &B)) return 62;
    ops->band_inverse((int)n, mx + 1, B, Ainv_dev);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// one grid of the hierarchy
// ---------------------------------------------------------------------------------------------------------
template <class Ops>
struct Level {
    Ops *ops = nullptr;
    int mx = 0, my = 0;
    size_t n = 0;
    double *g = nullptr, *vals = nullptr, *u = nullptr, *F = nullptr, *x = nullptr, *b = nullptr, *t = nullptr;
    double scale = 0.0;
    std::vector<double> omega;
    bool poisson = false, poisson_ready = false;

    void create(Ops *o, int mx_, int my_, const MinimalOpts &opt) {
        ops = o; mx = mx_; my = my_; n = (size_t)mx * my;
        poisson = opt.jacobian == 1;
        poisson_ready = false;
        g = ops->alloc(n); vals = ops->alloc(9 * n); u = ops->alloc(n); F = ops->alloc(n);
        x = ops->alloc(n); b = ops->alloc(n); t = ops->alloc(n);
        ops->minimal_sample(mx, my, opt.problem, opt.tent_H, opt.catenoid_c, g);
    }
    void destroy() {
        if (!ops) return;
        for (double *p : {g, vals, u, F, x, b, t}) ops->release(p);
        ops = nullptr;
    }
    void assemble(double q, bool F_known) {
        if (poisson) {                                         // the same at every iterate: filled once
            if (!poisson_ready) ops->poisson_stencil9(mx, my, vals);
            poisson_ready = true;
            return;
        }
        if (!F_known) ops->minimal_function(mx, my, q, u, g, F);
        ops->minimal_jacobian_fd(mx, my, q, u, g, F, vals);
    }
    // [PETSc] KSPSolve_Chebyshev, first kind, targets (0.1, 1.1) x the Gershgorin bound of D^-1 A
    void set_smoother(int its) {
        const double lam = ops->stencil9_gershgorin(mx, my, vals, t);
        const double emin = 0.1 * lam, emax = 1.1 * lam;
        scale = 2.0 / (emax + emin);
        const double alpha = 1.0 - scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
        double cm1 = 1.0, ck = mu;
        omega.clear();
        for (int i = 1; i < its; i++) {
            const double cp1 = 2.0 * mu * ck - cm1;
            omega.push_back(omegaprod * ck / cp1);
            cm1 = ck;
            ck = cp1;
        }
    }
    void mult(const double *in, double *out) { ops->stencil9_apply(mx, my, vals, in, out); }
};

// ---------------------------------------------------------------------------------------------------------
// multigrid preconditioner on the assembled level Jacobians; lev[0] is the finest
// ---------------------------------------------------------------------------------------------------------
template <class Ops>
struct AssembledMG {
    Ops *ops;
    std::vector<Level<Ops>> *lev;
    int its;
    double *Ainv = nullptr;
    int n0 = 0;

    int setup(double q) {
        std::vector<Level<Ops>> &L = *lev;
        if (L[0].poisson && L[0].poisson_ready && Ainv && n0 == (int)L.back().n) return 0;
        for (size_t l = 0; l < L.size(); l++) {
            if (l > 0) ops->inject2d(L[l].mx, L[l].my, L[l - 1].u, L[l].u);
            L[l].assemble(q, l == 0);
            if (l + 1 < L.size()) L[l].set_smoother(its);
        }
        Level<Ops> &C = L.back();
        if (C.n > 4225) return 61;                            // base grid larger than 65 x 65: refuse the dense solve
        if (!Ainv || n0 != (int)C.n) {
            if (Ainv) ops->release(Ainv);
            Ainv = ops->alloc(C.n * C.n);
            n0 = (int)C.n;
        }
        return stencil9_inverse_ops(ops, C.vals, C.mx, C.my, Ainv);
    }
    void destroy() {
        if (Ainv) ops->release(Ainv);
        Ainv = nullptr;
    }
    void smooth(Level<Ops> &L, bool zero_guess) {
        if (its <= 0) {
            if (zero_guess) ops->set(L.n, 0.0, L.x);
            return;
        }
        double *pm1 = L.x, *pk = L.t;
        if (zero_guess) ops->set(L.n, 0.0, pm1);
        ops->stencil9_lin(L.mx, L.my, L.vals, pm1, L.b, nullptr, 0.0, 1.0, L.scale, 1, pk);
        for (int i = 1; i < its; i++) {
            const double w = L.omega[i - 1];
            ops->stencil9_lin(L.mx, L.my, L.vals, pk, L.b, pm1, 1.0 - w, w, w * L.scale, 1, pm1);
            std::swap(pm1, pk);
        }
        if (pk != L.x) std::swap(L.x, L.t);
    }
    void cycle(size_t l, bool zero_guess) {
        std::vector<Level<Ops>> &LV = *lev;
        Level<Ops> &L = LV[l];
        if (l + 1 == LV.size()) {
            ops->dense_matvec((int)L.n, Ainv, L.b, L.x);
            return;
        }
        Level<Ops> &C = LV[l + 1];
        smooth(L, zero_guess);
        ops->stencil9_lin(L.mx, L.my, L.vals, L.x, L.b, nullptr, 0.0, 0.0, 1.0, 0, L.t);      // t = b - A x
        ops->restrict2d(L.mx, L.my, L.t, C.b);
        cycle(l + 1, true);
        ops->prolong_add2d(L.mx, L.my, C.x, L.x);
        smooth(L, false);
    }
    void apply(const double *r, double *z) {
        Level<Ops> &L = (*lev)[0];
        ops->copy(L.n, r, L.b);
        cycle(0, true);
        ops->copy(L.n, L.x, z);
    }
};

// the preconditioner a Krylov solve is handed: multigrid, the dense base-grid solve alone, or nothing
template <class Ops>
struct Precond {
    Ops *ops;
    AssembledMG<Ops> *mg = nullptr;
    const double *dense = nullptr;
    int n = 0;
    void apply(const double *r, double *z) {
        if (mg) mg->apply(r, z);
        else if (dense) ops->dense_matvec(n, dense, r, z);
        else ops->copy((size_t)n, r, z);
    }
};

// ---------------------------------------------------------------------------------------------------------
// Krylov solvers: vectors through Ops, scalars on the host
// ---------------------------------------------------------------------------------------------------------
struct KSPInfo { int its; bool converged; };

// [PETSc] KSPGMRES: left-preconditioned, restarted, x0 = 0, convergence on the preconditioned residual norm
// (mult(in, out): out = A in;  prec(r, z): z = M^-1 r)
template <class Ops, class Mult, class Prec>
KSPInfo gmres(Ops *ops, size_t n, Mult mult, const double *b, double *x, Prec prec, double rtol, double abstol, int restart,
              int max_it, std::vector<double *> &V, double *w, double *t) {
    ops->set(n, 0.0, x);
    prec(b, V[0]);
    double beta = ops->norm2(n, V[0]);
    const double ttol = std::max(rtol * beta, abstol);
    int its = 0;
    if (beta != beta) return {0, false};
    std::vector<double> H((size_t)(restart + 1) * restart), gv(restart + 1), cs(restart), sn(restart), y(restart);
    auto h = [&](int i, int k) -> double & { return H[(size_t)i * restart + k]; };
    while (beta > ttol && its < max_it) {
        std::fill(H.begin(), H.end(), 0.0);
        std::fill(gv.begin(), gv.end(), 0.0);
        gv[0] = beta;
        ops->axpby(n, 1.0 / beta, V[0], 0.0, nullptr, V[0]);
        int k = 0;
        while (k < restart && its < max_it) {
            mult(V[k], t);
            prec(t, w);
            if (ops->gmres_cgs() && restart < 64) {            // classical Gram-Schmidt ([PETSc]'s default): h = V^T w in one
                double hc[64];                                 // batch of dots (one read-back), then w -= V h
                ops->mdot(n, k + 1, V.data(), w, hc);
                for (int i = 0; i <= k; i++) {
                    h(i, k) = hc[i];
                    ops->axpy(n, -hc[i], V[i], w);
                }
            } else {
                for (int i = 0; i <= k; i++) {                 // modified Gram-Schmidt
                    h(i, k) = ops->dot(n, w, V[i]);
                    ops->axpy(n, -h(i, k), V[i], w);
                }
            }
            h(k + 1, k) = ops->norm2(n, w);
            if (h(k + 1, k) != 0.0) ops->axpby(n, 1.0 / h(k + 1, k), w, 0.0, nullptr, V[k + 1]);
            for (int i = 0; i < k; i++) {
                const double tmp = cs[i] * h(i, k) + sn[i] * h(i + 1, k);
                h(i + 1, k) = -sn[i] * h(i, k) + cs[i] * h(i + 1, k);
                h(i, k) = tmp;
            }
            const double d = hypot(h(k, k), h(k + 1, k));
            cs[k] = h(k, k) / d;
            sn[k] = h(k + 1, k) / d;
            h(k, k) = d;
            h(k + 1, k) = 0.0;
            gv[k + 1] = -sn[k] * gv[k];
            gv[k] = cs[k] * gv[k];
            beta = fabs(gv[k + 1]);
            its++;
            k++;
            if (beta <= ttol) break;
        }
        for (int i = k - 1; i >= 0; i--) {                     // back substitution  H y = g
            double s = gv[i];
            for (int j = i + 1; j < k; j++) s -= h(i, j) * y[j];
            y[i] = s / h(i, i);
        }
        for (int i = 0; i < k; i++) ops->axpy(n, y[i], V[i], x);
        if (beta <= ttol) break;
        mult(x, t);                                            // restart: r = M^-1 (b - A x)
        ops->axpby(n, 1.0, b, -1.0, t, t);
        prec(t, V[0]);
        beta = ops->norm2(n, V[0]);
    }
    return {its, beta <= ttol};
}

// [PETSc] KSPCG, preconditioned norm
template <class Ops, class Mult, class Prec>
KSPInfo cg(Ops *ops, size_t n, Mult mult, const double *b, double *x, Prec prec, double rtol, double abstol, int max_it,
           double *r, double *z, double *p, double *w) {
    ops->set(n, 0.0, x);
    ops->copy(n, b, r);
    prec(r, z);
    double beta = ops->dot(n, z, r), dp = ops->norm2(n, z), beta_old = 0.0;
    const double ttol = std::max(rtol * dp, abstol);
    int its = 0;
    while (dp > ttol && its < max_it) {
        if (its == 0) ops->copy(n, z, p);
        else ops->aypx(n, beta / beta_old, z, p);
        mult(p, w);
        const double a = beta / ops->dot(n, p, w);
        ops->axpy(n, a, p, x);
        ops->axpy(n, -a, w, r);
        prec(r, z);
        beta_old = beta;
        beta = ops->dot(n, z, r);
        dp = ops->norm2(n, z);
        its++;
    }
    return {its, dp <= ttol};
}

// ---------------------------------------------------------------------------------------------------------
// [PETSc] SNESLineSearchApply_BT (cubic).  On return w = x - lambda y, g = F(w).  Returns false when the search fails.
// (minlambda uses max|y_i| where PETSc uses max|y_i| / max(|x_i|, 1) <= it: the failure test is marginally laxer.)
// ---------------------------------------------------------------------------------------------------------
template <class Ops, class Fn>
bool linesearch_bt(Ops *ops, size_t n, Fn F, const double *x, const double *f, double fnorm, const double *y, const double *Jy,
                   double *w, double *g, double *gnorm_out, double *lambda_out) {
    const double alpha = 1.0e-4, steptol = 1.0e-12;
    const double yinf = ops->norminf(n, y);
    if (yinf == 0.0) {
        ops->copy(n, x, w);
        ops->copy(n, f, g);
        *gnorm_out = fnorm;
        *lambda_out = 0.0;
        return true;
    }
    const double minlambda = steptol / yinf;
    double initslope = ops->dot(n, f, Jy);
    if (initslope > 0.0) initslope = -initslope;
    if (initslope == 0.0) initslope = -1.0;
    auto trial = [&](double lam) {
        ops->axpby(n, 1.0, x, -lam, y, w);
        F(w, g);
        return ops->norm2(n, g);
    };
    auto clamp = [](double lamtemp, double lam) { return lamtemp > 0.5 * lam ? 0.5 * lam : (lamtemp <= 0.1 * lam ? 0.1 * lam : lamtemp); };
    double lam = 1.0, gnorm = trial(lam);
    if (0.5 * gnorm * gnorm <= 0.5 * fnorm * fnorm + lam * alpha * initslope) { *gnorm_out = gnorm; *lambda_out = lam; return true; }
    double lamprev = lam, gnormprev = gnorm;
    lam = clamp(-initslope / (gnorm * gnorm - fnorm * fnorm - 2.0 * initslope), lam);
    gnorm = trial(lam);
    if (0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope) { *gnorm_out = gnorm; *lambda_out = lam; return true; }
    for (int it = 0; it < 40; it++) {
        if (lam <= minlambda) return false;
        const double t1 = 0.5 * (gnorm * gnorm - fnorm * fnorm) - lam * initslope;
        const double t2 = 0.5 * (gnormprev * gnormprev - fnorm * fnorm) - lamprev * initslope;
        const double a = (t1 / (lam * lam) - t2 / (lamprev * lamprev)) / (lam - lamprev);
        const double b = (-lamprev * t1 / (lam * lam) + lam * t2 / (lamprev * lamprev)) / (lam - lamprev);
        const double d = std::max(b * b - 3.0 * a * initslope, 0.0);
        const double lamtemp = (a == 0.0) ? -initslope / (2.0 * b) : (-b + sqrt(d)) / (3.0 * a);
        lamprev = lam;
        gnormprev = gnorm;
        lam = clamp(lamtemp, lam);
        gnorm = trial(lam);
        if (0.5 * gnorm * gnorm < 0.5 * fnorm * fnorm + lam * alpha * initslope) { *gnorm_out = gnorm; *lambda_out = lam; return true; }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------
// [PETSc] SNESSolve_NEWTONLS on lev[0] (iterate in lev[0].u, updated in place)
// ---------------------------------------------------------------------------------------------------------
template <class Ops>
int newton(Ops *ops, std::vector<Level<Ops>> &lev, const MinimalOpts &opt, const Printer &pr, int indent, StageResult *res) {
    Level<Ops> &L = lev[0];
    const size_t n = L.n;
    const double q = opt.q;
    std::string pad((size_t)(2 * indent), ' ');
    auto F = [&](const double *u, double *f) { ops->minimal_function(L.mx, L.my, q, u, L.g, f); };
    double *y = ops->alloc(n), *Jy = ops->alloc(n), *w = ops->alloc(n), *gnew = ops->alloc(n), *t = ops->alloc(n);
    double *mfw = opt.mf_operator ? ops->alloc(n) : nullptr;
    double *kr = ops->alloc(n), *kz = nullptr, *kp = nullptr;
    std::vector<double *> V;
    if (opt.ksp_type == KSP_GMRES) {
        for (int i = 0; i <= opt.gmres_restart; i++) V.push_back(ops->alloc(n));
    } else {
        kz = ops->alloc(n);
        kp = ops->alloc(n);
    }
    AssembledMG<Ops> mg{ops, &lev, opt.smooth_its};
    const bool use_mg = opt.pc_type == PC_MG && lev.size() > 1;
    double *dense = nullptr;
    int rc = 0;
    memset(res, 0, sizeof *res);
    res->mx = L.mx;
    res->my = L.my;
    F(L.u, L.F);
    double fnorm = ops->norm2(n, L.F);
    res->fnorm[0] = fnorm;
    auto monitor = [&](int it, double v) {
        if (opt.snes_monitor == 2) pr.out("%s%3d SNES Function norm %s", pad.c_str(), it, g6(v).c_str());
        else if (opt.snes_monitor == 1) pr.out("%s%3d SNES Function norm %.12e", pad.c_str(), it, v);
    };
    ops->user_monitor(L.mx, L.my, 0, fnorm, indent, L.u);      // [PETSc] SNESMonitorSet monitors run before -snes_monitor's
    monitor(0, fnorm);
    int reason = 0;
    if (fnorm < opt.snes_atol) reason = SNES_CONVERGED_FNORM_ABS;
    const double ttol = opt.snes_rtol * fnorm;
    int it = 0;
    while (!reason && !rc) {
        if (it >= opt.snes_max_it || it >= MAX_NEWTON) { reason = SNES_DIVERGED_MAX_IT; break; }
        Precond<Ops> M{ops};
        M.n = (int)n;
        if (use_mg) {
            rc = mg.setup(q);
            if (rc) break;
            M.mg = &mg;
        } else if (!(opt.mf_operator && opt.pc_type == PC_NONE)) {        // (matrix-free and unpreconditioned: no matrix at all)
            L.assemble(q, true);
            if (opt.pc_type == PC_MG) {                        // a single level: the "multigrid" is the direct solve
                if (n > 4225) { rc = 61; break; }
                if (!dense) dense = ops->alloc(n * n);
                rc = stencil9_inverse_ops(ops, L.vals, L.mx, L.my, dense);
                if (rc) break;
                M.dense = dense;
            }
        }
        KSPInfo k;
        // -snes_mf_operator: [PETSc] MatMFFD, "wp": J v = (F(u + h v) - F(u)) / h, h = sqrt(eps) sqrt(1 + ||u||) / ||v||
        // (c/ch7/output/minimal.test3: its first stage, one unknown, is reproduced digit for digit with this h)
        const double unorm = opt.mf_operator ? ops->norm2(n, L.u) : 0.0;
        auto mult = [&](const double *in, double *out) {
            if (!opt.mf_operator) { L.mult(in, out); return; }
            const double vn = ops->norm2(n, in);
            if (vn == 0.0) { ops->set(n, 0.0, out); return; }
            const double h = 1.4901161193847656e-08 * sqrt(1.0 + unorm) / vn;
            ops->axpby(n, 1.0, L.u, h, in, mfw);
            F(mfw, out);
            ops->axpby(n, 1.0 / h, out, -1.0 / h, L.F, out);
        };
        auto prec = [&](const double *r, double *z) { M.apply(r, z); };
        if (opt.ksp_type == KSP_GMRES) k = gmres(ops, n, mult, L.F, y, prec, opt.ksp_rtol, 1.0e-50, opt.gmres_restart, opt.ksp_max_it, V, w, t);
        else k = cg(ops, n, mult, L.F, y, prec, opt.ksp_rtol, 1.0e-50, opt.ksp_max_it, kr, kz, kp, w);
        res->ksp_its[it] = k.its;
        if (opt.ksp_converged_reason)
            pr.out("%s    Linear solve %s due to %s iterations %d", pad.c_str(), k.converged ? "converged" : "did not converge",
                   k.converged ? "CONVERGED_RTOL" : "DIVERGED_ITS", k.its);
        mult(y, Jy);
        double gnorm = 0.0, lam = 0.0;
        if (!linesearch_bt(ops, n, F, L.u, L.F, fnorm, y, Jy, w, gnew, &gnorm, &lam)) { reason = SNES_DIVERGED_LINE_SEARCH; break; }
        res->lambda[it] = lam;
        ops->axpby(n, 1.0, w, -1.0, L.u, y);                   // the step actually taken (y is free now)
        const double snorm = ops->norm2(n, y), xnorm = ops->norm2(n, w);
        ops->copy(n, w, L.u);
        ops->copy(n, gnew, L.F);
        fnorm = gnorm;
        it++;
        res->fnorm[it] = fnorm;
        ops->user_monitor(L.mx, L.my, it, fnorm, indent, L.u);
        monitor(it, fnorm);
        if (fnorm != fnorm) reason = SNES_DIVERGED_FNORM_NAN;
        else if (fnorm < opt.snes_atol) reason = SNES_CONVERGED_FNORM_ABS;
        else if (fnorm <= ttol) reason = SNES_CONVERGED_FNORM_RELATIVE;
        else if (snorm < opt.snes_stol * xnorm) reason = SNES_CONVERGED_SNORM_RELATIVE;
        if (ops->error()) rc = ops->error();
    }
    res->its = it;
    res->nksp = it;
    res->reason = reason;
    // a residual that was substituted by the library's kernel is checked against the caller's callback once more, at the
    // iterate this grid converged to (ModelOps::verify_converged; a no-op for every other set of operations)
    if (!rc && !ops->verify_converged(L.mx, L.my, L.u, L.F)) rc = 68;
    if (!rc && opt.snes_converged_reason)
        pr.out("%s  Nonlinear solve %s due to %s iterations %d", pad.c_str(), reason > 0 ? "converged" : "did not converge",
               snes_reason_name(reason), it);
    mg.destroy();
    if (dense) ops->release(dense);
    for (double *p : {y, Jy, w, gnew, t, kr}) ops->release(p);
    if (mfw) ops->release(mfw);
    if (kz) ops->release(kz);
    if (kp) ops->release(kp);
    for (double *p : V) ops->release(p);
    return rc;
}

// the grids of one grid-sequence stage, finest first: [PETSc] PCMG coarsens the DMDA down to the -da_grid_x/_y base grid
// (or -pc_mg_levels); -pc_type none has the one grid
inline std::vector<std::pair<int, int>> level_shapes(const MinimalOpts &opt, int fx, int fy) {
    std::vector<std::pair<int, int>> s{{fx, fy}};
    if (opt.pc_type != PC_MG) return s;
    while (opt.mg_levels ? (int)s.size() < opt.mg_levels : true) {
        const int cx = s.back().first, cy = s.back().second;
        if (cx <= 3 || cy <= 3 || (cx - 1) % 2 || (cy - 1) % 2) break;
        if (!opt.mg_levels && cx == opt.grid_x && cy == opt.grid_y) break;
        s.push_back({(cx - 1) / 2 + 1, (cy - 1) / 2 + 1});
    }
    return s;
}
// every grid a solve will touch (all stages of the grid sequence, all levels)
inline std::vector<std::pair<int, int>> all_shapes(const MinimalOpts &opt) {
    std::vector<std::pair<int, int>> out;
    int mx = opt.grid_x, my = opt.grid_y;
    for (int r = 0; r < opt.refine; r++) { mx = 2 * mx - 1; my = 2 * my - 1; }
    for (int stage = 0; stage <= opt.grid_sequence; stage++) {
        if (stage > 0) { mx = 2 * mx - 1; my = 2 * my - 1; }
        for (auto &p : level_shapes(opt, mx, my))
            if (std::find(out.begin(), out.end(), p) == out.end()) out.push_back(p);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------
// minimal.c:main from DMDACreate2d to the error report (c/ch7/minimal.c:128-181)
// u_out: the final iterate (mx*my doubles in Ops memory, nullptr = not wanted; mx, my are in the result)
// ---------------------------------------------------------------------------------------------------------
// u0_host (optional): the caller's initial iterate on the first grid (host, mx*my doubles) -- then the problem is the
// caller's (its residual comes through Ops as a callback), nothing is known about an exact solution, and the final
// report is the caller's business (report = false).
template <class Ops>
int minimal_solve(Ops *ops, const MinimalOpts &opt, const Printer &pr, double **u_out, MinimalResult *R,
                  const double *u0_host = nullptr, bool report = true) {
    memset(R, 0, sizeof *R);
    R->errinf = -1.0;
    if (opt.grid_sequence + 1 > MAX_STAGES) return 60;
    int mx = opt.grid_x, my = opt.grid_y;
    for (int r = 0; r < opt.refine; r++) { mx = 2 * mx - 1; my = 2 * my - 1; }
    double *u_prev = nullptr;
    int rc = 0;
    std::vector<Level<Ops>> lev;
    for (int stage = 0; stage <= opt.grid_sequence && !rc; stage++) {
        if (stage > 0) { mx = 2 * mx - 1; my = 2 * my - 1; }
        std::vector<std::pair<int, int>> shapes = level_shapes(opt, mx, my);
        std::vector<Level<Ops>> next(shapes.size());
        for (size_t l = 0; l < shapes.size(); l++) next[l].create(ops, shapes[l].first, shapes[l].second, opt);
        Level<Ops> &L = next[0];
        if (stage == 0) {
            if (u0_host) ops->from_host(u0_host, L.u, L.n);                    // the caller's u_initial (minimal.c:150-158)
            else if (opt.exact_init) ops->copy(L.n, L.g, L.u);                 // FormExactFromG (minimal.c:191-208)
            else ops->initial_state2d(L.mx, L.my, L.g, L.u);                   // InitialState(ZEROS, gonboundary) (:157)
        } else {
            ops->set(L.n, 0.0, L.u);
            ops->prolong_add2d(L.mx, L.my, u_prev, L.u);                       // [PETSc] DMRefine + MatInterpolate
        }
        for (Level<Ops> &o : lev) o.destroy();
        lev.swap(next);
        rc = newton(ops, lev, opt, pr, opt.grid_sequence - stage, &R->stage[stage]);
        R->nstages = stage + 1;
        u_prev = lev[0].u;
        if (!rc && ops->error()) rc = ops->error();
    }
    if (!rc) {
        Level<Ops> &L = lev[0];
        R->mx = L.mx;
        R->my = L.my;
        const char *pname = opt.problem == 0 ? "tent" : "catenoid";
        if (report && opt.problem == 1 && opt.q == -0.5) {
            ops->axpby(L.n, 1.0, L.u, -1.0, L.g, L.t);
            R->errinf = ops->norminf(L.n, L.t);
            pr.out("done on %d x %d grid and problem %s:  error |u-uexact|_inf = %.5e", L.mx, L.my, pname, R->errinf);   // :177
        } else if (report) {
            pr.out("done on %d x %d grid and problem %s ...", L.mx, L.my, pname);                                       // :180
        }
        if (u_out) {
            *u_out = ops->alloc(L.n);
            ops->copy(L.n, L.u, *u_out);
        }
    }
    for (Level<Ops> &o : lev) o.destroy();
    R->error = rc;
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Recognising the caller's residual.  p4b_snes2d_solve takes the residual as a HOST callback (the FormFunctionLocal
// contract).  When that callback is c/ch7/minimal.c:210-282 itself -- the unchanged minimal.c under the PETSc-shaped shim --
// the library holds the same function as a kernel, and nine host evaluations per level Jacobian are nine too many.  So,
// before the solve, the callback is probed on every grid the solve will touch:
//   g      F(0) on the boundary rows is -g (minimal.c:227): the Dirichlet data of that grid, whatever g_bdry the caller uses
//   q      the one number of the model: the value for which the kernel reproduces the callback at a generic iterate
//          (-1/2 is tried first; otherwise a secant iteration on a fixed functional of F_kernel(q) - F_callback)
//   check  at a second generic iterate, on EVERY grid, kernel and callback must agree to rounding
// If all of that holds the solve runs with the residual on the device (ModelOps below: minimal_sample hands out the probed
// g); if anything does not, the caller's residual is not that model and the solve evaluates the callback on the host
// every time (CallbackOps).  Monitors are host callbacks either way.
// ---------------------------------------------------------------------------------------------------------
struct ProbedModel {
    bool ok = false;
    double q = 0.0;
    int callbacks = 0;
    std::vector<std::pair<std::pair<int, int>, std::vector<double>>> g;     // per grid: Dirichlet data (interior entries 0)
    const std::vector<double> *find(int mx, int my) const {
        for (auto &e : g)
            if (e.first.first == mx && e.first.second == my) return &e.second;
        return nullptr;
    }
};

// resid(mx, my, u_host, F_host) -> 0 on success: the caller's callback
template <class Ops, class Resid>
bool probe_minimal_model(Ops *ops, Resid resid, const MinimalOpts &opt, ProbedModel *M) {
    M->ok = false;
    unsigned long long lcg = 0x2545F4914F6CDD1DULL;
    auto rnd = [&]() { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(lcg >> 11) / 9007199254740992.0; };
    bool first = true;
    for (auto &shape : all_shapes(opt)) {
        const int mx = shape.first, my = shape.second;
        const size_t n = (size_t)mx * my;
        std::vector<double> u(n, 0.0), Fu(n), Fd(n), g(n, 0.0);
        auto bd = [&](size_t k) { const int j = (int)(k / mx), i = (int)(k - (size_t)j * mx); return i == 0 || j == 0 || i == mx - 1 || j == my - 1; };
        if (resid(mx, my, u.data(), Fu.data())) return false;                       // F(0): boundary rows are 0 - g
        M->callbacks++;
        for (size_t k = 0; k < n; k++) {
            if (Fu[k] != Fu[k]) return false;
            if (bd(k)) g[k] = -Fu[k];
        }
        double *dg = ops->alloc(n), *du = ops->alloc(n), *dF = ops->alloc(n);
        ops->from_host(g.data(), dg, n);
        auto kernel = [&](double q, std::vector<double> *out) {
            ops->minimal_function(mx, my, q, du, dg, dF);
            ops->to_host(dF, out->data(), n);
        };
        auto generic = [&]() {                                                     // g on the boundary, O(1) slopes inside
            for (size_t k = 0; k < n; k++) u[k] = bd(k) ? g[k] + 0.1 * (rnd() - 0.5) : 0.5 * rnd();
            ops->from_host(u.data(), du, n);
        };
        auto maxdiff = [&](const std::vector<double> &a, const std::vector<double> &b, double *scale) {
            double d = 0.0, sc = 1.0;
            for (size_t k = 0; k < n; k++) {
                const double e = fabs(a[k] - b[k]);
                if (!(e <= d)) d = e;
                sc = std::max(sc, fabs(b[k]));
            }
            *scale = sc;
            return d;
        };
        bool good = true;
        double sc = 1.0;
        if (first) {                                                                // identify q on the first grid
            generic();
            good = !resid(mx, my, u.data(), Fu.data());
            M->callbacks++;
            if (good) {
                kernel(-0.5, &Fd);
                if (maxdiff(Fd, Fu, &sc) <= 1.0e-11 * sc) M->q = -0.5;
                else {
                    std::vector<double> F0(n), F1(n), w(n);
                    double q0 = -0.5, q1 = 0.0;
                    F0 = Fd;
                    kernel(q1, &F1);
                    for (size_t k = 0; k < n; k++) w[k] = F1[k] - F0[k];
                    auto phi = [&](const std::vector<double> &F) { double s2 = 0.0; for (size_t k = 0; k < n; k++) s2 += w[k] * (F[k] - Fu[k]); return s2; };
                    double p0 = phi(F0), p1 = phi(F1);
                    for (int it = 0; it < 40 && p1 != p0 && p1 != 0.0; it++) {
                        const double q2 = q1 - p1 * (q1 - q0) / (p1 - p0);
                        if (!(fabs(q2) < 50.0)) { good = false; break; }
                        q0 = q1; p0 = p1; q1 = q2;
                        kernel(q1, &F1);
                        p1 = phi(F1);
                        if (fabs(q1 - q0) <= 1.0e-15 * std::max(1.0, fabs(q1))) break;
                    }
                    char txt[64];                                                   // an option typed as "-0.3" is strtod("-0.3")
                    snprintf(txt, sizeof txt, "%.12g", q1);
                    const double snapped = atof(txt);
                    kernel(snapped, &Fd);
                    if (good && maxdiff(Fd, Fu, &sc) <= 1.0e-11 * sc) M->q = snapped;
                    else {
                        kernel(q1, &Fd);
                        if (good && maxdiff(Fd, Fu, &sc) <= 1.0e-11 * sc) M->q = q1;
                        else good = false;
                    }
                }
            }
            first = false;
        }
        if (good) {                                                                 // verify on this grid
            generic();
            good = !resid(mx, my, u.data(), Fu.data());
            M->callbacks++;
            if (good) {
                kernel(M->q, &Fd);
                good = maxdiff(Fd, Fu, &sc) <= 1.0e-11 * sc;
            }
        }
        ops->release(dg);
        ops->release(du);
        ops->release(dF);
        if (!good || ops->error()) return false;
        M->g.push_back({shape, std::move(g)});
    }
    M->ok = true;
    return true;
}

// Base (DeviceOps in the library, HostOps in the CPU harness) with the probed Dirichlet data behind minimal_sample and
// the caller's monitor behind user_monitor
template <class Base>
struct ModelOps : Base {
    const ProbedModel *model = nullptr;
    std::function<int(int, int, int, double, int, const double *)> monitor;
    std::vector<double> hu;
    explicit ModelOps(const Base &b) : Base(b) {}
    void minimal_sample(int mx, int my, int, double, double, double *g) {
        const std::vector<double> *h = model->find(mx, my);
        if (!h) { if (!this->err) this->err = 67; return; }
        this->from_host(h->data(), g, h->size());
    }
    void user_monitor(int mx, int my, int its, double fnorm, int tablevel, const double *u) {
        if (!monitor || this->err) return;
        const size_t n = (size_t)mx * my;
        hu.resize(n);
        this->to_host(u, hu.data(), n);
        if (!this->err && monitor(mx, my, its, fnorm, tablevel, hu.data())) this->err = 66;
    }
    // The probes are a finite sample of the callback.  A callback that equals the model there but not where the solve
    // ends up (a term that acts for u > 1, an obstacle, ...) is caught here: at the converged iterate of every grid the
    // caller's residual is evaluated once more and must equal the kernel's to rounding; if not, the caller falls back
    // to the host-callback route (nk_device.cu).  One callback evaluation per grid.
    std::function<int(int, int, const double *, double *)> callback;
    double verify_worst = 0.0;
    bool verify_converged(int mx, int my, const double *u, const double *F) {
        if (!callback || this->err) return true;
        const size_t n = (size_t)mx * my;
        std::vector<double> hF(n), hFk(n);
        hu.resize(n);
        this->to_host(u, hu.data(), n);
        this->to_host(F, hFk.data(), n);
        if (this->err || callback(mx, my, hu.data(), hF.data())) return false;
        double d = 0.0, sc = 1.0;
        for (size_t k = 0; k < n; k++) {
            const double e = fabs(hF[k] - hFk[k]);
            if (!(e <= d)) d = e;
            sc = std::max(sc, fabs(hu[k]));
        }
        verify_worst = std::max(verify_worst, d / sc);
        return d <= 1.0e-10 * sc;
    }
};

}  // namespace nk
}  // namespace p4b
