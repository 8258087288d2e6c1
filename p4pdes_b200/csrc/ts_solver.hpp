// ts_solver.hpp -- host logic of the time stepping around c/ch5/pattern.c, in C++ over an abstract set of vector/kernel
// operations (`Ops`); companion of nk_solver.hpp (same pattern: DeviceOps in the library, plain C++ loops in
// oracle/native for CPU checks).  Statement for statement the logic of p4pdes_b200/pattern.py, which reproduces the
// goldens c/ch5/output/pattern.test1-4 verbatim:
//   -ts_type arkimex   [PETSc] TSARKIMEX3 = ARK3(2)4L[2]SA (Kennedy & Carpenter 2003) + TSAdaptBasic + MATCHSTEP
//   -ts_type beuler|cn [PETSc] TSTHETA (theta = 1 | 1/2 endpoint), fixed steps, Newton + bt on every step
//   -ts_type bdf       [PETSc] TSBDF, order 2: backward-Euler half-step restart, variable-step weights from Lagrange-basis
//                      derivatives, extrapolated initial guess, LTE from the next-higher difference, TSAdaptBasic.
//                      c/ch5/output/pattern.test5 pins the restart step (Newton counts 3, 2; next step 1.10972); later
//                      steps have no golden (checked for second-order accuracy against the oracle): parity unpinned there
//   -ts_type rk        [PETSc] TSRK default "3bs": Bogacki-Shampine 3(2), explicit, TSAdaptBasic on the embedded 2nd-order
//                      solution with order 3 (c/ch5/output/heat.test2 pins the step sequence); only through the
//                      callback route (p4b_ts_solve_callbacks: c/ch5/heat.c), and only for Ydot = G(t, Y)
//   stage solves       GMRES(30) preconditioned by a V cycle on the rediscretised, matrix-free stage operator
//                      J = shift*I - C L9 - G'(Y) (FormIJacobianLocal / FormRHSJacobianLocal, pattern.c:202-318)
#pragma once
#include <functional>

#include "nk_solver.hpp"

namespace p4b {
namespace nk {

enum { TS_ARKIMEX = 0, TS_BEULER = 1, TS_CN = 2, TS_BDF = 3, TS_RK = 4 };

struct PatternOpts {
    double L, Du, Dv, phi, kappa;          // -ptn_L -ptn_Du -ptn_Dv -ptn_phi -ptn_kappa (pattern.c:47-52)
    int no_rhsjacobian, call_back_report;  // -ptn_no_rhsjacobian -ptn_call_back_report
    int grid_x, grid_y, refine;
    int ts_type;                           // TS_ARKIMEX | TS_BEULER | TS_CN | TS_BDF
    double ts_dt, ts_max_time;
    int ts_max_steps;
    double ts_rtol, ts_atol;
    int ts_monitor;
    int pc_type;                           // PC_NONE | PC_MG
    int smooth_its;
    double mg_rscale;                      // -p4b_mg_rscale (1 = [PETSc] R = P^T; 0.25 = averaging restriction)
    double snes_rtol, snes_stol, snes_atol;
    int snes_max_it;
    double ksp_rtol;
    int ksp_max_it, gmres_restart;
    int snes_converged_reason, ksp_converged_reason;
};

inline void default_opts(PatternOpts *o) {
    memset(o, 0, sizeof *o);
    o->L = 2.5; o->Du = 8.0e-5; o->Dv = 4.0e-5; o->phi = 0.024; o->kappa = 0.06;
    o->grid_x = o->grid_y = 3;
    o->ts_type = TS_ARKIMEX; o->ts_dt = 5.0; o->ts_max_time = 200.0; o->ts_max_steps = 5000;      // pattern.c:115-118
    o->ts_rtol = o->ts_atol = 1.0e-4;
    o->pc_type = PC_MG; o->smooth_its = 2; o->mg_rscale = 1.0;
    o->snes_rtol = 1.0e-8; o->snes_stol = 1.0e-8; o->snes_atol = 1.0e-50; o->snes_max_it = 50;
    o->ksp_rtol = 1.0e-5; o->ksp_max_it = 10000; o->gmres_restart = 30;
}

constexpr int MAX_TS_STEPS_KEPT = 512;
struct PatternResult {
    int m, nsteps, rejected;
    long long ksp_its_total, newton_its_total;
    double t_final, dt_last;
    double step_t[MAX_TS_STEPS_KEPT], step_dt[MAX_TS_STEPS_KEPT];      // the first steps: time after, step taken
    int step_newton[MAX_TS_STEPS_KEPT];
    int error;
};

// PETSc's %g: an integral value prints with a trailing '.' ("5.", "200.")
inline std::string fmt_g(double v) {
    char b[64];
    snprintf(b, sizeof b, "%g", v);
    std::string s(b);
    const bool integral = s.find_first_not_of("-0123456789") == std::string::npos;
    return integral ? s + "." : s;
}

// dense inverse by Gauss-Jordan with partial pivoting (base grid of the periodic hierarchy: 2*m*m <= 512 unknowns)
inline int dense_inverse(std::vector<double> &A, int n, std::vector<double> *inv) {
    inv->assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) (*inv)[(size_t)i * n + i] = 1.0;
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int r = k + 1; r < n; r++)
            if (fabs(A[(size_t)r * n + k]) > fabs(A[(size_t)p * n + k])) p = r;
        if (A[(size_t)p * n + k] == 0.0) return 1;
        if (p != k)
            for (int c = 0; c < n; c++) {
                std::swap(A[(size_t)p * n + c], A[(size_t)k * n + c]);
                std::swap((*inv)[(size_t)p * n + c], (*inv)[(size_t)k * n + c]);
            }
        const double d = 1.0 / A[(size_t)k * n + k];
        for (int c = 0; c < n; c++) { A[(size_t)k * n + c] *= d; (*inv)[(size_t)k * n + c] *= d; }
        for (int r = 0; r < n; r++) {
            if (r == k) continue;
            const double f = A[(size_t)r * n + k];
            if (f == 0.0) continue;
            for (int c = 0; c < n; c++) { A[(size_t)r * n + c] -= f * A[(size_t)k * n + c]; (*inv)[(size_t)r * n + c] -= f * (*inv)[(size_t)k * n + c]; }
        }
    }
    return 0;
}

// host copy of the base-grid stage operator (FormIJacobianLocal minus FormRHSJacobianLocal), dense, row-major
inline void dense_stage_jacobian(int m, double shift, const double *Y, const PatternOpts &o, std::vector<double> *A) {
    const double h = o.L / m, C[2] = {o.Du / (6.0 * h * h), o.Dv / (6.0 * h * h)};
    const int n = m * m;
    A->assign((size_t)4 * n * n, 0.0);
    auto at = [&](int r, int c) -> double & { return (*A)[(size_t)r * 2 * n + c]; };
    static const int nb[8][3] = {{0, -1, 4}, {0, 1, 4}, {-1, 0, 4}, {1, 0, 4}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};
    for (int j = 0; j < m; j++)
        for (int i = 0; i < m; i++) {
            const int k = j * m + i;
            for (int c = 0; c < 2; c++) {
                const int r = 2 * k + c;
                at(r, r) += shift + 20.0 * C[c];
                for (auto &e : nb) at(r, 2 * (((j + e[0] + m) % m) * m + (i + e[1] + m) % m) + c) += -(double)e[2] * C[c];
            }
            if (Y) {
                const double u = Y[2 * k], v = Y[2 * k + 1];
                at(2 * k, 2 * k) -= -v * v - o.phi;
                at(2 * k, 2 * k + 1) -= -2.0 * u * v;
                at(2 * k + 1, 2 * k) -= v * v;
                at(2 * k + 1, 2 * k + 1) -= 2.0 * u * v - (o.phi + o.kappa);
            }
        }
}

template <class Ops>
struct PLevel {
    int m = 0;
    size_t n = 0;
    double *Y = nullptr, *x = nullptr, *b = nullptr, *t = nullptr;
    double scale = 0.0;
    std::vector<double> omega;
};

// J = shift*I - C L9 - G'(Y) on every level of the periodic hierarchy (lev[0] finest) and its V cycle
template <class Ops>
struct StageOperator {
    Ops *ops;
    const PatternOpts *opt;
    std::vector<PLevel<Ops>> lev;
    bool no_rhs = false;
    double shift = 0.0;
    double *Ainv = nullptr;
    bool ready = false;

    // n_field != 0: a field of that many doubles in a layout only the callbacks know (one level, no multigrid)
    void create(Ops *o, const PatternOpts *op, int m, bool imex, size_t n_field = 0) {
        ops = o; opt = op;
        no_rhs = op->no_rhsjacobian || imex;        // IMEX: the reaction is explicit, G' never enters the stage matrix
        std::vector<int> sizes{m};
        if (op->pc_type == PC_MG && !n_field)
            while (sizes.back() > op->grid_x && sizes.back() % 2 == 0) sizes.push_back(sizes.back() / 2);
        lev.resize(sizes.size());
        for (size_t l = 0; l < sizes.size(); l++) {
            PLevel<Ops> &L = lev[l];
            L.m = sizes[l];
            L.n = n_field ? n_field : (size_t)2 * L.m * L.m;
            L.Y = ops->alloc(L.n); L.x = ops->alloc(L.n); L.b = ops->alloc(L.n); L.t = ops->alloc(L.n);
        }
    }
    void destroy() {
        for (PLevel<Ops> &L : lev) {
            ops->release(L.Y);
            ops->release(L.x);
            ops->release(L.b);
            ops->release(L.t);
        }
        lev.clear();
        if (Ainv) ops->release(Ainv);
        Ainv = nullptr;
    }
    const double *Yof(const PLevel<Ops> &L) const { return no_rhs ? nullptr : L.Y; }
    void mult(const double *in, double *out) {
        ops->pattern_jac_apply(lev[0].m, *opt, shift, Yof(lev[0]), in, out);
    }
    int setup(double sh) {
        shift = sh;
        for (size_t l = 0; l < lev.size(); l++) {
            PLevel<Ops> &L = lev[l];
            if (l > 0 && !no_rhs) ops->pattern_inject(L.m, L.m, lev[l - 1].Y, L.Y);
            if (l + 1 < lev.size()) {
                const double lam = ops->pattern_jac_gershgorin(L.m, *opt, shift, Yof(L), L.t);
                const double emin = 0.1 * lam, emax = 1.1 * lam;
                L.scale = 2.0 / (emax + emin);
                const double alpha = 1.0 - L.scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
                double cm1 = 1.0, ck = mu;
                L.omega.clear();
                for (int i = 1; i < opt->smooth_its; i++) {
                    const double cp1 = 2.0 * mu * ck - cm1;
                    L.omega.push_back(omegaprod * ck / cp1);
                    cm1 = ck;
                    ck = cp1;
                }
            }
        }
        if (opt->pc_type != PC_MG) { ready = true; return 0; }
        PLevel<Ops> &C = lev.back();
        if (C.n > 512) return 61;
        std::vector<double> Yh, A, inv;
        if (!no_rhs) { Yh.resize(C.n); ops->to_host(C.Y, Yh.data(), C.n); }
        dense_stage_jacobian(C.m, shift, no_rhs ? nullptr : Yh.data(), *opt, &A);
        if (dense_inverse(A, (int)C.n, &inv)) return 62;
        if (!Ainv) Ainv = ops->alloc(C.n * C.n);
        ops->from_host(inv.data(), Ainv, C.n * C.n);
        ready = true;
        return 0;
    }
    void smooth(PLevel<Ops> &L, bool zero_guess) {
        const int its = opt->smooth_its;
        if (its <= 0) {
            if (zero_guess) ops->set(L.n, 0.0, L.x);
            return;
        }
        double *pm1 = L.x, *pk = L.t;
        if (zero_guess) ops->set(L.n, 0.0, pm1);
        ops->pattern_jac_lin(L.m, *opt, shift, Yof(L), pm1, L.b, nullptr, 0.0, 1.0, L.scale, 1, pk);
        for (int i = 1; i < its; i++) {
            const double w = L.omega[i - 1];
            ops->pattern_jac_lin(L.m, *opt, shift, Yof(L), pk, L.b, pm1, 1.0 - w, w, w * L.scale, 1, pm1);
            std::swap(pm1, pk);
        }
        if (pk != L.x) std::swap(L.x, L.t);
    }
    void cycle(size_t l, bool zero_guess) {
        PLevel<Ops> &L = lev[l];
        if (l + 1 == lev.size()) {
            ops->dense_matvec((int)L.n, Ainv, L.b, L.x);
            return;
        }
        PLevel<Ops> &C = lev[l + 1];
        smooth(L, zero_guess);
        ops->pattern_jac_lin(L.m, *opt, shift, Yof(L), L.x, L.b, nullptr, 0.0, 0.0, 1.0, 0, L.t);      // b - J x
        ops->pattern_restrict(C.m, C.m, L.t, C.b);
        if (opt->mg_rscale != 1.0) ops->axpby(C.n, opt->mg_rscale, C.b, 0.0, nullptr, C.b);
        cycle(l + 1, true);
        ops->pattern_prolong_add(C.m, C.m, C.x, L.x);
        smooth(L, false);
    }
    void precond(const double *r, double *z) {
        if (opt->pc_type != PC_MG) { ops->copy(lev[0].n, r, z); return; }
        ops->copy(lev[0].n, r, lev[0].b);
        cycle(0, true);
        ops->copy(lev[0].n, lev[0].x, z);
    }
};

// ARK3(2)4L[2]SA
static const double ARK3_G = 1767732205903.0 / 4055673282236.0;
static const double ARK3_AI[4][4] = {{0, 0, 0, 0}, {ARK3_G, ARK3_G, 0, 0},
    {2746238789719.0 / 10658868560708.0, -640167445237.0 / 6845629431997.0, ARK3_G, 0},
    {1471266399579.0 / 7840856788654.0, -4482444167858.0 / 7529755066697.0, 11266239266428.0 / 11593286722821.0, ARK3_G}};
static const double ARK3_AE[4][4] = {{0, 0, 0, 0}, {1767732205903.0 / 2027836641118.0, 0, 0, 0},
    {5535828885825.0 / 10492691773637.0, 788022342437.0 / 10882634858940.0, 0, 0},
    {6485989280629.0 / 16251701735622.0, -4246266847089.0 / 9704473918619.0, 10755448449292.0 / 10357097424841.0, 0}};
static const double ARK3_BH[4] = {2756255671327.0 / 12835298489170.0, -10771552573575.0 / 22201958757719.0,
                                  9247589265047.0 / 10645013368117.0, 2193209047091.0 / 5459859503100.0};

// [PETSc] TSAdaptChoose_Basic: the extra factor 1/2 only from the second consecutive rejection on; order = the order the
// error estimate was taken at (3 for ARK3(2)4L, k + 1 for BDF)
inline bool adapt_basic(double h, double enorm, bool prev_accept, double *hnext, int order = 3) {
    const bool accept = enorm <= 1.0;
    const double s = 0.9 * ((!accept && !prev_accept) ? 0.5 : 1.0);
    double hfac = enorm > 0.0 ? s * pow(enorm, -1.0 / (double)order) : INFINITY;
    hfac = std::min(std::max(hfac, 0.1), 10.0);
    *hnext = h * hfac;
    return accept;
}
// TS_EXACTFINALTIME_MATCHSTEP as TSAdaptChoose applies it; t = time after the accepted step
inline double match_step(double t, double hnext, double tmax) {
    if (t >= tmax) return hnext;
    const double hmax = tmax - t, tend = t + hnext;
    double out = hnext;
    if (tend > tmax) out = hmax;
    if (tend < tmax && hnext * 2.0 > hmax) out = hmax / 2.0;
    if (tend < tmax && hnext * 1.01 > hmax) out = hmax;
    return out;
}

// [PETSc] LagrangeBasisVals / LagrangeBasisDers (bdf.c): values and first derivatives at t of the Lagrange basis over T[0..n)
inline void lagrange_vals(int n, double t, const double *T, double *L) {
    for (int k = 0; k < n; k++) {
        L[k] = 1.0;
        for (int j = 0; j < n; j++)
            if (j != k) L[k] *= (t - T[j]) / (T[k] - T[j]);
    }
}
inline void lagrange_ders(int n, double t, const double *T, double *dL) {
    for (int k = 0; k < n; k++) {
        dL[k] = 0.0;
        for (int j = 0; j < n; j++) {
            if (j == k) continue;
            double p = 1.0 / (T[k] - T[j]);
            for (int l = 0; l < n; l++)
                if (l != k && l != j) p *= (t - T[l]) / (T[k] - T[l]);
            dL[k] += p;
        }
    }
}

// Y0 (optional, Ops memory, 2*m*m doubles): the caller's initial state -- then the run is the caller's (pattern.c under the
// PETSc-shaped shim prints its own banner and call-back report): only the solver's lines are printed.
// n_field != 0 (with Y0): the state is a field of n_field doubles whose layout only the callbacks behind Ops know (the
// method-of-lines system of any DMDA driver, e.g. c/ch5/heat.c: one component, Neumann in x, periodic in y); m is 0 then.
template <class Ops>
int pattern_solve(Ops *ops, const PatternOpts &opt, const Printer &pr, double **Y_out, PatternResult *R,
                  const double *Y0 = nullptr, size_t n_field = 0) {
    memset(R, 0, sizeof *R);
    const int mx = opt.grid_x << opt.refine, my = opt.grid_y << opt.refine;      // periodic: -da_refine doubles
    if (!n_field && mx != my) return 1;                                          // pattern.c:89
    if (n_field && (!Y0 || opt.pc_type == PC_MG)) return 62;
    const int m = n_field ? 0 : mx;
    R->m = m;
    if (!Y0) pr.out("running on %d x %d grid with square cells of side h = %.6f ...", m, m, opt.L / m);      // :94-96
    StageOperator<Ops> A;
    A.create(ops, &opt, m, opt.ts_type == TS_ARKIMEX, n_field);
    const size_t n = A.lev[0].n;
    double *Y = A.lev[0].Y;
    std::vector<double *> V;
    for (int i = 0; i <= opt.gmres_restart; i++) V.push_back(ops->alloc(n));
    double *w = ops->alloc(n), *t1 = ops->alloc(n), *Rv = ops->alloc(n), *d = ops->alloc(n), *Z = ops->alloc(n);
    std::vector<double *> extra;
    auto take = [&]() { double *p = ops->alloc(n); extra.push_back(p); return p; };
    if (Y0) ops->copy(n, Y0, Y);
    else ops->pattern_initial_state(m, m, opt.L, Y);                            // :146-179
    auto mult = [&](const double *in, double *out) { A.mult(in, out); };
    auto prec = [&](const double *r, double *z) { A.precond(r, z); };
    const double tmax = opt.ts_max_time;
    double t = 0.0, h = std::min(opt.ts_dt, opt.ts_max_time);   // [PETSc] TSSolve: MATCHSTEP clips the first step to the final time
    int k = 0, rc = 0;
    auto record = [&](double dt_taken, int newton) {
        if (k < MAX_TS_STEPS_KEPT) { R->step_t[k] = t; R->step_dt[k] = dt_taken; R->step_newton[k] = newton; }
    };
    if (opt.ts_type == TS_ARKIMEX) {
        double *Ynew = take(), *Yemb = take(), *zero = take();
        double *Ys[4], *FI[4], *FE[4];
        for (int i = 0; i < 4; i++) { Ys[i] = take(); FI[i] = take(); FE[i] = take(); }
        ops->set(n, 0.0, zero);
        while (t < tmax - 1e-12 * std::max(1.0, fabs(tmax)) && k < opt.ts_max_steps && !rc) {
            ops->set_step_size(h);                           // (what TSGetTimeStep answers inside a monitor)
            ops->ts_step(k, t, Y, n);                        // [PETSc] TSMonitor: step, time, solution
            if (opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(h).c_str(), fmt_g(t).c_str());
            bool prev_accept = true;
            double hnext = h;
            int newton_step = 0;
            while (!rc) {
                for (int i = 0; i < 4 && !rc; i++) {
                    double ci = 0.0;                            // stage time t + c_i h, c_i = sum_j a^I_ij (callbacks may depend on t)
                    for (int j = 0; j <= i; j++) ci += ARK3_AI[i][j];
                    ops->set_time(t + ci * h);
                    ops->copy(n, Y, Z);
                    for (int j = 0; j < i; j++) {
                        if (ARK3_AE[i][j] != 0.0) ops->axpy(n, h * ARK3_AE[i][j], FE[j], Z);
                        if (ARK3_AI[i][j] != 0.0) ops->axpy(n, h * ARK3_AI[i][j], FI[j], Z);
                    }
                    if (ARK3_AI[i][i] == 0.0) {                 // explicit first stage: Y_1 = Z, YdotI = -F(Y_1, 0)
                        ops->copy(n, Z, Ys[i]);
                        ops->pattern_ifunction(m, opt, Ys[i], zero, FI[i]);
                        ops->axpby(n, -1.0, FI[i], 0.0, nullptr, FI[i]);
                    } else {
                        // F(Y_i, shift (Y_i - Z)) = 0, shift = 1/(h a_ii): linear; [PETSc] runs Newton on it (rtol 1e-8)
                        const double shift = 1.0 / (h * ARK3_AI[i][i]);
                        if (!A.ready || A.shift != shift) { rc = A.setup(shift); if (rc) break; }
                        ops->copy(n, Ys[i - 1], Ys[i]);
                        auto resid = [&](const double *W, double *f) {
                            ops->axpby(n, shift, W, -shift, Z, d);
                            ops->pattern_ifunction(m, opt, W, d, f);
                        };
                        resid(Ys[i], Rv);
                        const double r0 = ops->norm2(n, Rv);
                        double rn = r0;
                        int its = 0;
                        while (rn > opt.snes_rtol * r0 && rn > opt.snes_atol && its < opt.snes_max_it) {
                            ops->set_linearisation(Ys[i]);      // (only a callback-defined operator needs it: F is linear here)
                            KSPInfo ki = gmres(ops, n, mult, Rv, d, prec, opt.ksp_rtol, 1.0e-50, opt.gmres_restart,
                                               opt.ksp_max_it, V, w, t1);
                            R->ksp_its_total += ki.its;
                            ops->axpy(n, -1.0, d, Ys[i]);
                            resid(Ys[i], Rv);
                            rn = ops->norm2(n, Rv);
                            its++;
                        }
                        newton_step += its;
                        if (rn != rn || (rn > opt.snes_rtol * r0 && rn > opt.snes_atol)) { rc = 64; break; }
                        ops->axpby(n, shift, Ys[i], -shift, Z, FI[i]);
                    }
                    ops->pattern_rhsfunction(m, opt, Ys[i], FE[i]);
                }
                if (rc) break;
                ops->copy(n, Y, Ynew);
                ops->copy(n, Y, Yemb);
                for (int j = 0; j < 4; j++) {
                    ops->axpy(n, h * ARK3_AI[3][j], FI[j], Ynew);
                    ops->axpy(n, h * ARK3_AI[3][j], FE[j], Ynew);
                    ops->axpy(n, h * ARK3_BH[j], FI[j], Yemb);
                    ops->axpy(n, h * ARK3_BH[j], FE[j], Yemb);
                }
                const double enorm = sqrt(ops->wrms2(n, Ynew, Yemb, opt.ts_atol, opt.ts_rtol) / (double)n);
                if (adapt_basic(h, enorm, prev_accept, &hnext)) break;
                prev_accept = false;
                R->rejected++;
                h = hnext;
            }
            if (rc) break;
            ops->copy(n, Ynew, Y);
            t += h;
            R->newton_its_total += newton_step;
            record(h, newton_step);
            R->dt_last = h;
            h = match_step(t, hnext, tmax);
            k++;
            if (ops->error()) rc = ops->error();
        }
        ops->set_step_size(h);
        if (!rc) ops->ts_step(k, t, Y, n);
        if (!rc && opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(h).c_str(), fmt_g(t).c_str());
    } else if (opt.ts_type == TS_RK) {
        // [PETSc] TSRK "3bs" (Bogacki & Shampine 1989): c = (0, 1/2, 3/4, 1), third-order weights = the last row,
        // embedded second-order weights (7/24, 1/4, 1/3, 1/8); TSAdaptBasic with the order of the method (3) -- the rule
        // c/ch5/output/heat.test2 pins (dt 0.001, 0.00226419, 0.00336791, ...; order 2 gives 0.00359127 instead)
        static const double RA[4][4] = {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.75, 0, 0}, {2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0, 0}};
        static const double RC[4] = {0.0, 0.5, 0.75, 1.0}, RBE[4] = {7.0 / 24.0, 0.25, 1.0 / 3.0, 0.125};
        double *K[4], *Ynew = take(), *Yemb = take();
        for (int i = 0; i < 4; i++) K[i] = take();
        while (t < tmax - 1e-12 * std::max(1.0, fabs(tmax)) && k < opt.ts_max_steps && !rc) {
            ops->set_step_size(h);
            ops->ts_step(k, t, Y, n);
            if (opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(h).c_str(), fmt_g(t).c_str());
            bool prev_accept = true;
            double hnext = h;
            while (!rc) {
                for (int i = 0; i < 4; i++) {
                    ops->copy(n, Y, Z);
                    for (int j = 0; j < i; j++)
                        if (RA[i][j] != 0.0) ops->axpy(n, h * RA[i][j], K[j], Z);
                    ops->set_time(t + RC[i] * h);
                    ops->pattern_rhsfunction(m, opt, Z, K[i]);
                }
                ops->copy(n, Y, Ynew);
                ops->copy(n, Y, Yemb);
                for (int j = 0; j < 4; j++) {
                    if (RA[3][j] != 0.0) ops->axpy(n, h * RA[3][j], K[j], Ynew);
                    ops->axpy(n, h * RBE[j], K[j], Yemb);
                }
                const double enorm = sqrt(ops->wrms2(n, Ynew, Yemb, opt.ts_atol, opt.ts_rtol) / (double)n);
                if (ops->error()) { rc = ops->error(); break; }
                if (enorm != enorm) { rc = 64; break; }
                if (adapt_basic(h, enorm, prev_accept, &hnext, 3)) break;
                prev_accept = false;
                R->rejected++;
                h = hnext;
            }
            if (rc) break;
            ops->copy(n, Ynew, Y);
            t += h;
            record(h, 0);
            R->dt_last = h;
            h = match_step(t, hnext, tmax);
            k++;
        }
        ops->set_step_size(h);
        if (!rc) ops->ts_step(k, t, Y, n);
        if (!rc && opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(h).c_str(), fmt_g(t).c_str());
    } else {
        double *Yprev = take(), *Ydot = take(), *G = take(), *affine = take(), *y = take(), *Jy = take(), *wv = take(), *gnew = take();
        // [PETSc] SNESSolve_NEWTONLS on F(W) = 0 from the guess X (in place), stage matrix shift*I - C L9 - G'(W)
        auto newton_solve = [&](const std::function<void(const double *, double *)> &F, double shift, double *X, int *its_out) {
            F(X, Rv);
            double fnorm = ops->norm2(n, Rv);
            const double ttol = opt.snes_rtol * fnorm;
            int reason = fnorm < opt.snes_atol ? SNES_CONVERGED_FNORM_ABS : 0, its = 0;
            while (!reason && !rc) {
                if (its >= opt.snes_max_it) { reason = SNES_DIVERGED_MAX_IT; break; }
                if (X != Y) ops->copy(n, X, Y);                  // the stage operator linearises about lev[0].Y
                rc = A.setup(shift);
                if (rc) break;
                ops->set_linearisation(X);
                KSPInfo ki = gmres(ops, n, mult, Rv, y, prec, opt.ksp_rtol, 1.0e-50, opt.gmres_restart, opt.ksp_max_it, V, w, t1);
                R->ksp_its_total += ki.its;
                if (opt.ksp_converged_reason)
                    pr.out("      Linear solve %s due to %s iterations %d", ki.converged ? "converged" : "did not converge",
                           ki.converged ? "CONVERGED_RTOL" : "DIVERGED_ITS", ki.its);
                A.mult(y, Jy);
                double gnorm = 0.0, lam = 0.0;
                if (!linesearch_bt(ops, n, F, X, Rv, fnorm, y, Jy, wv, gnew, &gnorm, &lam)) { reason = SNES_DIVERGED_LINE_SEARCH; break; }
                ops->axpby(n, 1.0, wv, -1.0, X, y);
                const double snorm = ops->norm2(n, y), xnorm = ops->norm2(n, wv);
                ops->copy(n, wv, X);
                ops->copy(n, gnew, Rv);
                fnorm = gnorm;
                its++;
                if (fnorm != fnorm) reason = SNES_DIVERGED_FNORM_NAN;
                else if (fnorm < opt.snes_atol) reason = SNES_CONVERGED_FNORM_ABS;
                else if (fnorm <= ttol) reason = SNES_CONVERGED_FNORM_RELATIVE;
                else if (snorm < opt.snes_stol * xnorm) reason = SNES_CONVERGED_SNORM_RELATIVE;
            }
            if (!rc && opt.snes_converged_reason)
                pr.out("    Nonlinear solve %s due to %s iterations %d", reason > 0 ? "converged" : "did not converge",
                       snes_reason_name(reason), its);
            *its_out = its;
            return reason;
        };
        if (opt.ts_type == TS_BDF) {
            // [PETSc] TSStep_BDF, order 2 (see the header of this file)
            const int order = 2;
            double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, *wk[8];
            for (int i = 0; i < 8; i++) wk[i] = take();
            double *V0 = take(), *lte = take();
            int kord = 0, nh = 0;
            bool restart = true;
            auto advance = [&](double tt, const double *X) {                          // TSBDF_Advance
                double *tail = wk[7];
                for (int i = 7; i >= 2; i--) { tm[i] = tm[i - 1]; wk[i] = wk[i - 1]; }
                nh = std::min(nh + 1, 7);
                tm[1] = tt;
                wk[1] = tail;
                ops->copy(n, X, tail);
            };
            auto stage = [&](double *X, int *its) {                                    // TSBDF_PreSolve + SNESSolve
                const int nn = std::max(kord, 1) + 1;
                double a[8];
                lagrange_ders(nn, tm[0], tm, a);
                ops->set(n, 0.0, V0);
                for (int i = 1; i < nn; i++) ops->axpy(n, a[i], wk[i], V0);
                const double shift = a[0];
                ops->set_time(tm[0]);
                std::function<void(const double *, double *)> F = [&, shift](const double *W, double *f) {
                    ops->axpby(n, shift, W, 1.0, V0, Ydot);                           // Ydot = shift W + V0
                    ops->pattern_ifunction(m, opt, W, Ydot, f);
                    ops->pattern_rhsfunction(m, opt, W, G);
                    ops->axpy(n, -1.0, G, f);
                };
                return newton_solve(F, shift, X, its);
            };
            double dt_next = h;
            while (t < tmax - 1e-12 * std::max(1.0, fabs(tmax)) && k < opt.ts_max_steps && !rc) {
                ops->set_step_size(h);
                ops->ts_step(k, t, Y, n);                        // [PETSc] TSMonitor: step, time, solution
            if (opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(h).c_str(), fmt_g(t).c_str());
                ops->copy(n, Y, Yprev);                              // lev[0].Y is the linearisation point from here on
                if (!restart) { kord = std::min(kord + 1, order); advance(t, Yprev); }
                bool accept = true;
                int newton_step = 0, its = 0;
                double hnext = h;
                while (!rc) {
                    if (restart) {                                                     // TSBDF_Restart
                        kord = 1; nh = 0;
                        advance(t, Yprev);
                        tm[0] = t + h / 2.0;
                        ops->copy(n, wk[1], wk[0]);
                        if (stage(wk[0], &its) <= 0 && !rc) rc = 64;
                        if (rc) break;
                        newton_step += its;
                        kord = std::min(2, order);
                        nh++;
                        ops->copy(n, wk[0], wk[2]);
                        tm[2] = tm[0];
                    }
                    tm[0] = t + h;
                    {                                                                  // TSBDF_Extrapolate
                        const int ne = std::min(kord - (accept ? 0 : 1) + 1, nh);
                        double c[8];
                        lagrange_vals(ne, tm[0], tm + 1, c);
                        ops->set(n, 0.0, wk[0]);
                        for (int i = 0; i < ne; i++) ops->axpy(n, c[i], wk[1 + i], wk[0]);
                    }
                    if (stage(wk[0], &its) <= 0 && !rc) rc = 64;
                    if (rc) break;
                    newton_step += its;
                    const int kl = std::min(kord, nh - 1);                             // TSEvaluateWLTE_BDF / TSBDF_VecLTE
                    double a[8], b[8];
                    lagrange_ders(kl + 1, tm[0], tm, a);
                    a[kl + 1] = 0.0;
                    lagrange_ders(kl + 2, tm[0], tm, b);
                    ops->copy(n, wk[0], lte);
                    for (int i = 0; i < kl + 2; i++) ops->axpy(n, (a[i] - b[i]) / a[0], wk[i], lte);
                    const double enorm = sqrt(ops->wrms2(n, wk[0], lte, opt.ts_atol, opt.ts_rtol) / (double)n);
                    if (adapt_basic(h, enorm, accept, &hnext, kl + 1)) break;
                    accept = false;
                    R->rejected++;
                    h = hnext;
                }
                if (rc) break;
                ops->copy(n, wk[0], Y);
                t += h;
                R->newton_its_total += newton_step;
                record(h, newton_step);
                R->dt_last = h;
                h = match_step(t, hnext, tmax);
                dt_next = h;
                restart = false;
                k++;
                if (ops->error()) rc = ops->error();
            }
            ops->set_step_size(dt_next);
            if (!rc) ops->ts_step(k, t, Y, n);
        if (!rc && opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(dt_next).c_str(), fmt_g(t).c_str());
        } else {
            const double theta = opt.ts_type == TS_CN ? 0.5 : 1.0;
            double dt_last = opt.ts_dt;
            while (t < tmax - 1e-14 * std::max(1.0, fabs(tmax)) && k < opt.ts_max_steps && !rc) {
                const double dt = std::min(opt.ts_dt, tmax - t);     // TS_EXACTFINALTIME_MATCHSTEP (:118)
                dt_last = dt;
                ops->set_step_size(dt);
                ops->ts_step(k, t, Y, n);                        // [PETSc] TSMonitor: step, time, solution
            if (opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(dt).c_str(), fmt_g(t).c_str());
                const double shift = 1.0 / (theta * dt);
                ops->copy(n, Y, Yprev);
                if (theta != 1.0) {
                    ops->set_time(t);
                    ops->set(n, 0.0, Ydot);
                    ops->pattern_ifunction(m, opt, Yprev, Ydot, affine);
                    ops->pattern_rhsfunction(m, opt, Yprev, G);
                    ops->axpy(n, -1.0, G, affine);
                }
                ops->set_time(t + dt);
                // F(W, (W - Yprev)/(theta dt)) - G(W) + (1 - theta)/theta [F(Yprev, 0) - G(Yprev)]
                std::function<void(const double *, double *)> F = [&](const double *W, double *f) {
                    ops->axpby(n, shift, W, -shift, Yprev, Ydot);
                    ops->pattern_ifunction(m, opt, W, Ydot, f);
                    ops->pattern_rhsfunction(m, opt, W, G);
                    ops->axpy(n, -1.0, G, f);
                    if (theta != 1.0) ops->axpy(n, (1.0 - theta) / theta, affine, f);
                };
                int its = 0;
                const int reason = newton_solve(F, shift, Y, &its);
                if (rc) break;
                if (reason <= 0) { rc = 64; break; }
                t += dt;
                R->newton_its_total += its;
                record(dt, its);
                R->dt_last = dt;
                k++;
                if (ops->error()) rc = ops->error();
            }
            ops->set_step_size(dt_last);
            if (!rc) ops->ts_step(k, t, Y, n);
        if (!rc && opt.ts_monitor) pr.out("%d TS dt %s time %s", k, fmt_g(dt_last).c_str(), fmt_g(t).c_str());
        }
    }
    R->nsteps = k;
    R->t_final = t;
    if (!rc && opt.call_back_report && !Y0) {                                   // pattern.c:127-135
        const char *name = opt.ts_type == TS_ARKIMEX ? "arkimex" : (opt.ts_type == TS_CN ? "cn" : (opt.ts_type == TS_BDF ? "bdf" : "beuler"));
        pr.out("CALL-BACK REPORT");
        pr.out("  solver type: %s", name);
        pr.out("  IFunction:   1  | IJacobian:   1");
        pr.out("  RHSFunction: 1  | RHSJacobian: %d", (opt.ts_type == TS_ARKIMEX || opt.no_rhsjacobian) ? 0 : 1);
    }
    if (!rc && Y_out) {
        *Y_out = ops->alloc(n);
        ops->copy(n, Y, *Y_out);
    }
    for (double *p : V) ops->release(p);
    extra.push_back(w);
    extra.push_back(t1);
    extra.push_back(Rv);
    extra.push_back(d);
    extra.push_back(Z);
    for (double *p : extra) ops->release(p);
    A.destroy();
    R->error = rc;
    return rc;
}

}  // namespace nk
}  // namespace p4b
