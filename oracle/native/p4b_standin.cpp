// p4b_standin.cpp -- HOST stand-in for the few C-ABI entry points the PETSc-shaped shim calls on the SNESNEWTONLS / TS paths
// (TEST INFRASTRUCTURE ONLY: never compiled into libp4b200.so, never loaded by the product).
//
// p4pdes_b200/shim/petscshim.c runs the reference's unchanged c/ch7/minimal.c by handing its FormFunctionLocal to
// p4b_snes2d_solve_monitored (include/p4b200.h).  On a machine without a GPU the shim's host logic -- option parsing,
// DMDALocalInfo and a[j][i] views around the callback, SNESMonitorSet monitors with the stage's DM and iterate, DM
// replacement under -snes_grid_sequence, the final report -- can still be exercised end to end if those entry points
// exist.  This file provides them over host memory: "device" pointers are malloc'ed, the Vec operations are loops, and
// the solve is the SAME template (p4pdes_b200/csrc/nk_solver.hpp) instantiated with HostCallbackOps of host_ops.hpp.
// oracle/Makefile links it with the shim source and the reference's minimal.c / pattern.c into
// oracle/_ref/{minimal,pattern}_shim_host, which tests/test_shim_minimal_cpu.py / test_shim_pattern_cpu.py compare with
// the reference's golden outputs (c/ch7/output/minimal.test*, c/ch5/output/pattern.test*).
// For fish.c the recognised finest-level operator is solved by Jacobi-preconditioned CG (no multigrid: iteration counts
// are not the device path's; the shim's work around the solve is what a CPU test can see).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/p4b200.h"
#include "host_ops.hpp"
#include "nk_solver.hpp"
#include "ts_solver.hpp"

using namespace p4b;

static char g_err[512] = "";
static int fail(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

struct p4b_ctx { int unused; };
static int g_route = 0, g_recognise = 1, g_cgs = 0;
static long long g_callbacks = 0;
struct p4b_mg { p4b_grid g; double diag, c[3]; };
struct p4b_sell { int n; std::vector<int> rp, ci; std::vector<double> v; };

extern "C" {

const char *p4b_last_error(void) { return g_err; }
int p4b_snes2d_last_route(void) { return g_route; }
int p4b_tune(const char *key, long value) {
    if (!strcmp(key, "recognise_residual")) { g_recognise = (int)value; return 0; }
    if (!strcmp(key, "gmres_cgs")) { g_cgs = (int)value; return 0; }
    return fail(62, "stand-in: unknown tuning key");
}
long long p4b_launch_count(void) { return 0; }
int p4b_set_ts_step_monitor(p4b_ts_step_fn fn, void *user) { HostOps::ts_fn() = fn; HostOps::ts_user() = user; return 0; }
// the VecSetRandom stream is host code in the product too (mg.cu): the same drand48 recurrence
unsigned long long p4b_rander48_seed(unsigned long seed) { return (((unsigned long long)(seed & 0xffffffffUL)) << 16) | 0x330EULL; }
int p4b_rander48_fill(unsigned long long *state, size_t n, double *out) {
    unsigned long long x = *state & 0xFFFFFFFFFFFFULL;
    for (size_t i = 0; i < n; i++) {
        x = (0x5DEECE66DULL * x + 0xBULL) & 0xFFFFFFFFFFFFULL;
        out[i] = (double)x * (1.0 / 281474976710656.0);
    }
    *state = x;
    return 0;
}
int p4b_ctx_create(int, void *, p4b_ctx **ctx) { *ctx = new p4b_ctx{0}; return 0; }
// -p4b_gpus N needs GPUs: the stand-in has none (the shim's multi-GPU route is covered by tests/test_gpu_multi.py)
int p4b_ctx_create_own_stream(int, p4b_ctx **) { return fail(70, "stand-in: -p4b_gpus needs CUDA devices"); }
int p4b_comm_unique_id(void *) { return fail(70, "stand-in: -p4b_gpus needs CUDA devices"); }
int p4b_comm_init(p4b_ctx *, const void *, int, int) { return fail(70, "stand-in: -p4b_gpus needs CUDA devices"); }
int p4b_mg_local_range(p4b_mg *, int *, int *, size_t *) { return fail(70, "stand-in: -p4b_gpus needs CUDA devices"); }
int p4b_ctx_destroy(p4b_ctx *ctx) { delete ctx; return 0; }
int p4b_malloc(p4b_ctx *, size_t bytes, void **dptr) { *dptr = malloc(bytes ? bytes : 1); return *dptr ? 0 : 55; }
int p4b_free(p4b_ctx *, void *dptr) { free(dptr); return 0; }
int p4b_memcpy_h2d(p4b_ctx *, void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
int p4b_memcpy_d2h(p4b_ctx *, void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); return 0; }

int p4b_vec_dot(p4b_ctx *, size_t n, const double *x, const double *y, double *r) { HostOps o; *r = o.dot(n, x, y); return 0; }
int p4b_vec_norm2(p4b_ctx *, size_t n, const double *x, double *r) { HostOps o; *r = o.norm2(n, x); return 0; }
int p4b_vec_norminf(p4b_ctx *, size_t n, const double *x, double *r) { HostOps o; *r = o.norminf(n, x); return 0; }
int p4b_vec_axpy(p4b_ctx *, size_t n, double a, const double *x, double *y) { HostOps o; o.axpy(n, a, x, y); return 0; }
int p4b_vec_aypx(p4b_ctx *, size_t n, double a, const double *x, double *y) { HostOps o; o.aypx(n, a, x, y); return 0; }

int p4b_minimal_default_opts(p4b_minimal_opts *o) {
    static_assert(sizeof(p4b_minimal_opts) == sizeof(nk::MinimalOpts), "p4b_minimal_opts and nk::MinimalOpts must agree");
    nk::default_opts(reinterpret_cast<nk::MinimalOpts *>(o));
    return 0;
}

int p4b_snes2d_solve_monitored(p4b_ctx *c, const p4b_minimal_opts *opts, p4b_residual2d_fn residual, p4b_monitor2d_fn monitor,
                               void *user, const double *u0_host, p4b_line_fn line, void *line_ctx, double *u_out_host,
                               size_t u_capacity, p4b_minimal_result *result) {
    static_assert(sizeof(p4b_minimal_result) == sizeof(nk::MinimalResult), "p4b_minimal_result and nk::MinimalResult must agree");
    if (!c || !opts || !residual || !u0_host || !result) return fail(62, "p4b_snes2d_solve: null argument");
    const nk::MinimalOpts &o = *reinterpret_cast<const nk::MinimalOpts *>(opts);
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    nk::Printer pr{line, line_ctx};
    double *u = nullptr;
    nk::MinimalResult &R = *reinterpret_cast<nk::MinimalResult *>(result);
    int rc = 0;
    g_route = 0;
    g_callbacks = 0;
    HostOps base;
    base.cgs = g_cgs != 0;
    nk::ProbedModel model;
    auto resid = [&](int mx, int my, const double *uh, double *Fh) { g_callbacks++; return residual(user, mx, my, uh, Fh); };
    if (g_recognise && nk::probe_minimal_model(&base, resid, o, &model)) {            // as nk_device.cu does
        nk::ModelOps<HostOps> ops(base);
        ops.model = &model;
        if (monitor)
            ops.monitor = [&](int mx, int my, int its, double fnorm, int tab, const double *uh) {
                return monitor(user, mx, my, its, fnorm, tab, uh);
            };
        ops.callback = resid;                   // re-verified at the converged iterate of every grid, as nk_device.cu does
        nk::MinimalOpts o2 = o;
        o2.q = model.q;
        g_route = 1;
        rc = nk::minimal_solve(&ops, o2, pr, u_out_host ? &u : nullptr, &R, u0_host, false);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && u_out_host) {
            const size_t n = (size_t)R.mx * R.my;
            if (u_capacity < n) rc = 63;
            else memcpy(u_out_host, u, sizeof(double) * n);
        }
        if (u) ops.release(u);
        u = nullptr;
        if (rc == 68) {
            fprintf(stderr, "[p4b200] SNES: the residual callback matched the library's kernel at the probes but not at a "
                            "converged iterate (deviation %.3e): solving again with the callback evaluated on the host\n",
                    ops.verify_worst);
            g_route = 2;
            rc = 0;
        }
    }
    if (g_route != 1) {
        g_route = 0;
        HostCallbackOps ops;
        ops.cgs = g_cgs != 0;
        ops.fn = residual;
        ops.mon = monitor;
        ops.user = user;
        rc = nk::minimal_solve(&ops, o, pr, u_out_host ? &u : nullptr, &R, u0_host, false);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && u_out_host) {
            const size_t n = (size_t)R.mx * R.my;
            if (u_capacity < n) rc = 63;
            else memcpy(u_out_host, u, sizeof(double) * n);
        }
        if (u) ops.release(u);
        g_callbacks += ops.callbacks;
    }
    if (getenv("P4B_STANDIN_REPORT"))
        fprintf(stderr, "standin: route %d, q %.17g, residual callbacks %lld\n", g_route, model.q, g_callbacks);
    if (rc == 61) return fail(61, "base grid of the multigrid hierarchy is larger than 65 x 65: use a coarser base grid");
    if (rc == 62) return fail(62, "base-grid Jacobian is singular");
    if (rc == 63) return fail(63, "u_out is too small for the final grid");
    if (rc == 65) return fail(65, "the residual callback returned an error");
    if (rc == 66) return fail(66, "the monitor callback returned an error");
    if (rc) return fail(rc, "p4b_snes2d_solve failed");
    return 0;
}

// ---- pattern.c: TSSolve of the shim (identify + verify against these two, then the solve) ----
int p4b_pattern_ifunction(p4b_ctx *, int mx, int my, double L, double Du, double Dv, const double *Y, const double *Ydot,
                          double *F) {
    if (mx != my) return fail(1, "pattern.c requires mx == my");
    HostOps o;
    nk::PatternOpts po;
    nk::default_opts(&po);
    po.L = L; po.Du = Du; po.Dv = Dv;
    o.pattern_ifunction(mx, po, Y, Ydot, F);
    return 0;
}
int p4b_pattern_rhsfunction(p4b_ctx *, int mx, int my, double phi, double kappa, const double *Y, double *G) {
    if (mx != my) return fail(1, "pattern.c requires mx == my");
    HostOps o;
    nk::PatternOpts po;
    nk::default_opts(&po);
    po.phi = phi; po.kappa = kappa;
    o.pattern_rhsfunction(mx, po, Y, G);
    return 0;
}
int p4b_pattern_default_opts(p4b_pattern_opts *o) {
    static_assert(sizeof(p4b_pattern_opts) == sizeof(nk::PatternOpts), "p4b_pattern_opts and nk::PatternOpts must agree");
    nk::default_opts(reinterpret_cast<nk::PatternOpts *>(o));
    return 0;
}
int p4b_pattern_solve_from(p4b_ctx *c, const p4b_pattern_opts *opts, const double *Y0, p4b_line_fn line, void *line_ctx,
                           double *Y_out, size_t Y_capacity, p4b_pattern_result *result) {
    static_assert(sizeof(p4b_pattern_result) == sizeof(nk::PatternResult), "p4b_pattern_result and nk::PatternResult must agree");
    if (!c || !opts || !result) return fail(62, "p4b_pattern_solve: null argument");
    const nk::PatternOpts &o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    if ((o.grid_x << o.refine) != (o.grid_y << o.refine)) return fail(1, "pattern.c requires mx == my");
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_BDF) return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3)");
    HostOps ops;
    ops.cgs = g_cgs != 0;
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr;
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, Y_out ? &Y : nullptr, &R, Y0);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc && Y_out) {
        const size_t n = (size_t)2 * R.m * R.m;
        if (Y_capacity < n) rc = 63;
        else memcpy(Y_out, Y, sizeof(double) * n);
    }
    if (Y) ops.release(Y);
    if (rc == 61) return fail(61, "base grid of the periodic hierarchy has more than 512 unknowns: use a coarser -da_grid_x/_y");
    if (rc == 62) return fail(62, "base-grid stage Jacobian is singular");
    if (rc == 63) return fail(63, "Y_out is too small for the grid");
    if (rc == 64) return fail(64, "TSSolve: a nonlinear (stage) solve did not converge");
    if (rc) return fail(rc, "p4b_pattern_solve failed");
    return 0;
}

int p4b_ts2d_solve(p4b_ctx *c, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction, p4b_rhsfunction2d_fn rhsfunction,
                   void *user, double *Y_inout_host, size_t Y_capacity, p4b_line_fn line, void *line_ctx,
                   p4b_pattern_result *result) {
    if (!c || !opts || !ifunction || !rhsfunction || !Y_inout_host || !result) return fail(62, "p4b_ts2d_solve: null argument");
    const nk::PatternOpts &o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if ((o.grid_x << o.refine) != (o.grid_y << o.refine)) return fail(1, "the device path needs mx == my");
    if (o.pc_type != nk::PC_NONE) return fail(56, "p4b_ts2d_solve: -pc_type none only");
    const int m = o.grid_x << o.refine;
    const size_t n = (size_t)2 * m * m;
    if (Y_capacity < n) return fail(63, "Y is too small for the grid");
    HostCallbackPatternOps ops;
    ops.cgs = g_cgs != 0;
    ops.ifn = ifunction;
    ops.gfn = rhsfunction;
    ops.user = user;
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr;
    std::vector<double> Y0(Y_inout_host, Y_inout_host + n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, &Y, &R, Y0.data());
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) memcpy(Y_inout_host, Y, sizeof(double) * n);
    if (Y) ops.release(Y);
    if (getenv("P4B_STANDIN_REPORT")) fprintf(stderr, "standin: ts2d callbacks %lld\n", ops.callbacks);
    if (rc == 64) return fail(64, "TSSolve: a nonlinear (stage) solve did not converge");
    if (rc == 65) return fail(65, "a callback returned an error");
    if (rc) return fail(rc, "p4b_ts2d_solve failed");
    return 0;
}

int p4b_ts_solve_callbacks(p4b_ctx *c, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction, p4b_rhsfunction2d_fn rhsfunction,
                           void *user, double *Y_inout_host, size_t n, p4b_line_fn line, void *line_ctx,
                           p4b_pattern_result *result) {
    if (!c || !opts || !ifunction || !rhsfunction || !Y_inout_host || !result || !n)
        return fail(62, "p4b_ts_solve_callbacks: null argument");
    nk::PatternOpts o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_RK) return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3), rk (4)");
    if (o.pc_type != nk::PC_NONE && o.ts_type != nk::TS_RK) return fail(56, "p4b_ts_solve_callbacks: -pc_type none only");
    o.pc_type = nk::PC_NONE;
    HostCallbackPatternOps ops;
    ops.cgs = g_cgs != 0;
    ops.ifn = ifunction;
    ops.gfn = rhsfunction;
    ops.user = user;
    ops.nfield = n;
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr;
    std::vector<double> Y0(Y_inout_host, Y_inout_host + n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, &Y, &R, Y0.data(), n);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) memcpy(Y_inout_host, Y, sizeof(double) * n);
    if (Y) ops.release(Y);
    if (getenv("P4B_STANDIN_REPORT")) fprintf(stderr, "standin: ts callbacks %lld\n", ops.callbacks);
    if (rc == 64) return fail(64, "TSSolve: a stage solve did not converge (or an explicit step produced NaN)");
    if (rc == 65) return fail(65, "a callback returned an error");
    if (rc) return fail(rc, "p4b_ts_solve_callbacks failed");
    return 0;
}
double p4b_ts_time_step(void) { return HostOps::ts_h(); }
int p4b_heat_rhs(p4b_ctx *, int mx, int my, double D0, const double *u, double *G) {
    HostHeatOps ops;
    ops.mx = mx; ops.my = my; ops.D0 = D0;
    ops.heat(0, 0.0, u, G);
    return 0;
}
int p4b_heat_jac_apply(p4b_ctx *, int mx, int my, double D0, double shift, const double *X, double *JX) {
    HostHeatOps ops;
    ops.mx = mx; ops.my = my; ops.D0 = D0;
    ops.heat(1, shift, X, JX);
    return 0;
}
int p4b_heat_solve(p4b_ctx *c, const p4b_pattern_opts *opts, int mx, int my, double D0, double *Y_inout_host, p4b_line_fn line,
                   void *line_ctx, p4b_pattern_result *result) {
    if (!c || !opts || !Y_inout_host || !result) return fail(62, "p4b_heat_solve: null argument");
    nk::PatternOpts o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_RK) return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3), rk (4)");
    if (o.pc_type != nk::PC_NONE && o.ts_type != nk::TS_RK) return fail(56, "p4b_heat_solve: -pc_type none only");
    o.pc_type = nk::PC_NONE;
    o.no_rhsjacobian = 0;
    const size_t n = (size_t)mx * my;
    HostHeatOps ops;
    ops.cgs = g_cgs != 0;
    ops.mx = mx; ops.my = my; ops.D0 = D0;
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr;
    std::vector<double> Y0(Y_inout_host, Y_inout_host + n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, &Y, &R, Y0.data(), n);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) memcpy(Y_inout_host, Y, sizeof(double) * n);
    if (Y) ops.release(Y);
    if (rc == 64) return fail(64, "TSSolve: a stage solve did not converge (or an explicit step produced NaN)");
    if (rc) return fail(rc, "p4b_heat_solve failed");
    return 0;
}

// ---- fish.c: the operator the shim's Mat type recognised (finest level) and a Jacobi-preconditioned CG on it.  There is
// NO multigrid here: iteration counts are not the device path's; what the stand-in lets a CPU test see is everything
// the shim does around the solve (callbacks, Mat recognition, Vec bookkeeping, the reference's own report lines). ----
int p4b_mg_default_opts(p4b_mg_opts *o) { memset(o, 0, sizeof *o); o->smooth_its = 2; o->cycle = P4B_CYCLE_V; return 0; }
int p4b_mg_create_stencil(p4b_ctx *, const p4b_grid *g, const p4b_mg_opts *, const double *coef, int nlev, p4b_mg **mg) {
    if (nlev < 1) return fail(62, "need the stencil of the finest level");
    p4b_mg *m = new p4b_mg;
    m->g = *g;
    m->diag = coef[0];
    for (int d = 0; d < 3; d++) m->c[d] = coef[1 + d];
    *mg = m;
    return 0;
}
int p4b_mg_destroy(p4b_mg *mg) { delete mg; return 0; }
int p4b_mg_matmult(p4b_mg *mg, const double *x, double *y) {
    // poissonfunctions.c:117-258: boundary rows are diag-only, interior rows couple to interior neighbours only
    const int M[3] = {mg->g.mx, mg->g.dim >= 2 ? mg->g.my : 1, mg->g.dim >= 3 ? mg->g.mz : 1};
    const long st[3] = {1, M[0], (long)M[0] * M[1]};
    auto bd = [&](const int *p) {
        for (int d = 0; d < mg->g.dim; d++)
            if (p[d] == 0 || p[d] == M[d] - 1) return true;
        return false;
    };
    for (int k = 0; k < M[2]; k++)
        for (int j = 0; j < M[1]; j++)
            for (int i = 0; i < M[0]; i++) {
                const int p[3] = {i, j, k};
                const long n = i + st[1] * j + st[2] * k;
                double v = mg->diag * x[n];
                if (!bd(p))
                    for (int d = 0; d < mg->g.dim; d++)
                        for (int sgn = -1; sgn <= 1; sgn += 2) {
                            int q[3] = {i, j, k};
                            q[d] += sgn;
                            if (!bd(q)) v -= mg->c[d] * x[n + sgn * st[d]];
                        }
                y[n] = v;
            }
    return 0;
}
int p4b_cg_solve(p4b_mg *mg, int pc_type, const double *b, double *x, double rtol, double abstol, int max_it, p4b_ksp_result *res) {
    const size_t n = (size_t)mg->g.mx * (mg->g.dim >= 2 ? mg->g.my : 1) * (mg->g.dim >= 3 ? mg->g.mz : 1);
    std::vector<double> r(b, b + n), z(n), p(n), w(n);
    HostOps o;
    memset(res, 0, sizeof *res);
    memset(x, 0, sizeof(double) * n);
    const double idiag = pc_type == P4B_PC_NONE ? 1.0 : 1.0 / mg->diag;       // (mg: Jacobi stands in, see above)
    for (size_t i = 0; i < n; i++) z[i] = idiag * r[i];
    double znorm = o.norm2(n, z.data()), beta = o.dot(n, r.data(), z.data());
    res->rnorm0 = znorm;
    res->hist[res->nhist++] = znorm;
    const double ttol = std::max(rtol * znorm, abstol);
    p = z;
    res->reason = P4B_DIVERGED_ITS;
    if (znorm <= ttol) res->reason = P4B_CONVERGED_ATOL;
    while (res->reason == P4B_DIVERGED_ITS && res->its < max_it) {
        p4b_mg_matmult(mg, p.data(), w.data());
        const double a = beta / o.dot(n, p.data(), w.data());
        o.axpy(n, a, p.data(), x);
        o.axpy(n, -a, w.data(), r.data());
        for (size_t i = 0; i < n; i++) z[i] = idiag * r[i];
        znorm = o.norm2(n, z.data());
        res->its++;
        if (res->nhist < P4B_MAX_HIST) res->hist[res->nhist++] = znorm;
        if (znorm <= ttol) { res->reason = znorm <= abstol ? P4B_CONVERGED_ATOL : P4B_CONVERGED_RTOL; break; }
        const double bnew = o.dot(n, r.data(), z.data());
        o.aypx(n, bnew / beta, z.data(), p.data());
        beta = bnew;
    }
    res->rnorm = znorm;
    return 0;
}

// ---- assembled matrices: CSR kept as given (the device library converts to SELL-32; the product of y = A x is the same) ----
int p4b_sell_create(p4b_ctx *, int nrows, const int *rowptr, const int *colind, const double *vals, p4b_sell **A) {
    p4b_sell *S = new p4b_sell;
    S->n = nrows;
    S->rp.assign(rowptr, rowptr + nrows + 1);
    S->ci.assign(colind, colind + rowptr[nrows]);
    S->v.assign(vals, vals + rowptr[nrows]);
    *A = S;
    return 0;
}
int p4b_sell_spmv(p4b_sell *A, const double *x, double *y) {
    for (int r = 0; r < A->n; r++) {
        double s2 = 0.0;
        for (int k = A->rp[r]; k < A->rp[r + 1]; k++) s2 += A->v[k] * x[A->ci[k]];
        y[r] = s2;
    }
    return 0;
}
int p4b_sell_destroy(p4b_sell *A) { delete A; return 0; }

}  // extern "C"
