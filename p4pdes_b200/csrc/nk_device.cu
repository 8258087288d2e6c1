// nk_device.cu -- p4b_minimal_solve: the Newton-Krylov-multigrid host logic of nk_solver.hpp with every vector
// operation a CUDA kernel of this library (vectors in HBM, scalars on the host).
//
// Replaces, for `./minimal -snes_fd_color -pc_type mg [-snes_grid_sequence k]` (c/ch8/cluster.sh:70), what the reference
// gets from [PETSc] SNESSolve: see nk_solver.hpp.  The Python host p4pdes_b200/minimal.py runs the same algorithm
// through the individual C-ABI calls; this entry point is the native form of it (one call, no interpreter between
// the kernels) and what a C host such as the PETSc-shaped shim binds.
#include <math.h>

#include "kernels.h"
#include "nk_solver.hpp"
#include "ts_solver.hpp"

namespace p4b {

// defined in mg.cu (the context owns the stream, the reduction scratch and the communicator)
cudaStream_t ctx_stream(p4b_ctx *c);
int ctx_rank(p4b_ctx *c);
int ctx_nranks(p4b_ctx *c);
int ctx_ring_halo(p4b_ctx *c, double *owned, size_t rowlen, int nrows);
int ctx_allgather(p4b_ctx *c, double *full, size_t count);

// process-wide TS step monitor (p4b_set_ts_step_monitor)
static p4b_ts_step_fn g_ts_step_fn = nullptr;
static void *g_ts_step_user = nullptr;
static double g_ts_time_step = 0.0;      // the step the integrator is about to take / proposes (p4b_ts_time_step)

struct DeviceOps {
    p4b_ctx *c;
    cudaStream_t st;
    int err = 0;
    std::vector<double> ts_host;
    void ts_step_n(int k, double t, const double *Y, size_t nloc) {
        if (!g_ts_step_fn || err) return;
        ts_host.resize(nloc);
        to_host(Y, ts_host.data(), nloc);
        if (!err && g_ts_step_fn(g_ts_step_user, k, t, ts_host.data(), nloc)) err = 66;
    }
    void ts_step(int k, double t, const double *Y, size_t n) { ts_step_n(k, t, Y, n); }
    void set_step_size(double h) { g_ts_time_step = h; }
    int error() const { return err; }
    void chk(int rc) { if (rc && !err) err = rc; }
    void cu(cudaError_t e) { if (e != cudaSuccess && !err) err = 70 + (int)e % 20; }

    double *alloc(size_t n) {
        double *p = nullptr;
        cu(cudaMallocAsync((void **)&p, sizeof(double) * (n ? n : 1), st));
        return p;
    }
    void release(double *p) { if (p) cu(cudaFreeAsync(p, st)); }
    void to_host(const double *s, double *d, size_t n) {
        cu(cudaMemcpyAsync(d, s, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        cu(cudaStreamSynchronize(st));
    }
    void from_host(const double *s, double *d, size_t n) {
        cu(cudaMemcpyAsync(d, s, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        cu(cudaStreamSynchronize(st));          // the host buffer may go away right after
    }
    double dot(size_t n, const double *x, const double *y) { double r = NAN; chk(p4b_vec_dot(c, n, x, y, &r)); return r; }
    bool gmres_cgs() const;
    void mdot(size_t n, int k, const double *const *X, const double *y, double *res) { chk(p4b_vec_mdot(c, n, k, X, y, res)); }
    double norm2(size_t n, const double *x) { double r = NAN; chk(p4b_vec_norm2(c, n, x, &r)); return r; }
    double norminf(size_t n, const double *x) { double r = NAN; chk(p4b_vec_norminf(c, n, x, &r)); return r; }
    void axpy(size_t n, double a, const double *x, double *y) { chk(p4b_vec_axpy(c, n, a, x, y)); }
    void aypx(size_t n, double a, const double *x, double *y) { chk(p4b_vec_aypx(c, n, a, x, y)); }
    void axpby(size_t n, double a, const double *x, double b, const double *y, double *out) { chk(p4b_vec_axpby(c, n, a, x, b, y, out)); }
    void copy(size_t n, const double *x, double *y) { chk(p4b_vec_copy(c, n, x, y)); }
    void set(size_t n, double a, double *y) { chk(p4b_vec_set(c, n, a, y)); }
    void minimal_sample(int mx, int my, int problem, double H, double cc, double *g) { chk(p4b_minimal_sample(c, mx, my, problem, H, cc, g)); }
    void minimal_function(int mx, int my, double q, const double *u, const double *g, double *F) { chk(p4b_minimal_function(c, mx, my, q, u, g, F)); }
    // the matrix minimal.c registers (Poisson2DJacobianLocal on the unit square, cx = cy = 1: minimal.c:77-78,136,142-145)
    void poisson_stencil9(int mx, int my, double *vals) { chk(p4b_poisson_stencil9(c, mx, my, 1.0, 1.0, 1.0, 1.0, vals)); }
    void minimal_jacobian_fd(int mx, int my, double q, const double *u, const double *g, const double *F0, double *vals) {
        chk(p4b_minimal_jacobian_fd(c, mx, my, q, u, g, F0, vals));
    }
    void stencil9_apply(int mx, int my, const double *vals, const double *x, double *y) { chk(p4b_stencil9_apply(c, mx, my, vals, x, y)); }
    void stencil9_lin(int mx, int my, const double *vals, const double *u, const double *b, const double *pm1, double ca, double cb,
                      double cg, int jacobi, double *out) {
        chk(p4b_stencil9_lin(c, mx, my, vals, u, b, pm1, ca, cb, cg, jacobi, out));
    }
    double stencil9_gershgorin(int mx, int my, const double *vals, double *work) {
        double r = NAN;
        chk(p4b_stencil9_gershgorin(c, mx, my, vals, work, &r));
        return r;
    }
    void inject2d(int cmx, int cmy, const double *uf, double *uc) { chk(p4b_inject2d(c, cmx, cmy, uf, uc)); }
    void user_monitor(int, int, int, double, int, const double *) {}      // only the callback form has one (CallbackOps)
    bool verify_converged(int, int, const double *, const double *) { return true; }   // only ModelOps has something to verify
    static p4b_grid grid2d(int mx, int my) {
        p4b_grid g;
        g.dim = 2; g.mx = mx; g.my = my; g.mz = 1;
        g.Lx = g.Ly = g.Lz = 1.0;
        g.cx = g.cy = g.cz = 1.0;
        return g;
    }
    void restrict2d(int fmx, int fmy, const double *rf, double *bc) { p4b_grid g = grid2d(fmx, fmy); chk(p4b_restrict(c, &g, rf, bc)); }
    void prolong_add2d(int fmx, int fmy, const double *xc, double *xf) { p4b_grid g = grid2d(fmx, fmy); chk(p4b_prolong_add(c, &g, xc, xf)); }
    void initial_state2d(int mx, int my, const double *g, double *u) { p4b_grid gr = grid2d(mx, my); chk(p4b_initial_state(c, &gr, g, 1, u)); }
    void dense_matvec(int n, const double *Ainv, const double *b, double *x) { chk(p4b_dense_matvec(c, n, Ainv, b, x)); }
    // base-grid inverse from the band LU factors (host): the n column solves run on the device
    void band_inverse(int n, int bw, const std::vector<double> &B, double *Ainv) {
        double *dB = alloc(B.size());
        from_host(B.data(), dB, B.size());
        chk(launch_band_inverse(st, n, bw, dB, Ainv));
        release(dB);
    }
    // pattern.c
    typedef nk::PatternOpts PO;
    double wrms2(size_t n, const double *x, const double *y, double atol, double rtol) {
        double r = NAN;
        chk(p4b_vec_wrms2(c, n, x, y, atol, rtol, &r));
        return r;
    }
    void pattern_initial_state(int mx, int my, double L, double *Y) { chk(p4b_pattern_initial_state(c, mx, my, L, Y)); }
    void pattern_ifunction(int m, const PO &o, const double *Y, const double *Ydot, double *F) {
        chk(p4b_pattern_ifunction(c, m, m, o.L, o.Du, o.Dv, Y, Ydot, F));
    }
    void pattern_rhsfunction(int m, const PO &o, const double *Y, double *G) { chk(p4b_pattern_rhsfunction(c, m, m, o.phi, o.kappa, Y, G)); }
    void pattern_jac_apply(int m, const PO &o, double shift, const double *Y, const double *X, double *out) {
        chk(p4b_pattern_jac_apply(c, m, m, o.L, o.Du, o.Dv, o.phi, o.kappa, shift, Y, X, out));
    }
    void pattern_jac_lin(int m, const PO &o, double shift, const double *Y, const double *X, const double *b, const double *pm1,
                         double ca, double cb, double cg, int jacobi, double *out) {
        chk(p4b_pattern_jac_lin(c, m, m, o.L, o.Du, o.Dv, o.phi, o.kappa, shift, Y, X, b, pm1, ca, cb, cg, jacobi, out));
    }
    double pattern_jac_gershgorin(int m, const PO &o, double shift, const double *Y, double *work) {
        double r = NAN;
        chk(p4b_pattern_jac_gershgorin(c, m, m, o.L, o.Du, o.Dv, o.phi, o.kappa, shift, Y, work, &r));
        return r;
    }
    void pattern_restrict(int Mx, int My, const double *rf, double *bc) { chk(p4b_pattern_restrict(c, Mx, My, rf, bc)); }
    void pattern_prolong_add(int Mx, int My, const double *xc, double *xf) { chk(p4b_pattern_prolong_add(c, Mx, My, xc, xf)); }
    void pattern_inject(int Mx, int My, const double *yf, double *yc) { chk(p4b_pattern_inject(c, Mx, My, yf, yc)); }
    void set_linearisation(const double *) {}             // the kernels take the state as an argument
    void set_time(double) {}                              // the model is autonomous; callbacks get the stage time
};

extern long long g_gmres_cgs;
inline bool DeviceOps::gmres_cgs() const { return g_gmres_cgs != 0; }

// ---------------------------------------------------------------------------------------------------------
// pattern.c on y-slabs (BASELINE config 5: 2048^2 on 8 GPUs).  The DMDA of c/ch5/pattern.c:79-84 is periodic in both
// directions with a box stencil of width 1; [PETSc] splits it over the ranks and DMGlobalToLocal fills one ghost row per
// side from the ring neighbours.  Here: rank r owns rows [r m/P, (r+1) m/P) of every level whose rows divide evenly
// into an even number per rank; the coarser levels (and always the base grid, whose operator is inverted densely) are
// REPLICATED on every rank -- computed redundantly, without communication -- after one all-gather of the restricted
// residual ([PETSc]'s PCREDUNDANT idea applied to whole levels, as the fish hierarchy does).  ts_solver.hpp does not
// know: it passes global sizes (m, n = 2 m^2) and this set of operations translates them --
//   vectors of a distributed level are stored [ghost row | owned rows | ghost row]; the pointer handed out is the first
//   owned row; BLAS-1 operations run over the owned rows, dot products and norms are all-reduced (p4b_vec_*);
//   stencil kernels (F, J x, smoother steps) exchange the operand's ghost rows first and run with ywrap = 0;
//   restriction / injection onto the first replicated level compute this rank's coarse rows in place and all-gather;
//   prolongation from it reads the rank's coarse rows (+ the wrapped row above) out of the replicated copy.
// Every node's arithmetic is the single-GPU kernels' arithmetic on the same neighbour values, so a run differs from the
// one-GPU run only by the rounding of the all-reduced dot products.
// ---------------------------------------------------------------------------------------------------------
struct SlabPlan {
    struct Lev { int m; bool dist; int ys, ym; };
    std::vector<Lev> lev;       // finest first
    int P = 1, rank = 0;
    // levels of the periodic hierarchy m, m/2, ... down to base (ts_solver.hpp StageOperator::create)
    int build(int m, int base, bool mg, int P_, int rank_) {
        P = P_; rank = rank_;
        std::vector<int> sizes{m};
        if (mg) while (sizes.back() > base && sizes.back() % 2 == 0) sizes.push_back(sizes.back() / 2);
        lev.clear();
        bool dist = true;
        for (size_t l = 0; l < sizes.size(); l++) {
            const int ml = sizes[l];
            const bool last = mg && sizes.size() > 1 && l + 1 == sizes.size();
            // a distributed level feeds its coarser level with ym / 2 rows per rank: ym must be even (>= 2)
            if (ml % P != 0 || (ml / P) % 2 != 0 || last) dist = false;
            Lev L;
            L.m = ml; L.dist = dist;
            L.ym = dist ? ml / P : ml;
            L.ys = dist ? rank * (ml / P) : 0;
            lev.push_back(L);
        }
        return lev[0].dist ? 0 : 60;
    }
    const Lev *find(int m) const {
        for (auto &L : lev) if (L.m == m) return &L;
        return nullptr;
    }
    const Lev *find_n(size_t n) const {
        for (auto &L : lev) if ((size_t)2 * L.m * L.m == n) return &L;
        return nullptr;
    }
};

struct SlabPatternOps : DeviceOps {
    SlabPlan plan;
    std::vector<std::pair<double *, double *>> owned_base;      // (pointer handed out, allocation)
    double *tmp = nullptr;
    size_t tmp_cap = 0;
    SlabPatternOps(p4b_ctx *c_, cudaStream_t st_) : DeviceOps{c_, st_} {}
    size_t local_n(size_t n) const {
        const SlabPlan::Lev *L = plan.find_n(n);
        return (L && L->dist) ? (size_t)2 * L->m * L->ym : n;
    }
    double *alloc(size_t n) {
        const SlabPlan::Lev *L = plan.find_n(n);
        if (!L || !L->dist) return DeviceOps::alloc(n);
        const size_t row = (size_t)2 * L->m;
        double *base = DeviceOps::alloc(row * (size_t)(L->ym + 2));
        if (!base) return nullptr;
        cu(cudaMemsetAsync(base, 0, sizeof(double) * row * (size_t)(L->ym + 2), st));
        owned_base.push_back({base + row, base});
        return base + row;
    }
    void release(double *p) {
        for (size_t i = 0; i < owned_base.size(); i++)
            if (owned_base[i].first == p) {
                DeviceOps::release(owned_base[i].second);
                owned_base.erase(owned_base.begin() + i);
                return;
            }
        DeviceOps::release(p);
    }
    double *scratch(size_t n) {
        if (n > tmp_cap) {
            if (tmp) DeviceOps::release(tmp);
            tmp = DeviceOps::alloc(n);
            tmp_cap = n;
        }
        return tmp;
    }
    void finish() { if (tmp) DeviceOps::release(tmp); tmp = nullptr; tmp_cap = 0; }
    // BLAS-1 over the owned rows; reductions are all-reduced inside p4b_vec_* (the context knows its communicator)
    double dot(size_t n, const double *x, const double *y) { return DeviceOps::dot(local_n(n), x, y); }
    void mdot(size_t n, int k, const double *const *X, const double *y, double *res) { DeviceOps::mdot(local_n(n), k, X, y, res); }
    double norm2(size_t n, const double *x) { return DeviceOps::norm2(local_n(n), x); }
    double norminf(size_t n, const double *x) { return DeviceOps::norminf(local_n(n), x); }
    double wrms2(size_t n, const double *x, const double *y, double atol, double rtol) {
        // [PETSc] TSErrorWeightedNorm2 divides the global sum by the global length: the caller does (ts_solver.hpp)
        return DeviceOps::wrms2(local_n(n), x, y, atol, rtol);
    }
    void axpy(size_t n, double a, const double *x, double *y) { DeviceOps::axpy(local_n(n), a, x, y); }
    void aypx(size_t n, double a, const double *x, double *y) { DeviceOps::aypx(local_n(n), a, x, y); }
    void axpby(size_t n, double a, const double *x, double b, const double *y, double *out) { DeviceOps::axpby(local_n(n), a, x, b, y, out); }
    void copy(size_t n, const double *x, double *y) { DeviceOps::copy(local_n(n), x, y); }
    void set(size_t n, double a, double *y) { DeviceOps::set(local_n(n), a, y); }
    void ts_step(int k, double t, const double *Y, size_t n) { ts_step_n(k, t, Y, local_n(n)); }
    void halo(const SlabPlan::Lev &L, const double *v) { chk(ctx_ring_halo(c, const_cast<double *>(v), (size_t)2 * L.m, L.ym)); }
    static void coef(const PO &o, int m, double *Cu, double *Cv) {
        const double h = o.L / (double)m;                      // pattern.c:246
        *Cu = o.Du / (6.0 * h * h);
        *Cv = o.Dv / (6.0 * h * h);
    }
    void pattern_initial_state(int mx, int my, double L, double *Y) {
        const SlabPlan::Lev *lv = plan.find(mx);
        if (!lv || !lv->dist) { DeviceOps::pattern_initial_state(mx, my, L, Y); return; }
        chk(launch_pattern_init(st, mx, my, L, Y, nullptr, 0.0, lv->ys, lv->ym));
    }
    void pattern_ifunction(int m, const PO &o, const double *Y, const double *Ydot, double *F) {
        const SlabPlan::Lev *lv = plan.find(m);
        if (!lv || !lv->dist) { DeviceOps::pattern_ifunction(m, o, Y, Ydot, F); return; }
        double Cu, Cv;
        coef(o, m, &Cu, &Cv);
        halo(*lv, Y);
        chk(launch_pattern_ifunction(st, m, lv->ym, Cu, Cv, 0, 0.0, Y, Ydot, F, 0));
    }
    void pattern_rhsfunction(int m, const PO &o, const double *Y, double *G) {
        const SlabPlan::Lev *lv = plan.find(m);
        if (!lv || !lv->dist) { DeviceOps::pattern_rhsfunction(m, o, Y, G); return; }
        chk(launch_pattern_rhs(st, m * lv->ym, o.phi, o.kappa, Y, G));
    }
    void jac(int mode, int m, const PO &o, double shift, const double *Y, const double *X, const double *b, const double *pm1,
             double ca, double cb, double cg, int jacobi, double *out) {
        const SlabPlan::Lev *lv = plan.find(m);
        double Cu, Cv;
        coef(o, m, &Cu, &Cv);
        if (!lv || !lv->dist) {
            chk(launch_pattern_jac(st, mode, m, m, Cu, Cv, shift, o.phi, o.kappa, Y, X, b, pm1, ca, cb, cg, jacobi, out, 1));
            return;
        }
        if (mode != 2) halo(*lv, X);
        chk(launch_pattern_jac(st, mode, m, lv->ym, Cu, Cv, shift, o.phi, o.kappa, Y, X, b, pm1, ca, cb, cg, jacobi, out, 0));
    }
    void pattern_jac_apply(int m, const PO &o, double shift, const double *Y, const double *X, double *out) {
        jac(0, m, o, shift, Y, X, nullptr, nullptr, 0, 0, 0, 0, out);
    }
    void pattern_jac_lin(int m, const PO &o, double shift, const double *Y, const double *X, const double *b, const double *pm1,
                         double ca, double cb, double cg, int jacobi, double *out) {
        jac(1, m, o, shift, Y, X, b, pm1, ca, cb, cg, jacobi, out);
    }
    double pattern_jac_gershgorin(int m, const PO &o, double shift, const double *Y, double *work) {
        jac(2, m, o, shift, Y, Y, nullptr, nullptr, 0, 0, 0, 0, work);
        return norminf((size_t)2 * m * m, work);          // max over the owned rows, all-reduced (max) over the ranks
    }
    // coarse level Mc x Mc, fine 2 Mc x 2 Mc
    void pattern_restrict(int Mx, int My, const double *rf, double *bc) {
        const SlabPlan::Lev *F = plan.find(2 * Mx), *C = plan.find(Mx);
        if (!F || !C || !F->dist) { DeviceOps::pattern_restrict(Mx, My, rf, bc); return; }
        halo(*F, rf);
        const int ymc = F->ym / 2;
        if (C->dist) { chk(launch_pattern_transfer(st, 0, Mx, ymc, rf, bc, 0)); return; }
        const size_t cnt = (size_t)2 * Mx * ymc;
        chk(launch_pattern_transfer(st, 0, Mx, ymc, rf, bc + (size_t)plan.rank * cnt, 0));
        chk(ctx_allgather(c, bc, cnt));
    }
    void pattern_inject(int Mx, int My, const double *yf, double *yc) {
        const SlabPlan::Lev *F = plan.find(2 * Mx), *C = plan.find(Mx);
        if (!F || !C || !F->dist) { DeviceOps::pattern_inject(Mx, My, yf, yc); return; }
        const int ymc = F->ym / 2;
        if (C->dist) { chk(launch_pattern_transfer(st, 2, Mx, ymc, yf, yc, 0)); return; }
        const size_t cnt = (size_t)2 * Mx * ymc;
        chk(launch_pattern_transfer(st, 2, Mx, ymc, yf, yc + (size_t)plan.rank * cnt, 0));
        chk(ctx_allgather(c, yc, cnt));
    }
    void pattern_prolong_add(int Mx, int My, const double *xc, double *xf) {
        const SlabPlan::Lev *F = plan.find(2 * Mx), *C = plan.find(Mx);
        if (!F || !C || !F->dist) { DeviceOps::pattern_prolong_add(Mx, My, xc, xf); return; }
        const int ymc = F->ym / 2;
        if (C->dist) {
            halo(*C, xc);
            chk(launch_pattern_transfer(st, 1, Mx, ymc, xc, xf, 0));
            return;
        }
        // the rank's coarse rows and the (periodically wrapped) row above them, out of the replicated copy
        const size_t row = (size_t)2 * Mx;
        const int ysc = F->ys / 2;
        double *t = scratch(row * (size_t)(ymc + 1));
        cu(cudaMemcpyAsync(t, xc + (size_t)ysc * row, sizeof(double) * row * (size_t)ymc, cudaMemcpyDeviceToDevice, st));
        cu(cudaMemcpyAsync(t + (size_t)ymc * row, xc + (size_t)((ysc + ymc) % Mx) * row, sizeof(double) * row,
                           cudaMemcpyDeviceToDevice, st));
        chk(launch_pattern_transfer(st, 1, Mx, ymc, t, xf, 0));
    }
};


// The same operations with the RESIDUAL supplied by the caller as a host callback -- the FormFunctionLocal contract of
// the reference's drivers (c/ch7/minimal.c:210-282 is registered with DMDASNESSetFunctionLocal, :138-140): the iterate
// is brought to the host, the callback fills F on the host exactly as PETSc would have it called (whole grid, natural
// ordering), F goes back.  Everything else (Jacobian differencing, Krylov, multigrid, line search algebra) stays on
// the device; the coloured finite-difference Jacobian costs nine callback evaluations per level.
struct CallbackOps : DeviceOps {
    p4b_residual2d_fn fn = nullptr;
    void *user = nullptr;
    p4b_monitor2d_fn mon = nullptr;
    std::vector<double> hu, hF;
    CallbackOps(p4b_ctx *c_, cudaStream_t st_, p4b_residual2d_fn f, p4b_monitor2d_fn m, void *u)
        : DeviceOps{c_, st_}, fn(f), user(u), mon(m) {}
    // [PETSc] SNESMonitorSet (c/ch7/minimal.c:146-148): the caller's monitor sees the current iterate on the host
    void user_monitor(int mx, int my, int its, double fnorm, int tablevel, const double *u) {
        if (!mon || err) return;
        const size_t n = (size_t)mx * my;
        hu.resize(n);
        to_host(u, hu.data(), n);
        if (!err && mon(user, mx, my, its, fnorm, tablevel, hu.data())) err = 66;
    }
    void minimal_sample(int, int, int, double, double, double *) {}
    void minimal_function(int mx, int my, double, const double *u, const double *, double *F) {
        const size_t n = (size_t)mx * my;
        hu.resize(n);
        hF.resize(n);
        to_host(u, hu.data(), n);
        if (!err) {
            const int rc = fn(user, mx, my, hu.data(), hF.data());
            if (rc && !err) err = 65;
        }
        from_host(hF.data(), F, n);
    }
    void minimal_jacobian_fd(int mx, int my, double q, const double *u, const double *g, const double *F0, double *vals) {
        const size_t n = (size_t)mx * my;
        double *up = alloc(n), *Fp = alloc(n);
        const double h = fd_step_wp(norm2(n, u));              // [PETSc] MatFDColoring "wp": one step for every column
        cu(cudaMemsetAsync(vals, 0, sizeof(double) * 9 * n, st));
        for (int cj = 0; cj < 3; cj++)
            for (int ci = 0; ci < 3; ci++) {
                chk(launch_fd_perturb(st, mx, my, ci, cj, h, u, up));
                minimal_function(mx, my, q, up, g, Fp);
                chk(launch_fd_extract(st, mx, my, ci, cj, h, F0, Fp, vals));
            }
        release(up);
        release(Fp);
    }
};

// The time-stepping operations with F(t, Y, Ydot) and G(t, Y) supplied by the caller as HOST callbacks (the
// DMDATSSetIFunctionLocal / DMDATSSetRHSFunctionLocal contract, c/ch5/pattern.c:103-114) and NO Jacobian: the stage operator
// is the differenced residual ([PETSc] MatMFFD "wp", as -snes_mf does),
//     J X = d/de [ F(Y + e X, shift (Y + e X)) - G(Y + e X) ]   (F = M Ydot + f(Y) with a constant M: the caller checks that)
// with the vectors and the Krylov / Newton / controller algebra on the device.  No assembled matrix means no multigrid:
// -pc_type none only.
struct CallbackPatternOps : DeviceOps {
    p4b_ifunction2d_fn ifn = nullptr;
    p4b_rhsfunction2d_fn gfn = nullptr;
    void *user = nullptr;
    const double *lin = nullptr;                           // linearisation point of the next operator products
    std::vector<double> hY, hD, hF;
    double *wY = nullptr, *wD = nullptr, *wF = nullptr, *wR0 = nullptr;
    const double *r0_for = nullptr;
    double r0_shift = 0.0;
    bool r0_rhs = false;
    long long callbacks = 0;
    double tcur = 0.0;                                     // stage time handed to the callbacks (ts_solver.hpp set_time)
    size_t nfield = 0;                                     // != 0: a field of that many doubles (p4b_ts_solve_callbacks), m is 0
    size_t fsize(int m) const { return nfield ? nfield : (size_t)2 * m * m; }
    void set_time(double t) { tcur = t; }
    CallbackPatternOps(p4b_ctx *c_, cudaStream_t st_, p4b_ifunction2d_fn f, p4b_rhsfunction2d_fn g, void *u)
        : DeviceOps{c_, st_}, ifn(f), gfn(g), user(u) {}
    void free_work() { for (double *p : {wY, wD, wF, wR0}) release(p); wY = wD = wF = wR0 = nullptr; }
    void pattern_ifunction(int m, const PO &, const double *Y, const double *Ydot, double *F) {
        const size_t n = fsize(m);
        hY.resize(n); hD.resize(n); hF.resize(n);
        to_host(Y, hY.data(), n);
        to_host(Ydot, hD.data(), n);
        callbacks++;
        if (!err && ifn(user, m, tcur, hY.data(), hD.data(), hF.data())) err = 65;
        from_host(hF.data(), F, n);
    }
    void pattern_rhsfunction(int m, const PO &, const double *Y, double *G) {
        const size_t n = fsize(m);
        hY.resize(n); hF.resize(n);
        to_host(Y, hY.data(), n);
        callbacks++;
        if (!err && gfn(user, m, tcur, hY.data(), hF.data())) err = 65;
        from_host(hF.data(), G, n);
    }
    void set_linearisation(const double *Y) { lin = Y; r0_for = nullptr; }
    // R(W) = F(W, shift W) - [rhs ? G(W) : 0]
    void resid(int m, const PO &o, double shift, bool rhs, const double *W, double *out) {
        const size_t n = fsize(m);
        axpby(n, shift, W, 0.0, nullptr, wD);
        pattern_ifunction(m, o, W, wD, out);
        if (rhs) { pattern_rhsfunction(m, o, W, wF); axpy(n, -1.0, wF, out); }
    }
    void pattern_jac_apply(int m, const PO &o, double shift, const double *Y, const double *X, double *out) {
        const size_t n = fsize(m);
        const bool rhs = Y != nullptr;                     // (nullptr: IMEX or -ptn_no_rhsjacobian, G' stays out)
        if (!lin) { if (!err) err = 68; return; }
        if (!wY) { wY = alloc(n); wD = alloc(n); wF = alloc(n); wR0 = alloc(n); }
        if (r0_for != lin || r0_shift != shift || r0_rhs != rhs) {          // R(Y): once per linearisation point
            resid(m, o, shift, rhs, lin, wR0);
            r0_for = lin; r0_shift = shift; r0_rhs = rhs;
        }
        const double xn = norm2(n, X);
        if (xn == 0.0) { set(n, 0.0, out); return; }
        const double h = 1.4901161193847656e-08 * sqrt(1.0 + norm2(n, lin)) / xn;
        axpby(n, 1.0, lin, h, X, wY);
        resid(m, o, shift, rhs, wY, out);
        axpby(n, 1.0 / h, out, -1.0 / h, wR0, out);
    }
};

// c/ch5/heat.c device-resident: G is the library's kernel (recognised by the caller: the shim probes the registered
// FormRHSFunctionLocal against it), F = Ydot (no IFunction), and the stage operator shift I - dG/du is applied matrix-free by
// the same kernel without its data -- exact, where the callback route differences the residual.  One level: -pc_type none.
struct HeatOps : DeviceOps {
    int mx, my;
    double D0;
    HeatOps(p4b_ctx *c_, cudaStream_t st_, int mx_, int my_, double D0_) : DeviceOps{c_, st_}, mx(mx_), my(my_), D0(D0_) {}
    void pattern_ifunction(int, const PO &, const double *, const double *Ydot, double *F) { copy((size_t)mx * my, Ydot, F); }
    void pattern_rhsfunction(int, const PO &, const double *Y, double *G) { chk(launch_heat_rhs(st, mx, my, D0, Y, G)); }
    void pattern_jac_apply(int, const PO &, double shift, const double *Y, const double *X, double *out) {
        // Y == nullptr: the right-hand side is explicit (IMEX) and stays out of the stage matrix, which is then shift I
        if (Y) chk(launch_heat_jac_apply(st, mx, my, D0, shift, X, out));
        else axpby((size_t)mx * my, shift, X, 0.0, nullptr, out);
    }
};

}  // namespace p4b

using namespace p4b;

extern "C" int p4b_minimal_default_opts(p4b_minimal_opts *o) {
    if (!o) return fail(62, "null options");
    static_assert(sizeof(p4b_minimal_opts) == sizeof(nk::MinimalOpts), "p4b_minimal_opts and nk::MinimalOpts must agree");
    nk::default_opts(reinterpret_cast<nk::MinimalOpts *>(o));
    return 0;
}

extern "C" int p4b_minimal_solve(p4b_ctx *c, const p4b_minimal_opts *opts, p4b_line_fn line, void *line_ctx, double *u_out,
                                 size_t u_capacity, p4b_minimal_result *result) {
    static_assert(sizeof(p4b_minimal_result) == sizeof(nk::MinimalResult), "p4b_minimal_result and nk::MinimalResult must agree");
    static_assert(sizeof(p4b_minimal_stage) == sizeof(nk::StageResult), "p4b_minimal_stage and nk::StageResult must agree");
    if (!c || !opts || !result) return fail(62, "p4b_minimal_solve: null argument");
    const nk::MinimalOpts &o = *reinterpret_cast<const nk::MinimalOpts *>(opts);
    if (o.problem != 0 && o.problem != 1) return fail(5, "unknown problem type");                                   // minimal.c:127
    if (o.problem == 0 && o.exact_init) return fail(2, "initialization with exact solution only possible for -mse_problem catenoid");
    if (o.problem == 1 && o.catenoid_c < 1.0) return fail(3, "catenoid exact solution only valid if c >= 1");       // :116
    if (o.exact_init && o.q != -0.5) return fail(4, "initialization with catenoid exact solution only possible if q=-0.5");
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    DeviceOps ops{c, ctx_stream(c)};
    nk::Printer pr{line, line_ctx};
    double *u = nullptr;
    nk::MinimalResult &R = *reinterpret_cast<nk::MinimalResult *>(result);
    int rc = nk::minimal_solve(&ops, o, pr, u_out ? &u : nullptr, &R);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc && u_out) {
        const size_t n = (size_t)R.mx * R.my;
        if (u_capacity < n) rc = 63;
        else if (cudaMemcpyAsync(u_out, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ops.st) != cudaSuccess) rc = 70;
    }
    if (u) cudaFreeAsync(u, ops.st);
    cudaStreamSynchronize(ops.st);
    if (rc == 61) return fail(61, "base grid of the multigrid hierarchy is larger than 65 x 65: use a coarser -da_grid_x/_y");
    if (rc == 62) return fail(62, "base-grid Jacobian is singular");
    if (rc == 63) return fail(63, "u_out holds %zu doubles, the final grid needs %d x %d", u_capacity, R.mx, R.my);
    if (rc) return fail(rc, "p4b_minimal_solve failed (%s)", p4b_last_error());
    return 0;
}

extern "C" int p4b_snes2d_solve(p4b_ctx *c, const p4b_minimal_opts *opts, p4b_residual2d_fn residual, void *user,
                                const double *u0_host, p4b_line_fn line, void *line_ctx, double *u_out_host,
                                size_t u_capacity, p4b_minimal_result *result) {
    return p4b_snes2d_solve_monitored(c, opts, residual, nullptr, user, u0_host, line, line_ctx, u_out_host, u_capacity, result);
}

namespace p4b {
long long g_recognise_residual = 1;     // p4b_tune("recognise_residual", 0): always evaluate the caller's residual on the host
long long g_gmres_cgs = 0;              // p4b_tune("gmres_cgs", 1): classical Gram-Schmidt with the dots of a step batched
}
static int g_snes2d_route = 0;

extern "C" int p4b_snes2d_last_route(void) { return g_snes2d_route; }

extern "C" int p4b_snes2d_solve_monitored(p4b_ctx *c, const p4b_minimal_opts *opts, p4b_residual2d_fn residual,
                                          p4b_monitor2d_fn monitor, void *user, const double *u0_host, p4b_line_fn line,
                                          void *line_ctx, double *u_out_host, size_t u_capacity, p4b_minimal_result *result) {
    if (!c || !opts || !residual || !u0_host || !result) return fail(62, "p4b_snes2d_solve: null argument");
    const nk::MinimalOpts &o = *reinterpret_cast<const nk::MinimalOpts *>(opts);
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    nk::Printer pr{line, line_ctx};
    double *u = nullptr;
    nk::MinimalResult &R = *reinterpret_cast<nk::MinimalResult *>(result);
    cudaStream_t st = ctx_stream(c);
    int rc = 0;
    g_snes2d_route = 0;
    // is the caller's residual the one this library has as a kernel (c/ch7/minimal.c:210-282)?  nk_solver.hpp, "Recognising
    // the caller's residual": probe it on every grid of the solve; on agreement the residual stays on the device
    DeviceOps base{c, st};
    nk::ProbedModel model;
    auto resid = [&](int mx, int my, const double *uh, double *Fh) { return residual(user, mx, my, uh, Fh); };
    if (g_recognise_residual && nk::probe_minimal_model(&base, resid, o, &model) && !base.error()) {
        nk::ModelOps<DeviceOps> ops(base);
        ops.model = &model;
        if (monitor)
            ops.monitor = [&](int mx, int my, int its, double fnorm, int tab, const double *uh) {
                return monitor(user, mx, my, its, fnorm, tab, uh);
            };
        ops.callback = resid;
        nk::MinimalOpts o2 = o;
        o2.q = model.q;
        g_snes2d_route = 1;
        rc = nk::minimal_solve(&ops, o2, pr, u_out_host ? &u : nullptr, &R, u0_host, false);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && u_out_host) {
            const size_t n = (size_t)R.mx * R.my;
            if (u_capacity < n) rc = 63;
            else ops.to_host(u, u_out_host, n);
        }
        if (rc == 68) {
            // the callback is NOT the model where the solve went: say so and solve again with the callback itself
            fprintf(stderr, "[p4b200] SNES: the residual callback matched the library's kernel at the probes but not at a "
                            "converged iterate (deviation %.3e): solving again with the callback evaluated on the host\n",
                    ops.verify_worst);
            if (u) { cudaFreeAsync(u, st); u = nullptr; }
            g_snes2d_route = 2;
            rc = 0;
        }
    }
    if (g_snes2d_route != 1) {
        const bool again = g_snes2d_route == 2;
        g_snes2d_route = 0;
        if (!again && base.error()) return fail(base.error(), "p4b_snes2d_solve: probing the residual failed (%s)", p4b_last_error());
        CallbackOps ops(c, st, residual, monitor, user);
        rc = nk::minimal_solve(&ops, o, pr, u_out_host ? &u : nullptr, &R, u0_host, false);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && u_out_host) {
            const size_t n = (size_t)R.mx * R.my;
            if (u_capacity < n) rc = 63;
            else ops.to_host(u, u_out_host, n);
        }
    }
    if (u) cudaFreeAsync(u, st);
    cudaStreamSynchronize(st);
    if (rc == 61) return fail(61, "base grid of the multigrid hierarchy is larger than 65 x 65: use a coarser base grid");
    if (rc == 62) return fail(62, "base-grid Jacobian is singular");
    if (rc == 63) return fail(63, "u_out holds %zu doubles, the final grid needs %d x %d", u_capacity, R.mx, R.my);
    if (rc == 65) return fail(65, "the residual callback returned an error");
    if (rc == 66) return fail(66, "the monitor callback returned an error");
    if (rc) return fail(rc, "p4b_snes2d_solve failed (%s)", p4b_last_error());
    return 0;
}

extern "C" int p4b_ts2d_solve(p4b_ctx *c, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction,
                              p4b_rhsfunction2d_fn rhsfunction, void *user, double *Y_inout_host, size_t Y_capacity,
                              p4b_line_fn line, void *line_ctx, p4b_pattern_result *result) {
    if (!c || !opts || !ifunction || !rhsfunction || !Y_inout_host || !result) return fail(62, "p4b_ts2d_solve: null argument");
    const nk::PatternOpts &o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    if ((o.grid_x << o.refine) != (o.grid_y << o.refine)) return fail(1, "the device path needs mx == my");
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_BDF) return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3)");
    if (o.pc_type != nk::PC_NONE)
        return fail(56, "p4b_ts2d_solve: callbacks without a Jacobian give a matrix-free stage operator; no matrix, no multigrid: "
                        "-pc_type none only");
    const int m = o.grid_x << o.refine;
    const size_t n = (size_t)2 * m * m;
    if (Y_capacity < n) return fail(63, "Y holds %zu doubles, the grid needs 2 x %d x %d", Y_capacity, m, m);
    CallbackPatternOps ops(c, ctx_stream(c), ifunction, rhsfunction, user);
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr, *Y0 = ops.alloc(n);
    ops.from_host(Y_inout_host, Y0, n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, &Y, &R, Y0);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) ops.to_host(Y, Y_inout_host, n);
    ops.release(Y0);
    if (Y) ops.release(Y);
    ops.free_work();
    cudaStreamSynchronize(ops.st);
    if (rc == 64) return fail(64, "TSSolve: a nonlinear (stage) solve did not converge");
    if (rc == 65) return fail(65, "a callback returned an error");
    if (rc) return fail(rc, "p4b_ts2d_solve failed (%s)", p4b_last_error());
    return 0;
}

extern "C" int p4b_ts_solve_callbacks(p4b_ctx *c, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction,
                                      p4b_rhsfunction2d_fn rhsfunction, void *user, double *Y_inout_host, size_t n,
                                      p4b_line_fn line, void *line_ctx, p4b_pattern_result *result) {
    if (!c || !opts || !ifunction || !rhsfunction || !Y_inout_host || !result || !n)
        return fail(62, "p4b_ts_solve_callbacks: null argument");
    const nk::PatternOpts &o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_RK)
        return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3), rk (4)");
    if (o.pc_type != nk::PC_NONE && o.ts_type != nk::TS_RK)
        return fail(56, "p4b_ts_solve_callbacks: callbacks without a Jacobian give a matrix-free stage operator; no matrix, no "
                        "multigrid: -pc_type none only");
    nk::PatternOpts o2 = o;
    o2.pc_type = nk::PC_NONE;
    CallbackPatternOps ops(c, ctx_stream(c), ifunction, rhsfunction, user);
    ops.nfield = n;
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr, *Y0 = ops.alloc(n);
    ops.from_host(Y_inout_host, Y0, n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o2, pr, &Y, &R, Y0, n);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) ops.to_host(Y, Y_inout_host, n);
    ops.release(Y0);
    if (Y) ops.release(Y);
    ops.free_work();
    cudaStreamSynchronize(ops.st);
    if (rc == 64) return fail(64, "TSSolve: a stage solve did not converge (or an explicit step produced NaN)");
    if (rc == 65) return fail(65, "a callback returned an error");
    if (rc) return fail(rc, "p4b_ts_solve_callbacks failed (%s)", p4b_last_error());
    return 0;
}

extern "C" double p4b_ts_time_step(void) { return g_ts_time_step; }

extern "C" int p4b_heat_rhs(p4b_ctx *c, int mx, int my, double D0, const double *u, double *G) {
    if (!c || !u || !G) return fail(62, "p4b_heat_rhs: null argument");
    if (mx < 3 || my < 3) return fail(60, "heat grid needs at least 3 nodes per dimension");
    return launch_heat_rhs(ctx_stream(c), mx, my, D0, u, G);
}
extern "C" int p4b_heat_jac_apply(p4b_ctx *c, int mx, int my, double D0, double shift, const double *X, double *JX) {
    if (!c || !X || !JX) return fail(62, "p4b_heat_jac_apply: null argument");
    if (mx < 3 || my < 3) return fail(60, "heat grid needs at least 3 nodes per dimension");
    return launch_heat_jac_apply(ctx_stream(c), mx, my, D0, shift, X, JX);
}
extern "C" int p4b_heat_solve(p4b_ctx *c, const p4b_pattern_opts *opts, int mx, int my, double D0, double *Y_inout_host,
                              p4b_line_fn line, void *line_ctx, p4b_pattern_result *result) {
    if (!c || !opts || !Y_inout_host || !result) return fail(62, "p4b_heat_solve: null argument");
    if (mx < 3 || my < 3) return fail(60, "heat grid needs at least 3 nodes per dimension");
    nk::PatternOpts o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_RK)
        return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3), rk (4)");
    if (o.pc_type != nk::PC_NONE && o.ts_type != nk::TS_RK)
        return fail(56, "p4b_heat_solve: the stage operator is applied matrix-free on one level: -pc_type none only");
    o.pc_type = nk::PC_NONE;
    o.no_rhsjacobian = 0;
    const size_t n = (size_t)mx * my;
    HeatOps ops(c, ctx_stream(c), mx, my, D0);
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr, *Y0 = ops.alloc(n);
    ops.from_host(Y_inout_host, Y0, n);
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = nk::pattern_solve(&ops, o, pr, &Y, &R, Y0, n);
    if (!rc && ops.error()) rc = ops.error();
    if (!rc) ops.to_host(Y, Y_inout_host, n);
    ops.release(Y0);
    if (Y) ops.release(Y);
    cudaStreamSynchronize(ops.st);
    if (rc == 64) return fail(64, "TSSolve: a stage solve did not converge (or an explicit step produced NaN)");
    if (rc) return fail(rc, "p4b_heat_solve failed (%s)", p4b_last_error());
    return 0;
}

extern "C" int p4b_pattern_default_opts(p4b_pattern_opts *o) {
    if (!o) return fail(62, "null options");
    static_assert(sizeof(p4b_pattern_opts) == sizeof(nk::PatternOpts), "p4b_pattern_opts and nk::PatternOpts must agree");
    nk::default_opts(reinterpret_cast<nk::PatternOpts *>(o));
    return 0;
}

extern "C" int p4b_set_ts_step_monitor(p4b_ts_step_fn fn, void *user) {
    g_ts_step_fn = fn;
    g_ts_step_user = user;
    return 0;
}

// host only: which levels of the periodic hierarchy m, m/2, ... grid_x are distributed over y-slabs on nranks ranks, and
// the rows rank `rank` owns on each (see SlabPlan above).  Arrays of length >= 32; returns the number of levels.
extern "C" int p4b_pattern_slab_plan(int m, int grid_x, int mg, int nranks, int rank, int *level_m, int *distributed, int *ys,
                                     int *ym) {
    if (m < 3 || grid_x < 3 || nranks < 1 || rank < 0 || rank >= nranks) return -fail(62, "p4b_pattern_slab_plan: bad argument");
    SlabPlan pl;
    if (pl.build(m, grid_x, mg != 0, nranks, rank) && nranks > 1)
        return -fail(60, "pattern on %d ranks: the %d rows of the grid must split into an even number of rows per rank", nranks, m);
    int nl = 0;
    for (auto &L : pl.lev) {
        if (nl >= 32) break;
        level_m[nl] = L.m; distributed[nl] = L.dist ? 1 : 0; ys[nl] = L.ys; ym[nl] = L.ym;
        nl++;
    }
    return nl;
}

extern "C" int p4b_pattern_solve(p4b_ctx *c, const p4b_pattern_opts *opts, p4b_line_fn line, void *line_ctx, double *Y_out,
                                 size_t Y_capacity, p4b_pattern_result *result) {
    return p4b_pattern_solve_from(c, opts, nullptr, line, line_ctx, Y_out, Y_capacity, result);
}

extern "C" int p4b_pattern_solve_from(p4b_ctx *c, const p4b_pattern_opts *opts, const double *Y0, p4b_line_fn line,
                                      void *line_ctx, double *Y_out, size_t Y_capacity, p4b_pattern_result *result) {
    static_assert(sizeof(p4b_pattern_result) == sizeof(nk::PatternResult), "p4b_pattern_result and nk::PatternResult must agree");
    if (!c || !opts || !result) return fail(62, "p4b_pattern_solve: null argument");
    const nk::PatternOpts &o = *reinterpret_cast<const nk::PatternOpts *>(opts);
    if (o.grid_x < 3 || o.grid_y < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    if ((o.grid_x << o.refine) != (o.grid_y << o.refine)) return fail(1, "pattern.c requires mx == my");            // pattern.c:89
    if (o.ts_type < nk::TS_ARKIMEX || o.ts_type > nk::TS_BDF) return fail(62, "ts_type: arkimex (0), beuler (1), cn (2), bdf (3)");
    nk::Printer pr{line, line_ctx};
    double *Y = nullptr;
    nk::PatternResult &R = *reinterpret_cast<nk::PatternResult *>(result);
    int rc = 0;
    if (ctx_nranks(c) > 1) {
        // y-slabs: Y0 / Y_out are this rank's rows (2 m * m/P doubles, p4b_pattern_slab), lines come from every rank
        SlabPatternOps ops(c, ctx_stream(c));
        const int m = o.grid_x << o.refine;
        if (ops.plan.build(m, o.grid_x, o.pc_type == nk::PC_MG, ctx_nranks(c), ctx_rank(c)))
            return fail(60, "pattern on %d ranks: the %d rows of the grid must split into an even number of rows per rank",
                        ctx_nranks(c), m);
        rc = nk::pattern_solve(&ops, o, pr, Y_out ? &Y : nullptr, &R, Y0);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && Y_out) {
            const size_t n = ops.local_n((size_t)2 * R.m * R.m);
            if (Y_capacity < n) rc = 63;
            else if (cudaMemcpyAsync(Y_out, Y, sizeof(double) * n, cudaMemcpyDeviceToDevice, ops.st) != cudaSuccess) rc = 70;
        }
        if (Y) ops.release(Y);
        ops.finish();
        cudaStreamSynchronize(ops.st);
    } else {
        DeviceOps ops{c, ctx_stream(c)};
        rc = nk::pattern_solve(&ops, o, pr, Y_out ? &Y : nullptr, &R, Y0);
        if (!rc && ops.error()) rc = ops.error();
        if (!rc && Y_out) {
            const size_t n = (size_t)2 * R.m * R.m;
            if (Y_capacity < n) rc = 63;
            else if (cudaMemcpyAsync(Y_out, Y, sizeof(double) * n, cudaMemcpyDeviceToDevice, ops.st) != cudaSuccess) rc = 70;
        }
        if (Y) cudaFreeAsync(Y, ops.st);
        cudaStreamSynchronize(ops.st);
    }
    if (rc == 61) return fail(61, "base grid of the periodic hierarchy has more than 512 unknowns: use a coarser -da_grid_x/_y");
    if (rc == 62) return fail(62, "base-grid stage Jacobian is singular");
    if (rc == 63) return fail(63, "Y_out holds %zu doubles, the grid needs 2 x %d x %d", Y_capacity, R.m, R.m);
    if (rc == 64) return fail(64, "TSSolve: a nonlinear (stage) solve did not converge");
    if (rc) return fail(rc, "p4b_pattern_solve failed (%s)", p4b_last_error());
    return 0;
}
