// bratu.cu -- the Liouville-Bratu equation of c/ch7/solns/bratu2D.c solved by FAS multigrid with nonlinear Gauss-Seidel
// smoothing: device kernels + the host logic of the cycle behind p4b_bratu_solve.
//
//   bratu_function_kernel   FormFunctionLocal, c/ch7/solns/bratu2D.c:196-225
//   bratu_ngs_kernel        NonlinearGS, :229-299 -- pointwise Newton on phi(u) = F_ij(u) - b_ij with [PETSc] SNESNGS's
//                           tolerances.  The reference sweeps lexicographically (sequential); here the sweep is RED-BLACK
//                           (two half sweeps, every node of a colour in parallel): same fixed point, different iterates --
//                           the north star's rule for smoothers ("Jacobi/Chebyshev ..., since SOR is sequential").
//   cycle                   [PETSc] SNESFAS -snes_fas_type full with the golden's components (oracle/bratu_oracle.py states
//                           what is and is not pinned): F cycle per outer iteration, NGS(2 sweeps) before and after, 4 x
//                           NGS(2) on the coarsest grid, FAS correction x_c0 = inject(x), b_c = F_c(x_c0) - R (F(x) - b),
//                           x += P (x_c - x_c0), R = P^T of the DMDA Q1 interpolation (transfer.cu).
// HBM-bound fp64 like everything else here: a residual moves 16 B/node (24 with b), a red-black sweep two passes.
#include <math.h>
#include <string.h>

#include <functional>
#include <vector>

#include "kernels.h"

namespace p4b {

cudaStream_t ctx_stream(p4b_ctx *c);

__device__ __forceinline__ double bratu_g(int exact, double x, double y) {
    if (!exact) return 0.0;                                   // g_zero, bratu2D.c:36-38
    const double r2 = (x + 1.0) * (x + 1.0) + (y + 1.0) * (y + 1.0), qq = r2 * r2 + 1.0;      // g_liouville, :40-45
    return log(32.0 * r2 / (qq * qq));
}

// F = residual [- b]
__global__ void __launch_bounds__(256) bratu_function_kernel(int mx, int my, double lambda, int exact,
                                                              const double *__restrict__ u, const double *__restrict__ b,
                                                              double *__restrict__ F) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= (long long)mx * my) return;
    const int j = (int)(n / mx), i = (int)(n - (long long)j * mx);
    const double hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1);
    const double uc = u[n];
    double f;
    if (i == 0 || j == 0 || i == mx - 1 || j == my - 1) {
        f = uc - bratu_g(exact, i * hx, j * hy);
    } else {
        f = (hy / hx) * (2.0 * uc - u[n - 1] - u[n + 1]) + (hx / hy) * (2.0 * uc - u[n - mx] - u[n + mx]) -
            hx * hy * lambda * exp(uc);
    }
    F[n] = b ? f - b[n] : f;
}

struct NgsTol { double atol, rtol, stol; int maxits; };

// one half sweep: the interior nodes with (i + j) % 2 == colour.  A thread owns ONE node of that colour (row j, column
// 2 k + ((colour + j) & 1)): every lane of a warp runs the Newton iteration -- with one thread per node of either colour
// half the lanes idled through it, and this kernel is bound by the fp64 exp and divide, not by HBM (measured at
// 20481^2: 0.19 of the HBM roofline).  Boundary nodes are set to g by the colour-0 launch.
__global__ void __launch_bounds__(256) bratu_ngs_kernel(int mx, int my, double lambda, int exact, int colour, NgsTol tol,
                                                         const double *__restrict__ b, double *u) {
    const int half = (mx + 1) / 2;                             // threads per row
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)half * my) return;
    const int j = (int)(t / half), k = (int)(t - (long long)j * half);
    const int i = 2 * k + ((colour + j) & 1);
    const double hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1);
    if (colour == 0) {
        // the boundary nodes of this thread's pair {2k, 2k+1} (whatever their colour): u = g  (bratu2D.c:252-254)
        for (int ii = 2 * k; ii <= 2 * k + 1 && ii < mx; ii++)
            if (ii == 0 || j == 0 || ii == mx - 1 || j == my - 1) u[(long long)j * mx + ii] = bratu_g(exact, ii * hx, j * hy);
    }
    if (i <= 0 || j <= 0 || i >= mx - 1 || j >= my - 1) return;
    const long long n = (long long)j * mx + i;
    const double hyhx = hy / hx, hxhy = hx / hy, dl = hx * hy * lambda, bij = b ? b[n] : 0.0;
    // neighbours on the boundary are read through g, so the result does not depend on whether the thread that sets
    // them has run yet
    auto nb = [&](int ii, int jj, long long q) {
        return (ii == 0 || jj == 0 || ii == mx - 1 || jj == my - 1) ? bratu_g(exact, ii * hx, jj * hy) : u[q];
    };
    const double sx = nb(i - 1, j, n - 1) + nb(i + 1, j, n + 1), sy = nb(i, j - 1, n - mx) + nb(i, j + 1, n + mx);
    double uu = u[n], phi0 = 0.0;
    for (int kk = 0; kk < tol.maxits; kk++) {
        const double e = exp(uu);
        const double phi = hyhx * (2.0 * uu - sx) + hxhy * (2.0 * uu - sy) - dl * e - bij;
        if (kk == 0) phi0 = phi;
        const double s = -phi / (2.0 * (hyhx + hxhy) - dl * e);
        uu += s;
        if (tol.atol > fabs(phi) || tol.rtol * fabs(phi0) > fabs(phi) || tol.stol * fabs(uu) > fabs(s)) break;
    }
    u[n] = uu;
}

__global__ void __launch_bounds__(256) bratu_exact_kernel(int mx, int my, int exact, double *__restrict__ g) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= (long long)mx * my) return;
    const int j = (int)(n / mx), i = (int)(n - (long long)j * mx);
    g[n] = bratu_g(exact, i * (1.0 / (mx - 1)), j * (1.0 / (my - 1)));
}

static const NgsTol NGS_DEFAULT = {1.0e-50, 1.0e-8, 1.0e-8, 50};      // as oracle/bratu_oracle.py restates [PETSc] SNESNGS

static int launch_bratu_function(cudaStream_t st, int mx, int my, double lambda, int exact, const double *u, const double *b,
                                 double *F) {
    const long long N = (long long)mx * my;
    bratu_function_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mx, my, lambda, exact, u, b, F);
    P4B_LAUNCH_CHECK();
    return 0;
}
static int launch_bratu_ngs(cudaStream_t st, int mx, int my, double lambda, int exact, int sweeps, const double *b, double *u) {
    const long long N = (long long)((mx + 1) / 2) * my;          // one thread per node of the colour
    for (int s = 0; s < sweeps; s++)
        for (int colour = 0; colour < 2; colour++) {
            bratu_ngs_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mx, my, lambda, exact, colour, NGS_DEFAULT, b, u);
            P4B_LAUNCH_CHECK();
        }
    return 0;
}

struct BratuLevel {
    int mx = 0, my = 0;
    size_t n = 0;
    double *u = nullptr, *b = nullptr, *r = nullptr, *x0 = nullptr;
    LevelDesc d;
};

}  // namespace p4b

using namespace p4b;

extern "C" {

int p4b_bratu_default_opts(p4b_bratu_opts *o) {
    memset(o, 0, sizeof *o);
    o->lambda = 1.0; o->exact = 0;
    o->grid_x = o->grid_y = 3; o->refine = 0; o->levels = 0;
    o->snes_rtol = 1.0e-8; o->snes_max_it = 10000;         // [PETSc] SNESFAS defaults
    o->smooth_sweeps = 1; o->smooth_its = 1;               // [PETSc] -fas_levels_snes_ngs_sweeps 1, -fas_levels_snes_max_it 1
    o->coarse_sweeps = 1; o->coarse_its = 50;
    o->full_cycle = 0;                                     // -snes_fas_type multiplicative (V cycles)
    return 0;
}

int p4b_bratu_function(p4b_ctx *c, int mx, int my, double lambda, int exact, const double *u, const double *b, double *F) {
    if (mx < 3 || my < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    return launch_bratu_function(ctx_stream(c), mx, my, lambda, exact, u, b, F);
}
int p4b_bratu_ngs(p4b_ctx *c, int mx, int my, double lambda, int exact, int sweeps, const double *b, double *u) {
    if (mx < 3 || my < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    return launch_bratu_ngs(ctx_stream(c), mx, my, lambda, exact, sweeps, b, u);
}
int p4b_bratu_exact(p4b_ctx *c, int mx, int my, int exact, double *g) {
    const long long N = (long long)mx * my;
    bratu_exact_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx_stream(c)>>>(mx, my, exact, g);
    P4B_LAUNCH_CHECK();
    return 0;
}

// ./bratu2D -snes_type fas ... : VecSet(u, 0); SNESSolve (bratu2D.c:131-133); error norm when -lb_exact (:146-157)
int p4b_bratu_solve(p4b_ctx *c, const p4b_bratu_opts *o, p4b_line_fn line, void *line_ctx, double *u_out, size_t u_capacity,
                    p4b_bratu_result *R) {
    if (!c || !o || !R) return fail(62, "p4b_bratu_solve: null argument");
    if (o->grid_x < 3 || o->grid_y < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    if (o->exact && o->lambda != 1.0) return fail(1, "Liouville exact solution only implemented for lambda = 1.0");   // :99-101
    cudaStream_t st = ctx_stream(c);
    memset(R, 0, sizeof *R);
    R->errinf = -1.0;
    // the DM hierarchy: the refined grid and its coarsenings down to the -da_grid base (or -snes_fas_levels of them)
    std::vector<std::pair<int, int>> shapes;
    {
        int mx = o->grid_x, my = o->grid_y;
        std::vector<std::pair<int, int>> up{{mx, my}};
        for (int r = 0; r < o->refine; r++) { mx = 2 * mx - 1; my = 2 * my - 1; up.push_back({mx, my}); }
        const int want = o->levels > 0 ? o->levels : (int)up.size();
        if (want > (int)up.size()) return fail(60, "cannot build %d FAS levels from -da_refine %d", want, o->refine);
        shapes.assign(up.end() - want, up.end());          // coarsest first
    }
    const int nl = (int)shapes.size(), top = nl - 1;
    std::vector<BratuLevel> L(nl);
    int rc = 0;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && !rc) rc = fail(71, "CUDA error %d (%s)", (int)e, cudaGetErrorString(e)); };
    for (int l = 0; l < nl && !rc; l++) {
        BratuLevel &V = L[l];
        V.mx = shapes[l].first; V.my = shapes[l].second;
        V.n = (size_t)V.mx * V.my;
        // stream-ordered allocations: the context keeps the device pool's memory between solves (p4b_ctx_create), so a
        // second solve of the same size pays no allocation (a 20481^2 hierarchy is 11 GB: cudaMalloc/cudaFree of it took
        // longer than the solve)
        cu(cudaMallocAsync((void **)&V.u, sizeof(double) * V.n, st));
        cu(cudaMallocAsync((void **)&V.r, sizeof(double) * V.n, st));
        if (l < top) { cu(cudaMallocAsync((void **)&V.b, sizeof(double) * V.n, st)); cu(cudaMallocAsync((void **)&V.x0, sizeof(double) * V.n, st)); }
        memset(&V.d, 0, sizeof V.d);
        V.d.nx = V.mx; V.d.ny = 1; V.d.nz = V.my; V.d.ax = 1; V.d.ay = 0; V.d.az = 1; V.d.zs = 0; V.d.zm = V.my;
    }
    auto chk = [&](int r_) { if (r_ && !rc) rc = r_; };
    auto say = [&](const char *fmt, auto... a) {
        if (!line) return;
        char buf[256];
        snprintf(buf, sizeof buf, fmt, a...);
        line(buf, line_ctx);
    };
    auto Fl = [&](int l, const double *u, const double *b, double *out) {
        R->residual_calls++;
        chk(launch_bratu_function(st, L[l].mx, L[l].my, o->lambda, o->exact, u, b, out));
    };
    auto smooth = [&](int l, int times, int sweeps) {
        for (int t = 0; t < times; t++) {
            R->ngs_calls++;
            chk(launch_bratu_ngs(st, L[l].mx, L[l].my, o->lambda, o->exact, sweeps, L[l].b, L[l].u));
        }
    };
    // FAS coarse-grid problem of level l (>= 1) from its current iterate: x_c0 = inject(x), b_c = F_c(x_c0) - R (F(x) - b)
    auto descend = [&](int l) {
        BratuLevel &F = L[l], &C = L[l - 1];
        Fl(l, F.u, F.b, F.r);
        chk(launch_inject2d(st, C.mx, C.my, F.mx, F.u, C.x0));
        chk(launch_restrict(st, F.d, C.d, F.r, C.r));                       // C.r = R (F(x) - b)
        Fl(l - 1, C.x0, nullptr, C.b);                                      // C.b = F_c(x_c0)
        chk(launch_axpy(st, (long long)C.n, -1.0, C.r, C.b));
        cu(cudaMemcpyAsync(C.u, C.x0, sizeof(double) * C.n, cudaMemcpyDeviceToDevice, st));
    };
    // x += P (x_c - x_c0)
    auto correct = [&](int l) {
        BratuLevel &F = L[l], &C = L[l - 1];
        chk(launch_axpby_out(st, (long long)C.n, 1.0, C.u, -1.0, C.x0, C.r));
        chk(launch_prolong_add(st, F.d, C.d, C.r, F.u));
    };
    std::function<void(int)> vcycle = [&](int l) {
        if (l == 0) { smooth(0, o->coarse_its, o->coarse_sweeps); return; }
        smooth(l, o->smooth_its, o->smooth_sweeps);
        descend(l);
        vcycle(l - 1);
        correct(l);
        smooth(l, o->smooth_its, o->smooth_sweeps);
    };
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cu(cudaEventCreate(&e0));
    cu(cudaEventCreate(&e1));
    cu(cudaEventRecord(e0, st));
    cu(cudaMemsetAsync(L[top].u, 0, sizeof(double) * L[top].n, st));       // VecSet(u, 0.0), bratu2D.c:132
    double f0 = 0.0, fn = 0.0;
    Fl(top, L[top].u, nullptr, L[top].r);
    chk(p4b_vec_norm2(c, L[top].n, L[top].r, &f0));
    R->fnorm[R->nnorm++] = f0;
    if (o->monitor) say("  0 SNES Function norm %g ", f0);
    int its = 0, reason = 0;
    while (!rc && !reason) {
        if (its >= o->snes_max_it) { reason = -5; break; }
        if (o->full_cycle && top > 0) {
            for (int l = top; l >= 1; l--) descend(l);                      // right-hand sides down the hierarchy
            smooth(0, o->coarse_its, o->coarse_sweeps);
            for (int l = 1; l <= top; l++) { correct(l); vcycle(l); }       // interpolate, one V cycle per level
        } else {
            vcycle(top);
        }
        its++;
        Fl(top, L[top].u, nullptr, L[top].r);
        chk(p4b_vec_norm2(c, L[top].n, L[top].r, &fn));
        if (R->nnorm < 64) R->fnorm[R->nnorm++] = fn;
        if (o->monitor) say("  %d SNES Function norm %g ", its, fn);
        if (fn != fn) reason = -4;
        else if (fn < 1.0e-50) reason = 2;
        else if (fn <= o->snes_rtol * f0) reason = 3;
    }
    cu(cudaEventRecord(e1, st));
    cu(cudaEventSynchronize(e1));
    float ms = 0;
    if (!rc) cudaEventElapsedTime(&ms, e0, e1);
    R->solve_ms = ms;
    R->its = its;
    R->reason = reason;
    R->mx = L[top].mx; R->my = L[top].my;
    if (!rc && o->converged_reason)
        say("Nonlinear solve %s due to %s iterations %d", reason > 0 ? "converged" : "did not converge",
            reason == 3 ? "CONVERGED_FNORM_RELATIVE" : (reason == 2 ? "CONVERGED_FNORM_ABS" : (reason == -5 ? "DIVERGED_MAX_IT" : "DIVERGED_FNORM_NAN")), its);
    if (!rc && o->exact) {                                                  // :146-157
        BratuLevel &T = L[top];
        chk(p4b_bratu_exact(c, T.mx, T.my, 1, T.r));
        chk(launch_axpby_out(st, (long long)T.n, 1.0, T.u, -1.0, T.r, T.r));
        chk(p4b_vec_norminf(c, T.n, T.r, &R->errinf));
    }
    if (!rc && u_out) {
        if (u_capacity < L[top].n) rc = fail(63, "u_out holds %zu doubles, the grid needs %d x %d", u_capacity, L[top].mx, L[top].my);
        else cu(cudaMemcpyAsync(u_out, L[top].u, sizeof(double) * L[top].n, cudaMemcpyDeviceToDevice, st));
    }
    for (BratuLevel &V : L)
        for (double *p : {V.u, V.b, V.r, V.x0})
            if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

}  // extern "C"
