/* fish_cpu.c -- CPU restatement (C99 + OpenMP) of the fish.c CG + geometric-multigrid path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Nothing under p4pdes_b200/ links or calls this; it is
 * used by tests/ (as a checker), by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs.  It is NOT PETSc: PETSc and MPI cannot be installed in this image, so
 * this file restates, one function per PETSc operation, the algorithm the reference executes:
 *
 *   form_function()   c/ch6/poissonfunctions.c:4-115    Poisson{1,2,3}DFunctionLocal
 *   op_apply()        c/ch6/poissonfunctions.c:117-258  the assembled Jacobian, applied matrix-free
 *   u_exact/f_rhs     c/ch6/fish.c:15-82
 *   interp/restrict   [PETSc] DMCreateInterpolation (DMDA Q1), MatRestrict = P^T   
 *   cheb_smooth()     [PETSc] KSPSolve_Chebyshev (first kind)                        
 *   sor_apply()       [PETSc] MatSOR, omega = 1, local symmetric sweep (goldens only)
 *   mcycle()          [PETSc] PCMGMCycle_Private, V or W                              
 *   cg()              [PETSc] KSPSolve_CG, preconditioned norm                        
 *
 * Pinned by tests/test_oracle_c.py: reproduces c/ch6/output/fish.test1 and fish.test3 (iteration counts
 * and error norms, SOR smoothing) and agrees with the NumPy oracle (oracle/fish_oracle.py, itself pinned
 * on all eight goldens) to 1e-12 on Chebyshev/Jacobi runs.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>

/* first touch by the threads that will stream the vector (static schedule, as every loop below): on a multi-socket host
 * a serial calloc / memcpy would put every page on one NUMA node */
static double *par_zeros(long n) {
    double *v = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!v) return NULL;
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; q++) v[q] = 0.0;
    return v;
}
static void par_copy(double *dst, const double *src, long n) {
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; q++) dst[q] = src[q];
}
static void par_zero(double *v, long n) {
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; q++) v[q] = 0.0;
}
#endif

typedef struct {
    int dim, m[3];
    double L[3], c[3];
    int problem;        /* 0 manupoly, 1 manuexp, 2 zero */
    int gonboundary;
    int levels;         /* 0 = down to 3^d */
    int cycle;          /* 1 V, 2 W */
    int smoother_ksp;   /* 0 chebyshev, 1 richardson */
    int smoother_pc;    /* 0 jacobi, 1 sor */
    int smooth_its;
    double emin, emax;  /* explicit Chebyshev bounds if emax > 0 */
    double rtol;
    int max_it;
    int threads;
} fishcpu_opts;

typedef struct {
    int its, nlevels, threads;
    double fnorm0, fnorm1, errinf, err2h, seconds;
    int nhist;
    double hist[256];
} fishcpu_result;

typedef struct {
    int dim, nx, ny, nz, ax, ay, az;
    double cx, cy, cz, diag, vol, hx, hy, hz;
    double emin, emax;
    long n;
    double *x, *b, *r, *t1, *t2;
} level_t;

static int is_bd(const level_t *g, int i, int j, int k) {
    return (g->ax && (i == 0 || i == g->nx - 1)) || (g->ay && (j == 0 || j == g->ny - 1)) ||
           (g->az && (k == 0 || k == g->nz - 1));
}

/* slots as in the device code: 2-D grids are (mx,1,my), 1-D (mx,1,1) */
static void level_init(level_t *g, int dim, const int *m, const double *L, const double *c) {
    memset(g, 0, sizeof *g);
    g->dim = dim;
    const double hx = L[0] / (m[0] - 1);
    const double hy = dim >= 2 ? L[1] / (m[1] - 1) : 1.0;
    const double hz = dim >= 3 ? L[2] / (m[2] - 1) : 1.0;
    if (dim == 1) {
        g->nx = m[0]; g->ny = 1; g->nz = 1; g->ax = 1;
        g->cx = c[0] / hx; g->diag = c[0] * 2.0 / hx; g->vol = hx; g->hx = hx; g->hy = 1; g->hz = 1;
    } else if (dim == 2) {
        g->nx = m[0]; g->ny = 1; g->nz = m[1]; g->ax = 1; g->az = 1;
        g->cx = c[0] * hy / hx; g->cz = c[1] * hx / hy; g->diag = 2.0 * (g->cx + g->cz);
        g->vol = hx * hy; g->hx = hx; g->hy = 1; g->hz = hy;
    } else {
        g->nx = m[0]; g->ny = m[1]; g->nz = m[2]; g->ax = g->ay = g->az = 1;
        const double dvol = hx * hy * hz;
        g->cx = c[0] * dvol / (hx * hx); g->cy = c[1] * dvol / (hy * hy); g->cz = c[2] * dvol / (hz * hz);
        g->diag = 2.0 * (g->cx + g->cy + g->cz);
        g->vol = dvol; g->hx = hx; g->hy = hy; g->hz = hz;
    }
    g->n = (long)g->nx * g->ny * g->nz;
}

static double u_exact(int dim, int problem, double x, double y, double z) {
    if (problem == 0) {
        double a = x * x * (1.0 - x * x);
        if (dim >= 2) a = a * y * y * (y * y - 1.0);
        if (dim >= 3) a = a * z * z * (z * z - 1.0);
        return a;
    }
    if (problem == 1) {
        if (dim == 1) return -exp(x);
        if (dim == 2) return -x * exp(y);
        return -x * exp(y + z);
    }
    return 0.0;
}

static double f_rhs(int dim, int problem, double x, double y, double z, const double *c) {
    if (problem == 0) {
        if (dim == 1) return c[0] * 12.0 * x * x - 2.0;
        const double aa = x * x * (1.0 - x * x), bb = y * y * (y * y - 1.0);
        const double ddaa = 2.0 * (1.0 - 6.0 * x * x), ddbb = 2.0 * (6.0 * y * y - 1.0);
        if (dim == 2) return -(c[0] * ddaa * bb + c[1] * aa * ddbb);
        const double cc = z * z * (z * z - 1.0), ddcc = 2.0 * (6.0 * z * z - 1.0);
        return -(c[0] * ddaa * bb * cc + c[1] * aa * ddbb * cc + c[2] * aa * bb * ddcc);
    }
    if (problem == 1) {
        if (dim == 1) return exp(x);
        if (dim == 2) return x * exp(y);
        return 2.0 * x * exp(y + z);
    }
    return 0.0;
}

static void coords(const level_t *g, int i, int j, int k, double *x, double *y, double *z) {
    const double X0 = i * g->hx, X1 = j * g->hy, X2 = k * g->hz;
    *x = X0;
    *y = g->dim == 3 ? X1 : (g->dim == 2 ? X2 : 0.0);
    *z = g->dim == 3 ? X2 : 0.0;
}

/* y = A u  (MatMult) */
static void op_apply(const level_t *g, const double *u, double *y) {
    const long plane = (long)g->nx * g->ny;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->nz; k++)
        for (int j = 0; j < g->ny; j++) {
            const long row = k * plane + (long)j * g->nx;
            const int jkb = (g->ay && (j == 0 || j == g->ny - 1)) || (g->az && (k == 0 || k == g->nz - 1));
            for (int i = 0; i < g->nx; i++) {
                const long n = row + i;
                double v = g->diag * u[n];
                if (!jkb && !(g->ax && (i == 0 || i == g->nx - 1))) {
                    if (g->ax) v -= g->cx * ((i - 1 > 0 ? u[n - 1] : 0.0) + (i + 1 < g->nx - 1 ? u[n + 1] : 0.0));
                    if (g->ay) v -= g->cy * ((j - 1 > 0 ? u[n - g->nx] : 0.0) + (j + 1 < g->ny - 1 ? u[n + g->nx] : 0.0));
                    if (g->az) v -= g->cz * ((k + 1 < g->nz - 1 ? u[n + plane] : 0.0) + (k - 1 > 0 ? u[n - plane] : 0.0));
                }
                y[n] = v;
            }
        }
}

/* r = b - A u  (MatResidual) */
static void op_residual(const level_t *g, const double *b, const double *u, double *r) {
    op_apply(g, u, r);
#pragma omp parallel for schedule(static)
    for (long n = 0; n < g->n; n++) r[n] = b[n] - r[n];
}

static double dot(long n, const double *a, const double *b) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* PCSOR: z = (D+U)^-1 D (D+L)^-1 r, natural ordering, omega 1 (sequential: goldens only) */
static void sor_apply(const level_t *g, const double *r, double *z) {
    const long plane = (long)g->nx * g->ny;
    for (int pass = 0; pass < 2; pass++) {
        for (long q = 0; q < g->n; q++) {
            const long n = pass == 0 ? q : g->n - 1 - q;
            const int k = (int)(n / plane), j = (int)((n % plane) / g->nx), i = (int)(n % g->nx);
            double s = pass == 0 ? r[n] : g->diag * z[n];   /* backward: rhs is D y */
            if (!is_bd(g, i, j, k)) {
                /* forward uses already-updated lower neighbours; backward uses upper neighbours */
                if (pass == 0) {
                    if (g->ax && i - 1 > 0) s += g->cx * z[n - 1];
                    if (g->ay && j - 1 > 0) s += g->cy * z[n - g->nx];
                    if (g->az && k - 1 > 0) s += g->cz * z[n - plane];
                } else {
                    if (g->ax && i + 1 < g->nx - 1) s += g->cx * z[n + 1];
                    if (g->ay && j + 1 < g->ny - 1) s += g->cy * z[n + g->nx];
                    if (g->az && k + 1 < g->nz - 1) s += g->cz * z[n + plane];
                }
            }
            z[n] = s / g->diag;
        }
    }
}

static void pc_apply(const level_t *g, int pc, const double *r, double *z) {
    if (pc == 1) { sor_apply(g, r, z); return; }
    const double di = 1.0 / g->diag;
#pragma omp parallel for schedule(static)
    for (long n = 0; n < g->n; n++) z[n] = di * r[n];
}

/* KSPChebyshev, `its` preconditioner applications; x in/out; uses g->r, g->t1, g->t2 */
static void cheb_smooth(level_t *g, int pc, int its, const double *b, double *x) {
    const double scale = 2.0 / (g->emax + g->emin), alpha = 1.0 - scale * g->emin;
    const double mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
    double cm1 = 1.0, ck = mu;
    double *pm1 = x, *pk = g->t1, *z = g->t2;
    op_residual(g, b, pm1, g->r);
    pc_apply(g, pc, g->r, z);
#pragma omp parallel for schedule(static)
    for (long n = 0; n < g->n; n++) pk[n] = pm1[n] + scale * z[n];
    for (int i = 1; i < its; i++) {
        op_residual(g, b, pk, g->r);
        const double cp1 = 2.0 * mu * ck - cm1, omega = omegaprod * ck / cp1;
        pc_apply(g, pc, g->r, z);
#pragma omp parallel for schedule(static)
        for (long n = 0; n < g->n; n++) pm1[n] = (1.0 - omega) * pm1[n] + omega * pk[n] + omega * scale * z[n];
        double *t = pm1; pm1 = pk; pk = t;
        cm1 = ck; ck = cp1;
    }
    if (pk != x) par_copy(x, pk, g->n);
}

static void rich_smooth(level_t *g, int pc, int its, const double *b, double *x) {
    for (int i = 0; i < its; i++) {
        op_residual(g, b, x, g->r);
        pc_apply(g, pc, g->r, g->t2);
#pragma omp parallel for schedule(static)
        for (long n = 0; n < g->n; n++) x[n] += g->t2[n];
    }
}

/* b_c = P^T r */
static void restrict_to(const level_t *F, const level_t *C, const double *rf, double *bc) {
    const long fplane = (long)F->nx * F->ny, cplane = (long)C->nx * C->ny;
#pragma omp parallel for collapse(2) schedule(static)
    for (int K = 0; K < C->nz; K++)
        for (int J = 0; J < C->ny; J++)
            for (int I = 0; I < C->nx; I++) {
                const int fi = F->ax ? 2 * I : I, fj = F->ay ? 2 * J : J, fk = F->az ? 2 * K : K;
                double s = 0.0;
                for (int dk = -F->az; dk <= F->az; dk++) {
                    const int kf = fk + dk;
                    if (kf < 0 || kf >= F->nz) continue;
                    for (int dj = -F->ay; dj <= F->ay; dj++) {
                        const int jf = fj + dj;
                        if (jf < 0 || jf >= F->ny) continue;
                        for (int di = -F->ax; di <= F->ax; di++) {
                            const int ifn = fi + di;
                            if (ifn < 0 || ifn >= F->nx) continue;
                            const double w = (dk ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (di ? 0.5 : 1.0);
                            s += w * rf[kf * fplane + (long)jf * F->nx + ifn];
                        }
                    }
                }
                bc[K * cplane + (long)J * C->nx + I] = s;
            }
}

/* x_f += P x_c */
static void prolong_add(const level_t *F, const level_t *C, const double *xc, double *xf) {
    const long fplane = (long)F->nx * F->ny, cplane = (long)C->nx * C->ny;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < F->nz; k++)
        for (int j = 0; j < F->ny; j++)
            for (int i = 0; i < F->nx; i++) {
                const int I0 = F->ax ? i >> 1 : i, oi = F->ax ? i & 1 : 0;
                const int J0 = F->ay ? j >> 1 : j, oj = F->ay ? j & 1 : 0;
                const int K0 = F->az ? k >> 1 : k, ok = F->az ? k & 1 : 0;
                double s = 0.0;
                for (int dk = 0; dk <= ok; dk++)
                    for (int dj = 0; dj <= oj; dj++)
                        for (int di = 0; di <= oi; di++)
                            s += xc[(K0 + dk) * cplane + (long)(J0 + dj) * C->nx + I0 + di];
                const double w = (oi ? 0.5 : 1.0) * (oj ? 0.5 : 1.0) * (ok ? 0.5 : 1.0);
                xf[k * fplane + (long)j * F->nx + i] += w * s;
            }
}

typedef struct {
    int nlev, cycle, sm_ksp, sm_pc, sm_its;
    level_t lev[16];
    double *chol;   /* dense lower Cholesky factor of the coarsest operator */
    long n0;
} mg_t;

static void coarse_solve(const mg_t *M, const double *b, double *x) {
    const long n = M->n0;
    const double *G = M->chol;
    for (long r = 0; r < n; r++) {
        double s = b[r];
        for (long q = 0; q < r; q++) s -= G[r * n + q] * x[q];
        x[r] = s / G[r * n + r];
    }
    for (long r = n - 1; r >= 0; r--) {
        double s = x[r];
        for (long q = r + 1; q < n; q++) s -= G[q * n + r] * x[q];
        x[r] = s / G[r * n + r];
    }
}

static void smooth(mg_t *M, int l, const double *b, double *x) {
    if (M->sm_ksp == 0) cheb_smooth(&M->lev[l], M->sm_pc, M->sm_its, b, x);
    else rich_smooth(&M->lev[l], M->sm_pc, M->sm_its, b, x);
}

static void mcycle(mg_t *M, int l) {
    level_t *g = &M->lev[l];
    if (l == 0) { coarse_solve(M, g->b, g->x); return; }
    level_t *c = &M->lev[l - 1];
    smooth(M, l, g->b, g->x);
    op_residual(g, g->b, g->x, g->r);
    restrict_to(g, c, g->r, c->b);
    par_zero(c->x, c->n);
    const int cycles = (l == 1 || M->cycle == 1) ? 1 : 2;
    for (int q = 0; q < cycles; q++) mcycle(M, l - 1);
    prolong_add(g, c, c->x, g->x);
    smooth(M, l, g->b, g->x);
}

static void pcmg_apply(mg_t *M, const double *r, double *z) {
    level_t *g = &M->lev[M->nlev - 1];
    par_copy(g->b, r, g->n);
    par_zero(g->x, g->n);
    mcycle(M, M->nlev - 1);
    par_copy(z, g->x, g->n);
}

static double lambda_max(const level_t *g) {
    const double PI = 3.14159265358979323846;
    double num = 0, den = 0;
    if (g->ax) { num += g->cx * cos(PI / (g->nx - 1)); den += g->cx; }
    if (g->ay) { num += g->cy * cos(PI / (g->ny - 1)); den += g->cy; }
    if (g->az) { num += g->cz * cos(PI / (g->nz - 1)); den += g->cz; }
    return 1.0 + num / den;
}

static int mg_setup(mg_t *M, const fishcpu_opts *o) {
    memset(M, 0, sizeof *M);
    int m[16][3];
    int nl = 1;
    memcpy(m[0], o->m, sizeof(int) * 3);
    for (;;) {
        if (o->levels > 0 && nl >= o->levels) break;
        int ok = 1;
        for (int d = 0; d < o->dim; d++)
            if (m[nl - 1][d] <= 3 || (m[nl - 1][d] - 1) % 2) ok = 0;
        if (!ok || nl >= 16) break;
        for (int d = 0; d < 3; d++) m[nl][d] = d < o->dim ? (m[nl - 1][d] - 1) / 2 + 1 : 1;
        nl++;
    }
    if (o->levels > 0 && nl < o->levels) return 60;
    M->nlev = nl;
    M->cycle = o->cycle; M->sm_ksp = o->smoother_ksp; M->sm_pc = o->smoother_pc; M->sm_its = o->smooth_its;
    for (int l = 0; l < nl; l++) {
        level_t *g = &M->lev[l];
        level_init(g, o->dim, m[nl - 1 - l], o->L, o->c);
        const double lam = o->smoother_pc == 1 ? 1.0 : lambda_max(g);
        if (o->emax > 0) { g->emin = o->emin; g->emax = o->emax; }
        else { g->emin = 0.1 * lam; g->emax = 1.1 * lam; }
        g->x = par_zeros(g->n); g->b = par_zeros(g->n); g->r = par_zeros(g->n);
        g->t1 = par_zeros(g->n); g->t2 = par_zeros(g->n);
        if (!g->x || !g->b || !g->r || !g->t1 || !g->t2) return 71;
    }
    /* coarsest operator, dense Cholesky  (PCLU on level 0) */
    level_t *g = &M->lev[0];
    const long n = g->n;
    if (n > 2400) return 61;
    M->n0 = n;
    double *A = calloc((size_t)n * n, sizeof(double));
    double *e = calloc(n, sizeof(double)), *col = calloc(n, sizeof(double));
    for (long q = 0; q < n; q++) {
        memset(e, 0, sizeof(double) * n);
        e[q] = 1.0;
        op_apply(g, e, col);
        for (long r = 0; r < n; r++) A[r * n + q] = col[r];
    }
    for (long c = 0; c < n; c++) {
        double d = A[c * n + c];
        for (long q = 0; q < c; q++) d -= A[c * n + q] * A[c * n + q];
        if (!(d > 0)) return 61;
        d = sqrt(d);
        A[c * n + c] = d;
        for (long r = c + 1; r < n; r++) {
            double s = A[r * n + c];
            for (long q = 0; q < c; q++) s -= A[r * n + q] * A[c * n + q];
            A[r * n + c] = s / d;
        }
    }
    M->chol = A;
    free(e); free(col);
    return 0;
}

static void mg_free(mg_t *M) {
    for (int l = 0; l < M->nlev; l++) {
        level_t *g = &M->lev[l];
        free(g->x); free(g->b); free(g->r); free(g->t1); free(g->t2);
    }
    free(M->chol);
}

static void form_function(const level_t *g, const fishcpu_opts *o, const double *u, double *F) {
    const long plane = (long)g->nx * g->ny;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < g->nz; k++)
        for (int j = 0; j < g->ny; j++)
            for (int i = 0; i < g->nx; i++) {
                const long n = k * plane + (long)j * g->nx + i;
                double x, y, z;
                coords(g, i, j, k, &x, &y, &z);
                if (is_bd(g, i, j, k)) {
                    const double s = g->dim == 1 ? o->c[0] * (2.0 / g->hx) : g->diag;
                    F[n] = (u[n] - u_exact(g->dim, o->problem, x, y, z)) * s;
                    continue;
                }
                double uw = 0, ue = 0, us = 0, un = 0, ud = 0, uu = 0, xx, yy, zz;
#define NB(ii, jj, kk, off, out)                                                          \
    if (is_bd(g, ii, jj, kk)) { coords(g, ii, jj, kk, &xx, &yy, &zz); out = u_exact(g->dim, o->problem, xx, yy, zz); } \
    else out = u[n + (off)];
                if (g->ax) { NB(i - 1, j, k, -1, uw) NB(i + 1, j, k, 1, ue) }
                if (g->ay) { NB(i, j - 1, k, -(long)g->nx, us) NB(i, j + 1, k, (long)g->nx, un) }
                if (g->az) { NB(i, j, k - 1, -plane, ud) NB(i, j, k + 1, plane, uu) }
#undef NB
                const double f = f_rhs(g->dim, o->problem, x, y, z, o->c);
                if (g->dim == 1) F[n] = o->c[0] * (2.0 * u[n] - uw - ue) / g->hx - g->hx * f;
                else {
                    double v = g->diag * u[n] - g->cx * (uw + ue);
                    if (g->ay) v -= g->cy * (us + un);
                    if (g->az) v -= g->cz * (uu + ud);
                    F[n] = v - g->vol * f;
                }
            }
}

static double now(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return (double)clock() / CLOCKS_PER_SEC;
#endif
}

/* SNESKSPONLY + KSPCG + PCMG.  u_out, b_out, y_out: optional arrays of n doubles. */
int fishcpu_solve(const fishcpu_opts *o, fishcpu_result *res, double *u_out, double *b_out, double *y_out) {
#ifdef _OPENMP
    if (o->threads > 0) omp_set_num_threads(o->threads);
    res->threads = omp_get_max_threads();
#else
    res->threads = 1;
#endif
    mg_t M;
    int rc = mg_setup(&M, o);
    if (rc) return rc;
    level_t *g = &M.lev[M.nlev - 1];
    const long n = g->n;
    double *u = par_zeros(n), *b = par_zeros(n), *x = par_zeros(n), *r = par_zeros(n), *z = par_zeros(n);
    double *p = par_zeros(n), *w = par_zeros(n);
    if (!u || !b || !x || !r || !z || !p || !w) return 71;
    const long plane = (long)g->nx * g->ny;
    if (o->gonboundary)
#pragma omp parallel for schedule(static)
        for (long q = 0; q < n; q++) {
            const int k = (int)(q / plane), j = (int)((q % plane) / g->nx), i = (int)(q % g->nx);
            if (is_bd(g, i, j, k)) {
                double xx, yy, zz;
                coords(g, i, j, k, &xx, &yy, &zz);
                u[q] = u_exact(g->dim, o->problem, xx, yy, zz);
            }
        }
    form_function(g, o, u, b);
    res->fnorm0 = sqrt(dot(n, b, b));
    res->nlevels = M.nlev;
    /* ---- KSPSolve_CG (timed) ---- */
    const double t0 = now();
    par_copy(r, b, n);
    pcmg_apply(&M, r, z);
    double beta = dot(n, z, r), dp = sqrt(dot(n, z, z)), beta_old = 0.0;
    res->nhist = 0;
    res->hist[res->nhist++] = dp;
    const double ttol = fmax(o->rtol * dp, 1e-50);
    int its = 0;
    while (dp > ttol && its < o->max_it) {
        if (its == 0) par_copy(p, z, n);
        else {
            const double bb = beta / beta_old;
#pragma omp parallel for schedule(static)
            for (long q = 0; q < n; q++) p[q] = z[q] + bb * p[q];
        }
        op_apply(g, p, w);
        const double a = beta / dot(n, p, w);
#pragma omp parallel for schedule(static)
        for (long q = 0; q < n; q++) { x[q] += a * p[q]; r[q] -= a * w[q]; }
        pcmg_apply(&M, r, z);
        beta_old = beta;
        beta = dot(n, z, r);
        dp = sqrt(dot(n, z, z));
        its++;
        if (res->nhist < 256) res->hist[res->nhist++] = dp;
    }
    res->seconds = now() - t0;
    res->its = its;
    /* u = u0 - y ; error norms (fish.c:248-276) */
    if (y_out) memcpy(y_out, x, sizeof(double) * n);
    if (b_out) memcpy(b_out, b, sizeof(double) * n);
    double einf = 0.0, e2 = 0.0;
#pragma omp parallel for schedule(static) reduction(max : einf) reduction(+ : e2)
    for (long q = 0; q < n; q++) {
        u[q] -= x[q];
        const int k = (int)(q / plane), j = (int)((q % plane) / g->nx), i = (int)(q % g->nx);
        double xx, yy, zz;
        coords(g, i, j, k, &xx, &yy, &zz);
        const double e = u[q] - u_exact(g->dim, o->problem, xx, yy, zz);
        if (fabs(e) > einf) einf = fabs(e);
        e2 += e * e;
    }
    double nc = 1.0;
    for (int d = 0; d < o->dim; d++) nc *= (double)(o->m[d] - 1);
    res->errinf = einf;
    res->err2h = sqrt(e2) / sqrt(nc);
    form_function(g, o, u, w);
    res->fnorm1 = sqrt(dot(n, w, w));
    if (u_out) memcpy(u_out, u, sizeof(double) * n);
    free(u); free(b); free(x); free(r); free(z); free(p); free(w);
    mg_free(&M);
    return 0;
}

/* host STREAM-triad figure so the CPU number can be read against its own roofline (HARDWARE.md:14-21) */
double fishcpu_stream_triad_gbs(long n, int reps, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double *a = malloc(sizeof(double) * n), *b = malloc(sizeof(double) * n), *c = malloc(sizeof(double) * n);
    if (!a || !b || !c) return -1.0;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) { a[i] = 1.0; b[i] = 2.0; c[i] = 0.5; }
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        const double t0 = now();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) a[i] = b[i] + 3.0 * c[i];
        const double t = now() - t0;
        if (t < best) best = t;
    }
    const double gbs = 24.0 * n / best / 1e9;
    free(a); free(b); free(c);
    return gbs;
}
