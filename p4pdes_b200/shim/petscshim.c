/* petscshim.c -- the PETSc-shaped host layer behind include/petsc.h (C99, NOT PETSc).
 *
 * Lets the reference's unchanged drivers (c/ch6/fish.c, c/ch7/minimal.c + c/ch6/poissonfunctions.c) run on a B200:
 * objects are thin host structs; every numerical operation of the solve goes through the C ABI of
 * include/p4b200.h into the CUDA kernels.  What runs on the host is what the reference itself runs
 * on the host: option parsing, the user's FormFunctionLocal / FormJacobianLocal callbacks (called
 * exactly as PETSc calls them: once per SNES function evaluation, once per multigrid level for the
 * rediscretised Jacobian, fish.c:7) and the final report.
 *
 * Plugin structure (the PETSc idiom): Mat types and PC types are looked up by name in registries
 * filled through MatRegister()/PCRegister(); the shim registers "stencilcuda" and "mg"/"jacobi"/"none".
 * There is no CPU solver here: -pc_type ilu/sor/... and friends fail with an explanatory error.
 */
#define _POSIX_C_SOURCE 200809L
#include <petsc.h>

#include <stdarg.h>
#include <strings.h>
#include <stdlib.h>
#include <time.h>

#include "p4b200.h"

/* ------------------------------------------------------------------------------------------------ */
/* globals: option database, device context, log                                                    */
/* ------------------------------------------------------------------------------------------------ */
#define MAXOPT 256
static struct { char *name; char *value; int used; } g_opt[MAXOPT];
static int g_nopt = 0;
static char g_prefix[64] = "";
static p4b_ctx *g_ctx = NULL;
static double g_flops = 0.0;
static double g_t_snes = 0.0, g_t_ksp = 0.0, g_t_ksp_dev_ms = 0.0, g_t_jac = 0.0, g_t_func = 0.0;
static int g_initialized = 0;
static int g_newton_solves = 0, g_ts_general = 0;

static double wall(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

PetscErrorCode PetscShimError(MPI_Comm comm, int line, const char *func, const char *file, PetscErrorCode code,
                              const char *msg) {
    (void)comm;
    fprintf(stderr, "[0]PETSC ERROR: %s", msg);
    if (!msg[0] || msg[strlen(msg) - 1] != '\n') fprintf(stderr, "\n");
    fprintf(stderr, "[0]PETSC ERROR: #1 %s() at %s:%d (p4b200 shim)\n", func, file, line);
    return code ? code : 1;
}
#define SHIM_ERR(code, msg) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, code, msg)
#define P4B(call)                                                                                        \
    do {                                                                                                 \
        int rc_ = (call);                                                                                \
        if (rc_) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc_, p4b_last_error()); \
    } while (0)

static PetscErrorCode ensure_ctx(void) {
    if (!g_ctx) P4B(p4b_ctx_create(0, NULL, &g_ctx));
    return 0;
}

/* ---- options ---- */
static int opt_find(const char *name) {
    int found = -1;
    for (int i = 0; i < g_nopt; i++)            /* [PETSc] the options database keeps the LAST value given */
        if (!strcmp(g_opt[i].name, name)) { g_opt[i].used = 1; found = i; }
    return found;
}
static const char *opt_value(const char *name) {
    int i = opt_find(name);
    return i < 0 ? NULL : g_opt[i].value;
}
static int opt_has(const char *name) { return opt_find(name) >= 0; }
static const char *binary_target(const char *optname) {        /* "binary:FILE" -> FILE */
    const char *v = opt_value(optname);
    return (v && !strncmp(v, "binary:", 7) && v[7]) ? v + 7 : NULL;
}

static int opt_bool(const char *name, int dflt) {
    int i = opt_find(name);
    if (i < 0) return dflt;
    const char *v = g_opt[i].value;
    if (!v) return 1;
    return !(v[0] == '0' || v[0] == 'f' || v[0] == 'F' || v[0] == 'n' || v[0] == 'N');
}

static int is_number(const char *s) {
    char *e;
    strtod(s, &e);
    return e != s;
}

PetscErrorCode PetscInitialize(int *argc, char ***argv, const char file[], const char help[]) {
    (void)file;
    g_nopt = 0;
    for (int i = 1; argc && i < *argc; i++) {
        const char *a = (*argv)[i];
        if (a[0] != '-' || is_number(a)) continue;
        if (g_nopt >= MAXOPT) {
            fprintf(stderr, "[p4b200 shim] more than %d options on the command line: %s and what follows are ignored\n", MAXOPT, a);
            break;
        }
        g_opt[g_nopt].name = strdup(a);
        g_opt[g_nopt].value = NULL;
        g_opt[g_nopt].used = 0;
        if (i + 1 < *argc) {
            const char *v = (*argv)[i + 1];
            if (v[0] != '-' || is_number(v)) { g_opt[g_nopt].value = strdup(v); i++; }
        }
        g_nopt++;
    }
    if (opt_has("-help") && help) printf("%s", help);
    g_initialized = 1;
    return 0;
}

PetscErrorCode PetscFinalize(void) {
    if (opt_has("-log_view")) {
        printf("------------------------------------------------------------------ p4b200 shim -log_view\n");
        printf("Time (sec):  SNESSolve %.6e   KSPSolve %.6e (device %.6e)\n", g_t_snes, g_t_ksp, g_t_ksp_dev_ms * 1e-3);
        printf("             FunctionEval(host callback) %.6e   JacobianEval(host callback, all levels) %.6e\n",
               g_t_func, g_t_jac);
        printf("Flop:  %.6e (user PetscLogFlops only)   kernel launches: %lld\n", g_flops, p4b_launch_count());
        if (g_ts_general) printf("TS: callbacks evaluated on the host, matrix-free stage operator (not the library's model)\n");
        if (g_newton_solves)
            printf("SNES newtonls: residual %s\n", p4b_snes2d_last_route()
                   ? "recognised as the library's kernel: evaluated on the device" : "evaluated by the host callback");
    }
    if (opt_has("-options_left")) {
        for (int i = 0; i < g_nopt; i++)
            if (!g_opt[i].used)
                printf("WARNING! option %s %s was set but not used\n", g_opt[i].name, g_opt[i].value ? g_opt[i].value : "");
    }
    if (g_ctx) { p4b_ctx_destroy(g_ctx); g_ctx = NULL; }
    for (int i = 0; i < g_nopt; i++) { free(g_opt[i].name); free(g_opt[i].value); }
    g_nopt = 0;
    return 0;
}

/* [PETSc] PetscVSNPrintf: a plain %g whose output shows neither '.' nor an exponent gets a '.' appended ("t0=0." in
 * c/ch5/output/heat.test1:1, "time 0." in every TS monitor line), so that reals read as reals.  The same is done here: each
 * plain %g is bracketed by two marker bytes in a copy of the format, and the bracketed text is fixed up after vsnprintf. */
static void petsc_vprintf(FILE *f, const char *format, va_list ap) {
    char fmt[1024], buf[4096];
    size_t k = 0;
    for (const char *p = format; *p && k + 6 < sizeof fmt; p++) {
        if (p[0] == '%' && p[1] == '%') { fmt[k++] = *p++; fmt[k++] = *p; continue; }
        if (p[0] == '%' && p[1] == 'g') { fmt[k++] = 1; fmt[k++] = '%'; fmt[k++] = 'g'; fmt[k++] = 2; p++; continue; }
        fmt[k++] = *p;
    }
    fmt[k] = 0;
    vsnprintf(buf, sizeof buf, fmt, ap);
    for (const char *p = buf; *p; p++) {
        if (*p != 1) { fputc(*p, f); continue; }
        int real = 0;
        for (p++; *p && *p != 2; p++) {
            if (*p == '.' || *p == 'e' || *p == 'n' || *p == 'i') real = 1;      /* 1.5, 1e-05, nan, inf */
            fputc(*p, f);
        }
        if (!real) fputc('.', f);
        if (!*p) break;
    }
}
PetscErrorCode PetscPrintf(MPI_Comm comm, const char format[], ...) {
    (void)comm;
    va_list ap;
    va_start(ap, format);
    petsc_vprintf(stdout, format, ap);
    va_end(ap);
    fflush(stdout);
    return 0;
}

PetscErrorCode PetscLogFlops(PetscLogDouble f) { g_flops += f; return 0; }

PetscErrorCode PetscShimOptionsBegin(MPI_Comm comm, const char prefix[], const char title[], const char mansec[]) {
    (void)comm; (void)title; (void)mansec;
    snprintf(g_prefix, sizeof g_prefix, "%s", prefix ? prefix : "");
    return 0;
}
PetscErrorCode PetscShimOptionsEnd(void) { g_prefix[0] = 0; return 0; }

static void full_name(const char *opt, char *out, size_t n) { snprintf(out, n, "-%s%s", g_prefix, opt + 1); }

PetscErrorCode PetscOptionsReal(const char opt[], const char text[], const char man[], PetscReal cur, PetscReal *value,
                                PetscBool *set) {
    (void)text; (void)man;
    char nm[128];
    full_name(opt, nm, sizeof nm);
    const char *v = opt_value(nm);
    if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
    *value = v ? strtod(v, NULL) : cur;
    return 0;
}
PetscErrorCode PetscOptionsInt(const char opt[], const char text[], const char man[], PetscInt cur, PetscInt *value,
                               PetscBool *set) {
    (void)text; (void)man;
    char nm[128];
    full_name(opt, nm, sizeof nm);
    const char *v = opt_value(nm);
    if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
    *value = v ? (PetscInt)strtol(v, NULL, 10) : cur;
    return 0;
}
PetscErrorCode PetscOptionsBool(const char opt[], const char text[], const char man[], PetscBool cur, PetscBool *value,
                                PetscBool *set) {
    (void)text; (void)man;
    char nm[128];
    full_name(opt, nm, sizeof nm);
    int has = opt_has(nm);
    if (set) *set = has ? PETSC_TRUE : PETSC_FALSE;
    *value = has ? (opt_bool(nm, 1) ? PETSC_TRUE : PETSC_FALSE) : cur;
    return 0;
}
PetscErrorCode PetscOptionsEnum(const char opt[], const char text[], const char man[], const char *const *list,
                                PetscEnum cur, PetscEnum *value, PetscBool *set) {
    (void)text; (void)man;
    char nm[128];
    full_name(opt, nm, sizeof nm);
    const char *v = opt_value(nm);
    if (set) *set = v ? PETSC_TRUE : PETSC_FALSE;
    *value = cur;
    if (!v) return 0;
    /* list = names..., "EnumName", "prefix", NULL  (fish.c:109-110) */
    int n = 0;
    while (list[n]) n++;
    n -= 2;
    for (int i = 0; i < n; i++)
        if (!strcasecmp(list[i], v)) { *value = (PetscEnum)i; return 0; }
    {
        char msg[256];
        snprintf(msg, sizeof msg, "Unknown option \"%s\" for %s", v, nm);
        SHIM_ERR(62, msg);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* objects                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
struct _p_DM {
    int dim, M[3], M0[3], dof, sw, refine, setup, refct;   /* M0: the grid before -da_refine (the multigrid base) */
    DMBoundaryType b[3];
    DMDAStencilType st;
    double cmin[3], cmax[3];
    void *appctx;
    DMDASNESFunctionFn *func;
    void *funcctx;
    DMDASNESJacobianFn *jac;
    void *jacctx;
    DMDATSIFunctionLocal ifunc;        /* the TS callback contract (pattern.c:103-114) */
    DMDATSRHSFunctionLocal rhsfunc;
    DMDATSIJacobianLocal ijac;
    DMDATSRHSJacobianLocal rhsjac;
    void *ifuncctx, *rhsfuncctx, *ijacctx, *rhsjacctx;
    Vec pool[8];
    int pool_busy[8];
};

enum { LOC_HOST = 1, LOC_DEV = 2 };
struct _p_Vec {
    size_t n;
    DM dm;
    double *h;          /* host copy */
    double *d;          /* device copy (lazy) */
    int valid;          /* LOC_HOST | LOC_DEV */
    void *tables;       /* pointer tables handed out by DMDAVecGetArray */
    int pooled;
};

/* Mat type "stencilcuda": the values inserted by MatSetValuesStencil are checked against the
 * constant-coefficient Dirichlet stencil of poissonfunctions.h:33-38 and reduced to (diag, cx, cy, cz). */
struct _p_Mat {
    DM dm;
    const char *type;
    double diag, c[3];
    int have_diag, have_c[3];
    long long rows_set;
    int general;        /* 1: the inserted values are NOT such a stencil */
    char why[160];
    int assembled;
    struct rd_check *rd; /* non-NULL: a TS Jacobian of a 2-component periodic DMDA is being checked (TSSolve) */
    /* Mat type "sellcuda": the inserted values are KEPT (any coefficients), row by row, and go to the device as a SELL-32
     * matrix ([PETSc] MATSELL / AIJ) for the SpMV kernel */
    int width;           /* slots per row: 3^dim (every stencil the DMDA can hold) */
    int *ccol, *ccnt;    /* column indices (natural ordering) and fill count per row */
    double *cval;
};

/* TS Jacobians (pattern.c:202-318): the rows inserted by the user's FormIJacobianLocal / FormRHSJacobianLocal are
 * compared, as they arrive, with the stage matrix the device path applies matrix-free,
 *   IJacobian:    shift I - C_c L9  (diag shift + 20 C_c, edges -4 C_c, corners -C_c, component c couples to c only)
 *   RHSJacobian:  the pointwise 2 x 2 block dG/d(u,v) of G = (-u v^2 + phi (1-u), u v^2 - (phi+kappa) v)
 * -- the same idea as the "stencilcuda" check above, for the two-species reaction-diffusion system. */
struct rd_check {
    int mode;           /* 1 IJacobian, 2 RHSJacobian */
    int m;
    double shift, C[2], phi, kappa;
    const double *Y;    /* host state, (u,v) interleaved */
    double maxdev, scale;
    long long rows;
    int bad_structure;
};

struct _p_PC {
    char type[32];
    int levels, cycle, smoother, smooth_its, have_eig;
    double emin, emax, est_lo, est_hi;
    int fuse;
    char levels_pc[32];
};
struct _p_KSP {
    char type[32];
    double rtol, abstol;
    int max_it, its, reason;
    int converged_reason_flag, monitor_flag;
    struct _p_PC pc;
};
struct _p_SNES {
    DM dm;
    char type[32];
    struct _p_KSP ksp;
    Vec sol;
    int monitor, monitor_short, converged_reason_flag, its;
    /* newtonls (minimal.c): tolerances, Jacobian source, grid sequencing, SNESMonitorSet monitors */
    double rtol, atol, stol;
    int max_it, fd_color, mf_operator, grid_sequence, gmres_restart, tablevel, reason, sym_check;
    double sym_tol;          /* -mat_is_symmetric <tol> */
    struct {
        PetscErrorCode (*f)(SNES, PetscInt, PetscReal, void *);
        void *ctx;
        PetscErrorCode (*destroy)(void **);
    } mon[5];
    int nmon;
};

/* ---- registries ---- */
#define MAXREG 16
static struct { char name[32]; PetscErrorCode (*create)(Mat); } g_matreg[MAXREG];
static struct { char name[32]; PetscErrorCode (*create)(PC); } g_pcreg[MAXREG];
static int g_nmatreg = 0, g_npcreg = 0;

PetscErrorCode MatRegister(const char name[], PetscErrorCode (*create)(Mat)) {
    if (g_nmatreg >= MAXREG) SHIM_ERR(62, "Mat registry full");
    snprintf(g_matreg[g_nmatreg].name, 32, "%s", name);
    g_matreg[g_nmatreg++].create = create;
    return 0;
}
PetscErrorCode PCRegister(const char name[], PetscErrorCode (*create)(PC)) {
    if (g_npcreg >= MAXREG) SHIM_ERR(62, "PC registry full");
    snprintf(g_pcreg[g_npcreg].name, 32, "%s", name);
    g_pcreg[g_npcreg++].create = create;
    return 0;
}
static PetscErrorCode MatCreate_StencilCUDA(Mat A) { A->type = MATSTENCILCUDA; return 0; }
static PetscErrorCode MatCreate_SellCUDA(Mat A) {
    const size_t nrows = (size_t)A->dm->M[0] * A->dm->M[1] * A->dm->M[2];
    if (A->dm->dof != 1) SHIM_ERR(56, "Mat type sellcuda: one degree of freedom per node");
    A->type = MATSELLCUDA;
    A->width = A->dm->dim == 1 ? 3 : (A->dm->dim == 2 ? 9 : 27);
    A->ccol = (int *)malloc(sizeof(int) * nrows * A->width);
    A->cval = (double *)malloc(sizeof(double) * nrows * A->width);
    A->ccnt = (int *)calloc(nrows, sizeof(int));
    if (!A->ccol || !A->cval || !A->ccnt) SHIM_ERR(55, "out of host memory for the assembled matrix");
    return 0;
}
static PetscErrorCode PCCreate_MG_P4B(PC pc) { snprintf(pc->type, 32, "%s", PCMG); return 0; }
static PetscErrorCode PCCreate_Jacobi_P4B(PC pc) { snprintf(pc->type, 32, "%s", PCJACOBI); return 0; }
static PetscErrorCode PCCreate_None_P4B(PC pc) { snprintf(pc->type, 32, "%s", PCNONE); return 0; }
static void register_all(void) {
    static int done = 0;
    if (done) return;
    done = 1;
    MatRegister(MATSTENCILCUDA, MatCreate_StencilCUDA);
    MatRegister(MATSELLCUDA, MatCreate_SellCUDA);
    PCRegister(PCMG, PCCreate_MG_P4B);
    PCRegister(PCJACOBI, PCCreate_Jacobi_P4B);
    PCRegister(PCNONE, PCCreate_None_P4B);
}

PetscErrorCode PCSetType(PC pc, PCType type) {
    register_all();
    for (int i = 0; i < g_npcreg; i++)
        if (!strcmp(g_pcreg[i].name, type)) return g_pcreg[i].create(pc);
    {
        char msg[256];
        snprintf(msg, sizeof msg,
                 "PC type %s is not provided by the p4b200 device path (registered: mg, jacobi, none); "
                 "PETSc's ilu/icc/sor/asm/bjacobi are sequential CPU algorithms", type);
        SHIM_ERR(86, msg);
    }
}

/* ---- DMDA ---- */
static PetscErrorCode da_create(int dim, const DMBoundaryType *b, DMDAStencilType st, const PetscInt *M, PetscInt dof,
                                PetscInt s, DM *da) {
    if (!g_initialized) SHIM_ERR(73, "PetscInitialize() must be called first");
    DM d = (DM)calloc(1, sizeof *d);
    d->dim = dim;
    for (int i = 0; i < 3; i++) {
        d->M[i] = i < dim ? M[i] : 1;
        d->b[i] = i < dim ? b[i] : DM_BOUNDARY_NONE;
        d->cmin[i] = 0.0;
        d->cmax[i] = 1.0;
        if (d->b[i] != DM_BOUNDARY_NONE && d->b[i] != DM_BOUNDARY_PERIODIC) {
            free(d);
            SHIM_ERR(56, "DM_BOUNDARY_NONE and DM_BOUNDARY_PERIODIC grids are provided by the shim");
        }
    }
    if (dim != 2 && (d->b[0] == DM_BOUNDARY_PERIODIC || d->b[1] == DM_BOUNDARY_PERIODIC || d->b[2] == DM_BOUNDARY_PERIODIC)) {
        free(d);
        SHIM_ERR(56, "periodic DMDAs are provided in 2-D (pattern.c:79-84, heat.c:60-65)");
    }
    if (dof < 1 || (dof > 1 && !(d->b[0] == DM_BOUNDARY_PERIODIC && d->b[1] == DM_BOUNDARY_PERIODIC))) {
        free(d);
        SHIM_ERR(56, "dof > 1 is provided on the periodic 2-D DMDA only");
    }
    d->dof = dof; d->sw = s; d->st = st; d->refct = 1;
    *da = d;
    return 0;
}
PetscErrorCode DMDACreate1d(MPI_Comm comm, DMBoundaryType bx, PetscInt M, PetscInt dof, PetscInt s, const PetscInt lx[],
                            DM *da) {
    (void)comm; (void)lx;
    DMBoundaryType b[3] = {bx, DM_BOUNDARY_NONE, DM_BOUNDARY_NONE};
    PetscInt m[3] = {M, 1, 1};
    return da_create(1, b, DMDA_STENCIL_STAR, m, dof, s, da);
}
PetscErrorCode DMDACreate2d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMDAStencilType st, PetscInt M, PetscInt N,
                            PetscInt m, PetscInt n, PetscInt dof, PetscInt s, const PetscInt lx[], const PetscInt ly[],
                            DM *da) {
    (void)comm; (void)m; (void)n; (void)lx; (void)ly;
    DMBoundaryType b[3] = {bx, by, DM_BOUNDARY_NONE};
    PetscInt mm[3] = {M, N, 1};
    return da_create(2, b, st, mm, dof, s, da);
}
PetscErrorCode DMDACreate3d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMBoundaryType bz, DMDAStencilType st,
                            PetscInt M, PetscInt N, PetscInt P, PetscInt m, PetscInt n, PetscInt p, PetscInt dof,
                            PetscInt s, const PetscInt lx[], const PetscInt ly[], const PetscInt lz[], DM *da) {
    (void)comm; (void)m; (void)n; (void)p; (void)lx; (void)ly; (void)lz;
    DMBoundaryType b[3] = {bx, by, bz};
    PetscInt mm[3] = {M, N, P};
    return da_create(3, b, st, mm, dof, s, da);
}
PetscErrorCode DMSetApplicationContext(DM dm, void *ctx) { dm->appctx = ctx; return 0; }
PetscErrorCode DMGetApplicationContext(DM dm, void *ctx) { *(void **)ctx = dm->appctx; return 0; }
PetscErrorCode DMSetFromOptions(DM dm) {
    const char *v;
    const char *names[3] = {"-da_grid_x", "-da_grid_y", "-da_grid_z"};
    for (int i = 0; i < dm->dim; i++)
        if ((v = opt_value(names[i]))) dm->M[i] = atoi(v);
    if ((v = opt_value("-da_refine"))) dm->refine = atoi(v);
    return 0;
}
PetscErrorCode DMSetUp(DM dm) {
    if (dm->setup) return 0;
    memcpy(dm->M0, dm->M, sizeof dm->M0);
    for (int i = 0; i < dm->dim; i++)       /* -da_refine n: M <- 1 + 2^n (M-1), periodic M <- 2^n M   */
        dm->M[i] = dm->b[i] == DM_BOUNDARY_PERIODIC ? (dm->M[i] << dm->refine) : 1 + (1 << dm->refine) * (dm->M[i] - 1);
    dm->setup = 1;
    return 0;
}
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax,
                                         PetscReal zmin, PetscReal zmax) {
    da->cmin[0] = xmin; da->cmax[0] = xmax;
    da->cmin[1] = ymin; da->cmax[1] = ymax;
    da->cmin[2] = zmin; da->cmax[2] = zmax;
    return 0;
}
PetscErrorCode DMGetBoundingBox(DM dm, PetscReal gmin[], PetscReal gmax[]) {
    for (int i = 0; i < dm->dim; i++) {
        if (gmin) gmin[i] = dm->cmin[i];
        if (gmax) gmax[i] = dm->cmax[i];
    }
    return 0;
}
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info) {
    if (!da->setup) SHIM_ERR(73, "DMSetUp() must be called before DMDAGetLocalInfo()");
    memset(info, 0, sizeof *info);
    info->da = da; info->dim = da->dim; info->dof = da->dof; info->sw = da->sw;
    info->mx = da->M[0]; info->my = da->M[1]; info->mz = da->M[2];
    info->xs = info->ys = info->zs = 0;
    info->xm = da->M[0]; info->ym = da->M[1]; info->zm = da->M[2];
    info->gxs = info->gys = info->gzs = 0;        /* one logical rank owns the whole grid */
    info->gxm = da->M[0]; info->gym = da->M[1]; info->gzm = da->M[2];
    if (da->b[0] == DM_BOUNDARY_PERIODIC) { info->gxs = -da->sw; info->gxm += 2 * da->sw; }   /* periodic ghosts */
    if (da->b[1] == DM_BOUNDARY_PERIODIC) { info->gys = -da->sw; info->gym += 2 * da->sw; }
    info->bx = da->b[0]; info->by = da->b[1]; info->bz = da->b[2];
    info->st = da->st;
    return 0;
}
static size_t da_n(DM dm) { return (size_t)dm->M[0] * dm->M[1] * dm->M[2] * (size_t)dm->dof; }

static PetscErrorCode vec_new(DM dm, Vec *v) {
    Vec x = (Vec)calloc(1, sizeof *x);
    if (!x) SHIM_ERR(55, "out of host memory for a Vec");
    x->n = da_n(dm);
    x->dm = dm;
    x->h = (double *)calloc(x->n, sizeof(double));
    if (!x->h) { free(x); SHIM_ERR(55, "out of host memory for a Vec"); }
    x->valid = LOC_HOST;
    *v = x;
    return 0;
}
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *g) {
    if (!dm->setup) SHIM_ERR(73, "DMSetUp() must be called before DMCreateGlobalVector()");
    return vec_new(dm, g);
}
PetscErrorCode DMGetGlobalVector(DM dm, Vec *g) {
    for (int i = 0; i < 8; i++)
        if (dm->pool[i] && !dm->pool_busy[i]) { dm->pool_busy[i] = 1; *g = dm->pool[i]; return 0; }
    for (int i = 0; i < 8; i++)
        if (!dm->pool[i]) {
            PetscCall(DMCreateGlobalVector(dm, &dm->pool[i]));
            dm->pool[i]->pooled = 1;
            dm->pool_busy[i] = 1;
            *g = dm->pool[i];
            return 0;
        }
    SHIM_ERR(77, "DMGetGlobalVector pool exhausted");
}
PetscErrorCode DMRestoreGlobalVector(DM dm, Vec *g) {
    for (int i = 0; i < 8; i++)
        if (dm->pool[i] == *g) { dm->pool_busy[i] = 0; *g = NULL; return 0; }
    SHIM_ERR(62, "vector was not obtained with DMGetGlobalVector");
}
/* one logical rank on a non-periodic DMDA: the ghosted local vector IS the global one */
PetscErrorCode DMGetLocalVector(DM dm, Vec *l) { return vec_new(dm, l); }
static void vec_free(Vec v);
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *l) {
    (void)dm;
    if (l && *l) { vec_free(*l); *l = NULL; }
    return 0;
}
static PetscErrorCode vec_to_host(Vec v);
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l) {
    (void)dm;
    if (mode != INSERT_VALUES) SHIM_ERR(56, "DMGlobalToLocal: INSERT_VALUES only");
    if (g->n != l->n) SHIM_ERR(75, "DMGlobalToLocal: incompatible vector sizes");
    PetscCall(vec_to_host(g));
    memcpy(l->h, g->h, g->n * sizeof(double));
    l->valid = LOC_HOST;
    return 0;
}
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l) { (void)dm; (void)g; (void)mode; (void)l; return 0; }

static void vec_free(Vec v) {
    if (!v) return;
    if (v->d && g_ctx) p4b_free(g_ctx, v->d);
    free(v->h);
    free(v->tables);
    free(v);
}
PetscErrorCode VecDestroy(Vec *v) {
    if (!v || !*v) return 0;
    if ((*v)->pooled) SHIM_ERR(62, "cannot VecDestroy a vector owned by DMGetGlobalVector");
    vec_free(*v);
    *v = NULL;
    return 0;
}
PetscErrorCode DMDestroy(DM *dm) {
    if (!dm || !*dm) return 0;
    if (--(*dm)->refct > 0) { *dm = NULL; return 0; }     /* SNES keeps its own reference (fish.c:245-249) */
    for (int i = 0; i < 8; i++) vec_free((*dm)->pool[i]);
    free(*dm);
    *dm = NULL;
    return 0;
}

static PetscErrorCode vec_to_host(Vec v) {
    if (v->valid & LOC_HOST) return 0;
    P4B(p4b_memcpy_d2h(g_ctx, v->h, v->d, v->n * sizeof(double)));
    v->valid |= LOC_HOST;
    return 0;
}
static PetscErrorCode vec_to_dev(Vec v) {
    PetscCall(ensure_ctx());
    if (!v->d) P4B(p4b_malloc(g_ctx, v->n * sizeof(double), (void **)&v->d));
    if (v->valid & LOC_DEV) return 0;
    P4B(p4b_memcpy_h2d(g_ctx, v->d, v->h, v->n * sizeof(double)));
    v->valid |= LOC_DEV;
    return 0;
}

/* a[k][j][i] views with global indices (one rank owns everything, so no offsets are needed) */
static void *make_tables_dof(int dim, const int *M, int dof, double *base) {
    if (dim == 2) {         /* a[j][i] with i counting nodes of dof doubles each (pattern.c: Field **aY) */
        double **rows = (double **)malloc(sizeof(double *) * (size_t)M[1]);
        for (int j = 0; j < M[1]; j++) rows[j] = base + (size_t)j * M[0] * dof;
        return rows;
    }
    return NULL;
}
static void *make_tables(int dim, const int *M, double *base) {
    if (dim == 1) return NULL;
    if (dim == 2) return make_tables_dof(dim, M, 1, base);
    size_t np = (size_t)M[2], nr = (size_t)M[2] * M[1];
    char *blk = (char *)malloc(sizeof(double **) * np + sizeof(double *) * nr);
    double ***planes = (double ***)blk;
    double **rows = (double **)(blk + sizeof(double **) * np);
    for (size_t k = 0; k < np; k++) {
        planes[k] = rows + k * M[1];
        for (int j = 0; j < M[1]; j++) planes[k][j] = base + (k * M[1] + j) * (size_t)M[0];
    }
    return planes;
}
static PetscErrorCode get_array(DM da, Vec vec, void *array, int write) {
    PetscCall(vec_to_host(vec));
    if (write) vec->valid = LOC_HOST;
    free(vec->tables);
    vec->tables = da->dof > 1 ? make_tables_dof(da->dim, da->M, da->dof, vec->h) : make_tables(da->dim, da->M, vec->h);
    *(void **)array = da->dim == 1 ? (void *)vec->h : vec->tables;
    return 0;
}
PetscErrorCode DMDAVecGetArray(DM da, Vec vec, void *array) { return get_array(da, vec, array, 1); }
PetscErrorCode DMDAVecGetArrayRead(DM da, Vec vec, void *array) { return get_array(da, vec, array, 0); }
PetscErrorCode DMDAVecRestoreArray(DM da, Vec vec, void *array) {
    (void)da;
    free(vec->tables);
    vec->tables = NULL;
    *(void **)array = NULL;
    return 0;
}
PetscErrorCode DMDAVecRestoreArrayRead(DM da, Vec vec, void *array) { return DMDAVecRestoreArray(da, vec, array); }

PetscErrorCode DMDASNESSetFunctionLocal(DM dm, InsertMode imode, DMDASNESFunctionFn *func, void *ctx) {
    (void)imode;
    dm->func = func; dm->funcctx = ctx;
    return 0;
}
PetscErrorCode DMDASNESSetJacobianLocal(DM dm, DMDASNESJacobianFn *func, void *ctx) {
    dm->jac = func; dm->jacctx = ctx;
    return 0;
}

/* ---- Vec operations: BLAS-1 on the device (p4b_vec_*), trivial fills on the host ---- */
PetscErrorCode VecSet(Vec x, PetscScalar a) {
    for (size_t i = 0; i < x->n; i++) x->h[i] = a;
    x->valid = LOC_HOST;
    return 0;
}
/* [PETSc] PetscRandom, default type rander48: the drand48 recurrence seeded 0x12345678 (+ 76543 * rank); restated from
 * memory of PETSc's rander48.c, UNPINNED (no PETSc here, no golden uses a random vector): include/p4b200.h */
struct _p_PetscRandom { unsigned long long state; };
static struct _p_PetscRandom g_default_random;
static int g_default_random_set = 0;
PetscErrorCode PetscRandomCreate(MPI_Comm comm, PetscRandom *r) {
    (void)comm;
    *r = (PetscRandom)malloc(sizeof(struct _p_PetscRandom));
    if (!*r) SHIM_ERR(55, "out of memory");
    (*r)->state = p4b_rander48_seed(0x12345678UL);
    return 0;
}
PetscErrorCode PetscRandomDestroy(PetscRandom *r) {
    if (r && *r) { free(*r); *r = NULL; }
    return 0;
}
PetscErrorCode VecSetRandom(Vec x, PetscRandom r) {
    if (!r) {       /* [PETSc] VecSetRandom(x, NULL): one library-owned generator, its stream continues across calls */
        if (!g_default_random_set) { g_default_random.state = p4b_rander48_seed(0x12345678UL); g_default_random_set = 1; }
        r = &g_default_random;
    }
    P4B(p4b_rander48_fill(&r->state, x->n, x->h));
    x->valid = LOC_HOST;
    return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) {
    if (x->n != y->n) SHIM_ERR(75, "VecAXPY: incompatible vector sizes");
    PetscCall(vec_to_dev(x));
    PetscCall(vec_to_dev(y));
    P4B(p4b_vec_axpy(g_ctx, y->n, a, x->d, y->d));
    y->valid = LOC_DEV;
    return 0;
}
PetscErrorCode VecAYPX(Vec y, PetscScalar b, Vec x) {
    if (x->n != y->n) SHIM_ERR(75, "VecAYPX: incompatible vector sizes");
    PetscCall(vec_to_dev(x));
    PetscCall(vec_to_dev(y));
    P4B(p4b_vec_aypx(g_ctx, y->n, b, x->d, y->d));
    y->valid = LOC_DEV;
    return 0;
}
PetscErrorCode VecScale(Vec x, PetscScalar a) {
    PetscCall(vec_to_dev(x));
    P4B(p4b_vec_aypx(g_ctx, x->n, a - 1.0, x->d, x->d));    /* x = x + (a-1) x */
    x->valid = LOC_DEV;
    return 0;
}
PetscErrorCode VecCopy(Vec x, Vec y) {
    PetscCall(vec_to_host(x));
    memcpy(y->h, x->h, x->n * sizeof(double));
    y->valid = LOC_HOST;
    return 0;
}
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val) {
    PetscCall(vec_to_dev(x));
    PetscCall(vec_to_dev(y));
    P4B(p4b_vec_dot(g_ctx, x->n, x->d, y->d, val));
    return 0;
}
PetscErrorCode VecNorm(Vec x, NormType type, PetscReal *val) {
    PetscCall(vec_to_dev(x));
    if (type == NORM_2 || type == NORM_FROBENIUS) P4B(p4b_vec_norm2(g_ctx, x->n, x->d, val));
    else if (type == NORM_INFINITY) P4B(p4b_vec_norminf(g_ctx, x->n, x->d, val));
    else SHIM_ERR(56, "VecNorm: only NORM_2 and NORM_INFINITY are provided");
    return 0;
}
PetscErrorCode VecGetSize(Vec x, PetscInt *size) { *size = (PetscInt)x->n; return 0; }
PetscErrorCode VecDuplicate(Vec v, Vec *newv) { return vec_new(v->dm, newv); }

/* ---- Mat "stencilcuda" ---- */
static PetscErrorCode mat_new(DM dm, Mat *mat) {
    register_all();
    Mat A = (Mat)calloc(1, sizeof *A);
    A->dm = dm;
    const char *want = opt_value("-mat_type");
    if (!want || !strcmp(want, "aij")) want = MATSTENCILCUDA;      /* DMCreateMatrix default -> our structured type */
    for (int i = 0; i < g_nmatreg; i++)
        if (!strcmp(g_matreg[i].name, want)) {
            PetscCall(g_matreg[i].create(A));
            *mat = A;
            return 0;
        }
    free(A);
    SHIM_ERR(86, "unknown -mat_type (registered: stencilcuda, sellcuda)");
}
PetscErrorCode DMCreateMatrix(DM dm, Mat *mat) { return mat_new(dm, mat); }
PetscErrorCode MatDestroy(Mat *mat) {
    if (mat && *mat) { free((*mat)->ccol); free((*mat)->cval); free((*mat)->ccnt); free(*mat); *mat = NULL; }
    return 0;
}
PetscErrorCode MatZeroEntries(Mat A) {
    if (A->ccnt) {
        const size_t nrows = (size_t)A->dm->M[0] * A->dm->M[1] * A->dm->M[2];
        for (size_t r = 0; r < nrows * (size_t)A->width; r++) A->cval[r] = 0.0;      /* keeps the nonzero pattern, as PETSc does */
    }
    A->have_diag = A->have_c[0] = A->have_c[1] = A->have_c[2] = 0;
    A->rows_set = 0; A->general = 0; A->assembled = 0;
    return 0;
}
static PetscErrorCode rd_check_rows(struct rd_check *k, PetscInt m, const MatStencil idxm[], PetscInt n,
                                    const MatStencil idxn[], const PetscScalar v[]) {
    for (int r = 0; r < m; r++) {
        const int ri = idxm[r].i, rj = idxm[r].j, rc = idxm[r].c;
        if (rc < 0 || rc > 1 || ri < 0 || ri >= k->m || rj < 0 || rj >= k->m) { k->bad_structure = 1; continue; }
        if (k->mode == 1) {
            if (n != 9) k->bad_structure = 1;
            for (int c = 0; c < n; c++) {
                const int di = idxn[c].i - ri, dj = idxn[c].j - rj;       /* ghost indices -1, m are legal (periodic) */
                if (idxn[c].c != rc || di < -1 || di > 1 || dj < -1 || dj > 1) { k->bad_structure = 1; continue; }
                const double want = (!di && !dj) ? k->shift + 20.0 * k->C[rc] : ((!di || !dj) ? -4.0 * k->C[rc] : -k->C[rc]);
                const double dev = fabs(v[r * n + c] - want);
                if (dev > k->maxdev) k->maxdev = dev;
            }
        } else {
            const double u = k->Y[2 * ((size_t)rj * k->m + ri)], w = k->Y[2 * ((size_t)rj * k->m + ri) + 1];
            if (n != 2) k->bad_structure = 1;
            for (int c = 0; c < n; c++) {
                if (idxn[c].i != ri || idxn[c].j != rj || idxn[c].c < 0 || idxn[c].c > 1) { k->bad_structure = 1; continue; }
                const int cc = idxn[c].c;
                const double want = rc == 0 ? (cc == 0 ? -w * w - k->phi : -2.0 * u * w)
                                            : (cc == 0 ? w * w : 2.0 * u * w - (k->phi + k->kappa));
                const double dev = fabs(v[r * n + c] - want);
                if (dev > k->maxdev) k->maxdev = dev;
            }
        }
        k->rows++;
    }
    return 0;
}
/* Mat type "sellcuda": INSERT_VALUES of one or more rows */
static PetscErrorCode sell_set_rows(Mat A, PetscInt m, const MatStencil idxm[], PetscInt n, const MatStencil idxn[],
                                    const PetscScalar v[]) {
    const DM dm = A->dm;
    const int dim = dm->dim, W = A->width;
    for (int r = 0; r < m; r++) {
        const int ri = idxm[r].i, rj = dim >= 2 ? idxm[r].j : 0, rk = dim >= 3 ? idxm[r].k : 0;
        if (ri < 0 || ri >= dm->M[0] || rj < 0 || rj >= dm->M[1] || rk < 0 || rk >= dm->M[2])
            SHIM_ERR(63, "MatSetValuesStencil: row outside the grid");
        const size_t row = ((size_t)rk * dm->M[1] + rj) * dm->M[0] + ri;
        for (int c = 0; c < n; c++) {
            const int ci = idxn[c].i, cj = dim >= 2 ? idxn[c].j : 0, ck = dim >= 3 ? idxn[c].k : 0;
            if (ci < 0 || ci >= dm->M[0] || cj < 0 || cj >= dm->M[1] || ck < 0 || ck >= dm->M[2])
                SHIM_ERR(63, "MatSetValuesStencil: column outside the (non-periodic) grid");
            const int col = (int)(((size_t)ck * dm->M[1] + cj) * dm->M[0] + ci);
            int slot = -1;
            for (int k = 0; k < A->ccnt[row]; k++)
                if (A->ccol[row * W + k] == col) { slot = k; break; }
            if (slot < 0) {
                if (A->ccnt[row] >= W) SHIM_ERR(63, "MatSetValuesStencil: more entries in a row than the DMDA stencil holds");
                slot = A->ccnt[row]++;
                A->ccol[row * W + slot] = col;
            }
            A->cval[row * W + slot] = v[r * n + c];
        }
        A->rows_set++;
    }
    return 0;
}
static void mat_reject(Mat A, const char *why) {
    if (!A->general) { A->general = 1; snprintf(A->why, sizeof A->why, "%s", why); }
}
static int node_is_bdry(const DM dm, int i, int j, int k) {
    if (i == 0 || i == dm->M[0] - 1) return 1;
    if (dm->dim >= 2 && (j == 0 || j == dm->M[1] - 1)) return 1;
    if (dm->dim >= 3 && (k == 0 || k == dm->M[2] - 1)) return 1;
    return 0;
}
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil idxm[], PetscInt n, const MatStencil idxn[],
                                   const PetscScalar v[], InsertMode addv) {
    if (addv != INSERT_VALUES) { mat_reject(A, "ADD_VALUES insertion"); return 0; }
    const DM dm = A->dm;
    const int dim = dm->dim;
    if (A->rd) return rd_check_rows(A->rd, m, idxm, n, idxn, v);
    if (A->ccnt) return sell_set_rows(A, m, idxm, n, idxn, v);
    for (int r = 0; r < m; r++) {
        const int ri = idxm[r].i, rj = dim >= 2 ? idxm[r].j : 0, rk = dim >= 3 ? idxm[r].k : 0;
        const int rb = node_is_bdry(dm, ri, rj, rk);
        int nb_seen = 0, have_d = 0;
        for (int c = 0; c < n; c++) {
            const int ci = idxn[c].i, cj = dim >= 2 ? idxn[c].j : 0, ck = dim >= 3 ? idxn[c].k : 0;
            const double val = v[r * n + c];
            const int di = ci - ri, dj = cj - rj, dk = ck - rk;
            if (!di && !dj && !dk) {
                if (!A->have_diag) { A->diag = val; A->have_diag = 1; }
                else if (val != A->diag) mat_reject(A, "diagonal is not constant");
                have_d = 1;
                continue;
            }
            const int dir = (di && !dj && !dk && (di == 1 || di == -1)) ? 0
                          : (!di && dj && !dk && (dj == 1 || dj == -1)) ? 1
                          : (!di && !dj && dk && (dk == 1 || dk == -1)) ? 2 : -1;
            if (dir < 0) { mat_reject(A, "entry outside the 3/5/7-point star"); continue; }
            if (rb) { mat_reject(A, "off-diagonal entry in a boundary row"); continue; }
            if (node_is_bdry(dm, ci, cj, ck)) { mat_reject(A, "column to a boundary node was not dropped"); continue; }
            if (!A->have_c[dir]) { A->c[dir] = -val; A->have_c[dir] = 1; }
            else if (-val != A->c[dir]) mat_reject(A, "off-diagonal coefficient is not constant");
            nb_seen++;
        }
        if (!have_d) mat_reject(A, "row without a diagonal entry");
        if (!rb) {      /* an interior row must couple to every interior neighbour */
            int expect = 0;
            if (ri - 1 > 0) expect++;
            if (ri + 1 < dm->M[0] - 1) expect++;
            if (dim >= 2) { if (rj - 1 > 0) expect++; if (rj + 1 < dm->M[1] - 1) expect++; }
            if (dim >= 3) { if (rk - 1 > 0) expect++; if (rk + 1 < dm->M[2] - 1) expect++; }
            if (nb_seen != expect) mat_reject(A, "interior row does not couple to all interior neighbours");
        }
        A->rows_set++;
    }
    return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) {
    if (t == MAT_FINAL_ASSEMBLY) A->assembled = 1;
    return 0;
}
static PetscErrorCode mat_to_grid(Mat A, p4b_grid *g) {
    memset(g, 0, sizeof *g);
    g->dim = A->dm->dim;
    g->mx = A->dm->M[0]; g->my = A->dm->M[1]; g->mz = A->dm->M[2];
    g->Lx = A->dm->cmax[0] - A->dm->cmin[0];
    g->Ly = A->dm->cmax[1] - A->dm->cmin[1];
    g->Lz = A->dm->cmax[2] - A->dm->cmin[2];
    g->cx = g->cy = g->cz = 1.0;        /* the real coefficients travel separately (p4b_mg_create_stencil) */
    return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) {
    /* matrix-free apply of the recognised stencil: y = A x */
    if (A->general) SHIM_ERR(56, "MatMult: matrix is not a constant-coefficient stencil");
    p4b_grid g;
    p4b_mg_opts o;
    p4b_mg *mg = NULL;
    PetscCall(mat_to_grid(A, &g));
    p4b_mg_default_opts(&o);
    o.levels = 1;
    double coef[4] = {A->diag, A->c[0], A->c[1], A->c[2]};
    PetscCall(vec_to_dev(x));
    PetscCall(vec_to_dev(y));
    /* a one-level hierarchy gives access to the stencil kernel with explicit coefficients:
     * residual with b = 0 and a sign flip would cost an extra pass, so use r = 0 - A x then scale */
    P4B(p4b_mg_create_stencil(g_ctx, &g, &o, coef, 1, &mg));
    P4B(p4b_mg_matmult(mg, x->d, y->d));
    P4B(p4b_mg_destroy(mg));
    y->valid = LOC_DEV;
    return 0;
}

/* ---- SNES / KSP ---- */
PetscErrorCode SNESCreate(MPI_Comm comm, SNES *snes) {
    (void)comm;
    register_all();
    SNES s = (SNES)calloc(1, sizeof *s);
    snprintf(s->type, 32, "%s", SNESNEWTONLS);
    snprintf(s->ksp.type, 32, "%s", KSPGMRES);       /* PETSc defaults; fish.c overrides both (:231-233) */
    s->ksp.rtol = 1e-5; s->ksp.abstol = 1e-50; s->ksp.max_it = 10000;
    s->ksp.pc.type[0] = 0;
    s->ksp.pc.levels = 0; s->ksp.pc.cycle = P4B_CYCLE_V; s->ksp.pc.smoother = P4B_SMOOTH_CHEBYSHEV;
    s->ksp.pc.smooth_its = 2; s->ksp.pc.est_lo = 0.1; s->ksp.pc.est_hi = 1.1; s->ksp.pc.fuse = 1;
    snprintf(s->ksp.pc.levels_pc, 32, "sor");        /* PETSc's default level smoother PC */
    s->rtol = 1e-8; s->atol = 1e-50; s->stol = 1e-8; s->max_it = 50; s->gmres_restart = 30;   /* [PETSc] SNES defaults */
    *snes = s;
    return 0;
}
PetscErrorCode SNESSetDM(SNES snes, DM dm) { snes->dm = dm; dm->refct++; return 0; }
PetscErrorCode SNESGetDM(SNES snes, DM *dm) { *dm = snes->dm; return 0; }
PetscErrorCode SNESSetType(SNES snes, SNESType type) { snprintf(snes->type, 32, "%s", type); return 0; }
PetscErrorCode SNESGetKSP(SNES snes, KSP *ksp) { *ksp = &snes->ksp; return 0; }
PetscErrorCode KSPSetType(KSP ksp, KSPType type) { snprintf(ksp->type, 32, "%s", type); return 0; }
PetscErrorCode KSPGetPC(KSP ksp, PC *pc) { *pc = &ksp->pc; return 0; }
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits) {
    (void)dtol;
    if (rtol != PETSC_DEFAULT) ksp->rtol = rtol;
    if (abstol != PETSC_DEFAULT) ksp->abstol = abstol;
    if (maxits != PETSC_DEFAULT) ksp->max_it = maxits;
    return 0;
}
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its) { *its = ksp->its; return 0; }
PetscErrorCode SNESGetIterationNumber(SNES snes, PetscInt *it) { *it = snes->its; return 0; }
PetscErrorCode SNESGetSolution(SNES snes, Vec *x) { *x = snes->sol; return 0; }

PetscErrorCode SNESSetFromOptions(SNES snes) {
    const char *v;
    KSP ksp = &snes->ksp;
    PC pc = &ksp->pc;
    if ((v = opt_value("-snes_type"))) snprintf(snes->type, 32, "%s", v);
    if ((v = opt_value("-ksp_type"))) snprintf(ksp->type, 32, "%s", v);
    if ((v = opt_value("-ksp_rtol"))) ksp->rtol = strtod(v, NULL);
    if ((v = opt_value("-ksp_atol"))) ksp->abstol = strtod(v, NULL);
    if ((v = opt_value("-ksp_max_it"))) ksp->max_it = atoi(v);
    if ((v = opt_value("-pc_type"))) PetscCall(PCSetType(pc, v));
    if ((v = opt_value("-pc_mg_levels"))) pc->levels = atoi(v);
    if ((v = opt_value("-pc_mg_cycle_type"))) pc->cycle = (v[0] == 'w' || v[0] == 'W') ? P4B_CYCLE_W : P4B_CYCLE_V;
    if ((v = opt_value("-mg_levels_ksp_type"))) {
        if (!strcmp(v, "chebyshev")) pc->smoother = P4B_SMOOTH_CHEBYSHEV;
        else if (!strcmp(v, "richardson")) pc->smoother = P4B_SMOOTH_RICHARDSON;
        else SHIM_ERR(86, "-mg_levels_ksp_type must be chebyshev or richardson on the device path");
    }
    if ((v = opt_value("-mg_levels_ksp_max_it"))) pc->smooth_its = atoi(v);
    if ((v = opt_value("-mg_levels_pc_type"))) snprintf(pc->levels_pc, 32, "%s", v);
    if ((v = opt_value("-mg_levels_ksp_chebyshev_eigenvalues"))) {
        if (sscanf(v, "%lf,%lf", &pc->emin, &pc->emax) != 2) SHIM_ERR(62, "need emin,emax");
        pc->have_eig = 1;
    }
    if ((v = opt_value("-mg_levels_ksp_chebyshev_esteig"))) {
        double a, b, c, d;
        if (sscanf(v, "%lf,%lf,%lf,%lf", &a, &b, &c, &d) == 4) { pc->est_lo = b; pc->est_hi = d; }
    }
    if (opt_has("-p4b_no_fuse")) pc->fuse = 0;
    ksp->converged_reason_flag = opt_has("-ksp_converged_reason");
    ksp->monitor_flag = opt_has("-ksp_monitor");
    snes->monitor_short = opt_has("-snes_monitor_short");
    snes->monitor = opt_has("-snes_monitor");
    snes->converged_reason_flag = opt_has("-snes_converged_reason");
    if ((v = opt_value("-snes_rtol"))) snes->rtol = strtod(v, NULL);
    if ((v = opt_value("-snes_atol"))) snes->atol = strtod(v, NULL);
    if ((v = opt_value("-snes_stol"))) snes->stol = strtod(v, NULL);
    if ((v = opt_value("-snes_max_it"))) snes->max_it = atoi(v);
    if ((v = opt_value("-ksp_gmres_restart"))) snes->gmres_restart = atoi(v);
    if ((v = opt_value("-snes_grid_sequence"))) snes->grid_sequence = atoi(v);
    snes->fd_color = opt_bool("-snes_fd_color", 0);
    snes->mf_operator = opt_bool("-snes_mf_operator", 0);
    if (opt_has("-snes_mf"))
        SHIM_ERR(56, "-snes_mf is not provided by the shim (-snes_fd_color or -snes_mf_operator for newtonls)");
    if (snes->mf_operator && !strcmp(snes->type, SNESKSPONLY))
        SHIM_ERR(56, "-snes_mf_operator is provided for -snes_type newtonls only");
    if (!strcmp(snes->type, SNESKSPONLY) && snes->grid_sequence)
        SHIM_ERR(56, "-snes_grid_sequence is provided for -snes_type newtonls only");
    if ((v = opt_value("-mat_is_symmetric"))) { snes->sym_tol = strtod(v, NULL); snes->sym_check = 1; }
    else if (opt_has("-mat_is_symmetric")) { snes->sym_tol = 0.0; snes->sym_check = 1; }
    if (opt_has("-pc_mg_galerkin")) SHIM_ERR(56, "-pc_mg_galerkin is not provided (levels are rediscretised, fish.c:7)");
    return 0;
}

static void print_snes_norm(SNES snes, int it, double fnorm) {
    if (snes->monitor_short) {
        if (fnorm > 1e-9) printf("  %d SNES Function norm %g\n", it, fnorm);              /* SNESMonitorDefaultShort */
        else if (fnorm > 1e-11) printf("  %d SNES Function norm %5.3e\n", it, fnorm);
        else printf("  %d SNES Function norm < 1.e-11\n", it);
    } else if (snes->monitor) {
        printf("  %d SNES Function norm %14.12e\n", it, fnorm);
    }
}

/* call the user's FormFunctionLocal on the host exactly as PETSc does (ghosted local array == the global
 * array for one rank on a non-periodic DMDA) */
static PetscErrorCode compute_function(SNES snes, Vec u, Vec F) {
    DM dm = snes->dm;
    DMDALocalInfo info;
    void *au, *aF;
    if (!dm->func) SHIM_ERR(73, "no residual callback: call DMDASNESSetFunctionLocal()");
    const double t0 = wall();
    PetscCall(DMDAGetLocalInfo(dm, &info));
    PetscCall(DMDAVecGetArrayRead(dm, u, &au));
    PetscCall(DMDAVecGetArray(dm, F, &aF));
    PetscErrorCode rc = dm->func(&info, au, aF, dm->funcctx);
    PetscCall(DMDAVecRestoreArrayRead(dm, u, &au));
    PetscCall(DMDAVecRestoreArray(dm, F, &aF));
    g_t_func += wall() - t0;
    return rc;
}

static const char *reason_name(int r) {
    switch (r) {
        case P4B_CONVERGED_RTOL: return "CONVERGED_RTOL";
        case P4B_CONVERGED_ATOL: return "CONVERGED_ATOL";
        case P4B_DIVERGED_ITS: return "DIVERGED_ITS";
        case P4B_DIVERGED_DTOL: return "DIVERGED_DTOL";
        case P4B_DIVERGED_INDEFINITE_MAT: return "DIVERGED_INDEFINITE_MAT";
        case P4B_DIVERGED_INDEFINITE_PC: return "DIVERGED_INDEFINITE_PC";
        default: return "DIVERGED_NANORINF";
    }
}

/* ---- SNESNEWTONLS: the callback-contract solve of the C ABI (p4b_snes2d_solve_monitored) --------------------------
 * What c/ch7/minimal.c:130-161 sets up -- a 2-D DMDA, FormFunctionLocal registered with DMDASNESSetFunctionLocal,
 * optionally a monitor -- becomes one library call: Newton + bt line search, GMRES/CG, multigrid on finite-difference
 * coloured Jacobians of the user's residual, -snes_grid_sequence; vectors and algebra on the device, the user's
 * callbacks on the host, called as PETSc calls them (whole grid of one logical rank, a[j][i] views). */
static SNES g_mon_snes = NULL;            /* the SNES whose monitors are running (PetscObjectGetTabLevel) */

static int newton_residual(void *user, int mx, int my, const double *u, double *F) {
    SNES snes = (SNES)user;
    DM dm = snes->dm;
    struct _p_DM cd = *dm;                /* the DMDA of this level / grid-sequence stage: same box, mx x my nodes */
    cd.M[0] = mx; cd.M[1] = my; cd.M[2] = 1;
    memset(cd.pool, 0, sizeof cd.pool);
    memset(cd.pool_busy, 0, sizeof cd.pool_busy);
    DMDALocalInfo info;
    const double t0 = wall();
    if (DMDAGetLocalInfo(&cd, &info)) return 1;
    void *au = make_tables(2, cd.M, (double *)u), *aF = make_tables(2, cd.M, F);
    PetscErrorCode rc = dm->func(&info, au, aF, dm->funcctx);
    free(au);
    free(aF);
    g_t_func += wall() - t0;
    return (int)rc;
}

static int newton_monitor(void *user, int mx, int my, int its, double fnorm, int tablevel, const double *u) {
    SNES snes = (SNES)user;
    if (!snes->nmon) return 0;
    /* SNESGetDM / SNESGetSolution inside a monitor give the current stage's grid and iterate (minimal.c:300-309) */
    struct _p_DM cd = *snes->dm;
    cd.M[0] = mx; cd.M[1] = my; cd.M[2] = 1;
    memset(cd.pool, 0, sizeof cd.pool);
    memset(cd.pool_busy, 0, sizeof cd.pool_busy);
    struct _p_Vec v;
    memset(&v, 0, sizeof v);
    v.n = (size_t)mx * my; v.dm = &cd; v.h = (double *)u; v.valid = LOC_HOST;
    DM save_dm = snes->dm;
    Vec save_sol = snes->sol;
    const int save_its = snes->its;
    snes->dm = &cd; snes->sol = &v; snes->its = its; snes->tablevel = tablevel;
    g_mon_snes = snes;
    PetscErrorCode rc = 0;
    for (int i = 0; i < snes->nmon && !rc; i++) rc = snes->mon[i].f(snes, its, fnorm, snes->mon[i].ctx);
    g_mon_snes = NULL;
    snes->dm = save_dm; snes->sol = save_sol; snes->its = save_its; snes->tablevel = 0;
    free(v.tables);
    for (int i = 0; i < 8; i++) vec_free(cd.pool[i]);
    fflush(stdout);
    return (int)rc;
}

static void newton_line(const char *line, void *ctx) { (void)ctx; puts(line); }

/* Without -snes_fd_color PETSc fills Newton's matrix (under -snes_mf_operator: the preconditioner's) by calling the Jacobian
 * callback the driver REGISTERED.  minimal.c:142-145 registers Poisson2DJacobianLocal (c/ch6/poissonfunctions.c:152-193), which
 * the library has as a kernel (p4b_poisson_stencil9: unit square, cx = cy = 1).  Before that kernel stands in for the
 * callback, the callback is run -- on the grid the solve starts on and on the one it ends on -- and what it inserts is
 * compared with the kernel's matrix: a constant star stencil with eliminated boundary columns (the "stencilcuda" recognition
 * of MatSetValuesStencil) whose three numbers are the kernel's.  Any other callback is refused, naming -snes_fd_color. */
static PetscErrorCode check_registered_poisson_jacobian(SNES snes) {
    DM dm = snes->dm;
    int gx = dm->M[0], gy = dm->M[1];
    for (int pass = 0; pass < (snes->grid_sequence > 0 ? 2 : 1); pass++) {
        if (pass == 1)
            for (int k = 0; k < snes->grid_sequence; k++) { gx = 2 * gx - 1; gy = 2 * gy - 1; }
        struct _p_DM cd = *dm;
        cd.M[0] = gx; cd.M[1] = gy; cd.M[2] = 1;
        memset(cd.pool, 0, sizeof cd.pool);
        memset(cd.pool_busy, 0, sizeof cd.pool_busy);
        const size_t n = (size_t)gx * gy;
        double *u = (double *)calloc(n, sizeof(double));
        Mat J = (Mat)calloc(1, sizeof *J);
        if (!u || !J) { free(u); free(J); SHIM_ERR(55, "out of host memory for the Jacobian check"); }
        J->dm = &cd;
        J->type = MATSTENCILCUDA;
        DMDALocalInfo info;
        PetscCall(DMDAGetLocalInfo(&cd, &info));
        void *au = make_tables(2, cd.M, u);
        const double t0 = wall();
        PetscErrorCode rc = dm->jac(&info, au, J, J, dm->jacctx);
        g_t_jac += wall() - t0;
        free(au);
        free(u);
        const double hx = 1.0 / (gx - 1), hy = 1.0 / (gy - 1), scx = hy / hx, scy = hx / hy, diag = 2.0 * (scx + scy);
        const int general = J->general, complete = J->rows_set == (long long)n && J->have_diag;
        const double dev = complete ? fmax(fabs(J->diag - diag), fmax(J->have_c[0] ? fabs(J->c[0] - scx) : (gx > 3 ? 1.0 : 0.0),
                                                                     J->have_c[1] ? fabs(J->c[1] - scy) : (gy > 3 ? 1.0 : 0.0)))
                                    : 1.0;
        char why[200];
        snprintf(why, sizeof why, "%s", general ? J->why : "");
        free(J);
        if (rc) return rc;
        if (general || !complete || !(dev <= 1.0e-12 * diag)) {
            char msg[512];
            snprintf(msg, sizeof msg, "the registered Jacobian callback is not Poisson2DJacobianLocal on the unit square with "
                     "cx = cy = 1 on the %d x %d grid (%s%s; deviation %.3e): the device path has that matrix only -- pass "
                     "-snes_fd_color to difference the residual instead", gx, gy, general ? why : "",
                     complete ? "" : " not every row was set", dev);
            SHIM_ERR(56, msg);
        }
    }
    return 0;
}

static PetscErrorCode snes_solve_newtonls(SNES snes, Vec x) {
    KSP ksp = &snes->ksp;
    PC pc = &ksp->pc;
    DM dm = snes->dm;
    if (!dm || !dm->func) SHIM_ERR(73, "no residual callback: call DMDASNESSetFunctionLocal()");
    if (dm->dim != 2 || dm->dof != 1)
        SHIM_ERR(56, "-snes_type newtonls is provided for 2-D DMDAs with one degree of freedom (minimal.c); "
                     "fish.c is linear: -snes_type ksponly (fish.c:230-231)");
    /* which matrix: -snes_fd_color differences the residual; without it PETSc calls the REGISTERED Jacobian callback --
     * minimal.c:142-145 registers Poisson2DJacobianLocal ('thus ONLY APPROXIMATE').  The library has that matrix as a
     * kernel (p4b_poisson_stencil9); the callback's rows are checked against it before it is used (below). */
    const char *mfp = opt_value("-p4b_mf_pmat");
    if (mfp && strcmp(mfp, "fd") && strcmp(mfp, "poisson")) SHIM_ERR(56, "-p4b_mf_pmat: fd or poisson");
    const int registered = !snes->fd_color && (!snes->mf_operator || (mfp && !strcmp(mfp, "poisson")));
    if (registered) {
        if (!dm->jac) SHIM_ERR(73, "no Jacobian callback registered (DMDASNESSetJacobianLocal): pass -snes_fd_color");
        PetscCall(check_registered_poisson_jacobian(snes));
    }
    p4b_minimal_opts o;
    P4B(p4b_minimal_default_opts(&o));
    if (!strcmp(ksp->type, KSPGMRES)) o.ksp_type = 0;
    else if (!strcmp(ksp->type, KSPCG)) o.ksp_type = 1;
    else SHIM_ERR(56, "-ksp_type: gmres and cg are provided on the device path");
    if (!pc->type[0])
        SHIM_ERR(56, "PETSc's default PC (ILU(0) on one rank) is sequential and not provided on the device: "
                     "pass -pc_type mg or -pc_type none");
    if (!strcmp(pc->type, PCMG)) o.pc_type = 1;
    else if (!strcmp(pc->type, PCNONE)) o.pc_type = 0;
    else SHIM_ERR(56, "newtonls: -pc_type mg and -pc_type none are provided");
    if (o.pc_type == 1) {
        if (strcmp(pc->levels_pc, "jacobi"))
            SHIM_ERR(56, "PCMG's default level smoother PC (SOR) is sequential and not provided on the device: "
                         "pass -mg_levels_pc_type jacobi");
        if (pc->smoother != P4B_SMOOTH_CHEBYSHEV) SHIM_ERR(56, "newtonls: the level smoother is chebyshev + jacobi");
        if (pc->cycle != P4B_CYCLE_V) SHIM_ERR(56, "newtonls: -pc_mg_cycle_type v only");
    }
    o.grid_x = dm->M0[0]; o.grid_y = dm->M0[1]; o.refine = dm->refine;
    o.grid_sequence = snes->grid_sequence;
    o.ksp_rtol = ksp->rtol; o.ksp_max_it = ksp->max_it; o.gmres_restart = snes->gmres_restart;
    o.mg_levels = pc->levels; o.smooth_its = pc->smooth_its;
    o.snes_rtol = snes->rtol; o.snes_stol = snes->stol; o.snes_atol = snes->atol; o.snes_max_it = snes->max_it;
    o.snes_monitor = snes->monitor_short ? 2 : (snes->monitor ? 1 : 0);
    o.snes_converged_reason = snes->converged_reason_flag;
    o.ksp_converged_reason = ksp->converged_reason_flag;
    /* -snes_mf_operator: the Krylov operator is the differenced residual ([PETSc] MatMFFD); PETSc would build the
     * preconditioner from the registered (Poisson) Jacobian callback, here it is built from the FD-coloured Jacobian of the
     * residual -- the better matrix; Newton iterates depend on it only through the inexactness of the linear solves */
    o.mf_operator = snes->mf_operator && !snes->fd_color;
    o.jacobian = registered;
    PetscCall(ensure_ctx());
    {   /* -p4b_recognise_residual 0: evaluate the registered FormFunctionLocal on the host every time, also when it is the
         * residual the library has as a kernel (p4b200.h, "Recognition") */
        const char *v = opt_value("-p4b_recognise_residual");
        P4B(p4b_tune("recognise_residual", v ? atol(v) : 1));
        v = opt_value("-p4b_gmres_cgs");                 /* GMRES orthogonalisation: 0 modified (default), 1 classical, batched dots */
        P4B(p4b_tune("gmres_cgs", v ? atol(v) : 0));
    }
    const double t0 = wall();
    PetscCall(vec_to_host(x));
    int fx = dm->M[0], fy = dm->M[1];
    for (int k = 0; k < snes->grid_sequence; k++) { fx = 2 * fx - 1; fy = 2 * fy - 1; }
    const size_t nf = (size_t)fx * fy;
    double *uf = (double *)malloc(sizeof(double) * nf);
    if (!uf) SHIM_ERR(55, "out of host memory for the solution");
    p4b_minimal_result *R = (p4b_minimal_result *)calloc(1, sizeof *R);
    fflush(stdout);
    int rc = p4b_snes2d_solve_monitored(g_ctx, &o, newton_residual, snes->nmon ? newton_monitor : NULL, snes, x->h,
                                        newton_line, NULL, uf, nf, R);
    fflush(stdout);
    g_t_snes += wall() - t0;
    /* which route ran is always said (on stderr: stdout stays what PETSc would print) */
    if (!rc && p4b_snes2d_last_route() == 1)
        fprintf(stderr, "[p4b200] SNES: the registered FormFunctionLocal equals the library's minimal-surface residual (probed "
                        "on every grid, re-verified at each converged iterate): evaluated on the device.  "
                        "-p4b_recognise_residual 0 keeps it a host callback.\n");
    if (rc) {
        free(uf);
        free(R);
        return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, p4b_last_error());
    }
    g_newton_solves++;
    snes->its = R->stage[R->nstages - 1].its;
    snes->reason = R->stage[R->nstages - 1].reason;
    ksp->its = snes->its ? R->stage[R->nstages - 1].ksp_its[snes->its - 1] : 0;
    /* under -snes_grid_sequence the SNES ends up with the refined DM and a solution on it: the caller fetches both
     * with SNESGetDM / SNESGetSolution (minimal.c:163-165) */
    if (R->mx != dm->M[0] || R->my != dm->M[1]) {
        DM fine = (DM)malloc(sizeof *fine);
        *fine = *dm;
        fine->M[0] = R->mx; fine->M[1] = R->my;
        fine->refine = dm->refine + snes->grid_sequence;
        fine->refct = 1;
        memset(fine->pool, 0, sizeof fine->pool);
        memset(fine->pool_busy, 0, sizeof fine->pool_busy);
        vec_free(snes->sol);
        snes->sol = NULL;
        DM old = dm;
        DMDestroy(&old);                      /* the SNES's reference; the caller still holds (and destroys) its own */
        snes->dm = dm = fine;
    }
    if (snes->sol && snes->sol->n != nf) { vec_free(snes->sol); snes->sol = NULL; }
    if (!snes->sol) PetscCall(vec_new(dm, &snes->sol));
    memcpy(snes->sol->h, uf, sizeof(double) * nf);
    snes->sol->valid = LOC_HOST;
    if (x->n == nf) {                         /* PETSc's SNESSolve leaves the solution in x as well */
        memcpy(x->h, uf, sizeof(double) * nf);
        x->valid = LOC_HOST;
    }
    free(uf);
    free(R);
    return 0;
}

PetscErrorCode SNESMonitorSet(SNES snes, PetscErrorCode (*f)(SNES, PetscInt, PetscReal, void *), void *mctx,
                              PetscErrorCode (*monitordestroy)(void **)) {
    if (snes->nmon >= 5) SHIM_ERR(63, "too many monitors set");
    snes->mon[snes->nmon].f = f;
    snes->mon[snes->nmon].ctx = mctx;
    snes->mon[snes->nmon].destroy = monitordestroy;
    snes->nmon++;
    return 0;
}

/* ---- what a monitor uses to print (minimal.c:333-343): one rank, stdout ---- */
struct _p_PetscViewer { int tab; };
static struct _p_PetscViewer g_stdout_viewer = {0};
PetscViewer PETSC_VIEWER_STDOUT_(MPI_Comm comm) { (void)comm; return &g_stdout_viewer; }
PetscErrorCode PetscViewerASCIIAddTab(PetscViewer viewer, PetscInt tabs) { viewer->tab += tabs; return 0; }
PetscErrorCode PetscViewerASCIISubtractTab(PetscViewer viewer, PetscInt tabs) { viewer->tab -= tabs; return 0; }
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer viewer, const char format[], ...) {
    va_list ap;
    for (int i = 0; i < viewer->tab; i++) fputs("  ", stdout);
    va_start(ap, format);
    petsc_vprintf(stdout, format, ap);
    va_end(ap);
    return 0;
}
PetscErrorCode PetscObjectGetComm(PetscObject obj, MPI_Comm *comm) { (void)obj; *comm = PETSC_COMM_WORLD; return 0; }
PetscErrorCode PetscObjectGetTabLevel(PetscObject obj, PetscInt *tab) {
    *tab = (g_mon_snes && (void *)obj == (void *)g_mon_snes) ? g_mon_snes->tablevel : 0;
    return 0;
}
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
    (void)op; (void)comm;
    if (datatype != MPIU_REAL) return 1;
    memcpy(recvbuf, sendbuf, sizeof(PetscReal) * (size_t)count);     /* one rank: every reduction is the identity */
    return 0;
}

/* fish.test2,5,8 run `-snes_fd_color -mat_is_symmetric tol`: PETSc differences the (linear) residual instead of calling
 * the Jacobian callback and reports the symmetry of the result.  The device operator comes from the Jacobian callback
 * (recognised stencil), so
 *   -snes_fd_color      is honoured by CHECKING that the callback's matrix is the Jacobian of the residual:
 *                       F(u0 + v) - F(u0) = A v for a pseudo-random v, the residual evaluated by the user's host
 *                       callback, A v by the device MatMult; a disagreement is an error, not a silent substitution
 *   -mat_is_symmetric   (x, A y) = (y, A x) for two pseudo-random vectors to the tolerance, on the device; PETSc prints
 *                       its line once for the matrix DMCreateMatrix hands out and once per Jacobian assembly */
static PetscErrorCode ksponly_matrix_checks(SNES snes, p4b_mg *mg, p4b_sell *As, Vec u, Vec F0) {
    DM dm = snes->dm;
    Vec v = NULL, w = NULL, Av = NULL, Aw = NULL;
    PetscCall(vec_new(dm, &v));
    PetscCall(vec_new(dm, &w));
    PetscCall(vec_new(dm, &Av));
    PetscCall(vec_new(dm, &Aw));
    unsigned long long lcg = 0xD1B54A32D192ED03ULL;
    for (size_t i = 0; i < v->n; i++) {
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        v->h[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        w->h[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
    }
    PetscCall(vec_to_dev(v));
    PetscCall(vec_to_dev(w));
    PetscCall(vec_to_dev(Av));
    PetscCall(vec_to_dev(Aw));
    if (As) {                                          /* the assembled type: the same checks on the SELL matrix */
        P4B(p4b_sell_spmv(As, v->d, Av->d));
        P4B(p4b_sell_spmv(As, w->d, Aw->d));
    } else {
        P4B(p4b_mg_matmult(mg, v->d, Av->d));
        P4B(p4b_mg_matmult(mg, w->d, Aw->d));
    }
    Av->valid = Aw->valid = LOC_DEV;
    if (snes->fd_color) {
        Vec up = NULL, Fp = NULL;
        PetscCall(vec_new(dm, &up));
        PetscCall(vec_new(dm, &Fp));
        PetscCall(vec_to_host(u));
        PetscCall(vec_to_host(F0));
        for (size_t i = 0; i < up->n; i++) up->h[i] = u->h[i] + v->h[i];
        PetscCall(compute_function(snes, up, Fp));
        PetscCall(vec_to_host(Av));
        double dev = 0.0, scale = 0.0;
        for (size_t i = 0; i < up->n; i++) {
            const double e = fabs((Fp->h[i] - F0->h[i]) - Av->h[i]);
            if (!(e <= dev)) dev = e;
            if (fabs(Av->h[i]) > scale) scale = fabs(Av->h[i]);
        }
        vec_free(up);
        vec_free(Fp);
        if (!(dev <= 1.0e-10 * (scale > 1.0 ? scale : 1.0))) {
            char msg[256];
            snprintf(msg, sizeof msg, "-snes_fd_color: the matrix of the Jacobian callback is not the Jacobian of the residual "
                     "(F(u+v) - F(u) - A v: max %.3e); the device path does not difference the residual of a linear problem", dev);
            SHIM_ERR(56, msg);
        }
    }
    if (snes->sym_check) {
        double xAy = 0.0, yAx = 0.0, nx = 0.0, nAy = 0.0;
        P4B(p4b_vec_dot(g_ctx, v->n, v->d, Aw->d, &xAy));
        P4B(p4b_vec_dot(g_ctx, v->n, w->d, Av->d, &yAx));
        P4B(p4b_vec_norm2(g_ctx, v->n, v->d, &nx));
        P4B(p4b_vec_norm2(g_ctx, v->n, Aw->d, &nAy));
        const int sym = fabs(xAy - yAx) <= (snes->sym_tol > 1.0e-13 ? snes->sym_tol : 1.0e-13) * nx * nAy;
        for (int k = 0; k < 2; k++)         /* DMCreateMatrix's (empty) matrix, then the assembled Jacobian */
            printf(sym || k == 0 ? "Matrix is symmetric (tolerance %g)\n" : "Matrix is not symmetric (tolerance %g)\n", snes->sym_tol);
    }
    vec_free(v);
    vec_free(w);
    vec_free(Av);
    vec_free(Aw);
    return 0;
}

/* KSPONLY with -mat_type sellcuda: the Jacobian callback's values are kept as inserted (any coefficients), converted to
 * CSR on the host and to SELL-32 on the device (p4b_sell_create); KSPCG runs here as a host loop over the library's
 * SpMV and Vec kernels ([PETSc] KSPSolve_CG, preconditioned norm).  Preconditioners: none, or jacobi (D^-1 as a
 * second, diagonal SELL matrix); PCMG needs the structured type (-mat_type stencilcuda): its coarse operators are
 * rediscretised stencils, not assembled matrices. */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
static PetscErrorCode ksponly_solve_assembled(SNES snes, Vec u, Vec F, Vec Y, double fnorm0) {
    DM dm = snes->dm;
    KSP ksp = &snes->ksp;
    PC pc = &ksp->pc;
    (void)fnorm0;
    if (strcmp(pc->type, PCNONE) && strcmp(pc->type, PCJACOBI))
        SHIM_ERR(56, "-mat_type sellcuda: -pc_type none or jacobi (PCMG coarsens the structured type: -mat_type stencilcuda)");
    Mat J = NULL;
    PetscCall(mat_new(dm, &J));
    DMDALocalInfo info;
    void *au = NULL;
    const double t_jac0 = wall();
    PetscCall(DMDAGetLocalInfo(dm, &info));
    PetscCall(DMDAVecGetArrayRead(dm, u, &au));
    PetscErrorCode rc = dm->jac(&info, au, J, J, dm->jacctx);
    PetscCall(DMDAVecRestoreArrayRead(dm, u, &au));
    if (rc) { MatDestroy(&J); return rc; }
    const size_t n = u->n;
    const int W = J->width;
    if (J->rows_set < (long long)n) { MatDestroy(&J); SHIM_ERR(73, "-mat_type sellcuda: the Jacobian callback did not set every row"); }
    /* CSR with sorted columns; the diagonal for Jacobi */
    int *rowptr = (int *)malloc(sizeof(int) * (n + 1)), *colind = (int *)malloc(sizeof(int) * n * W);
    double *vals = (double *)malloc(sizeof(double) * n * W);
    if (!rowptr || !colind || !vals) {
        free(rowptr); free(colind); free(vals);
        MatDestroy(&J);
        SHIM_ERR(55, "out of host memory for the CSR copy");
    }
    double dmin = 1e300, dmax = -1e300;
    size_t nnz = 0, ndiag = 0;
    for (size_t r = 0; r < n; r++) {
        rowptr[r] = (int)nnz;
        const int cnt = J->ccnt[r];
        int order[27];
        for (int k = 0; k < cnt; k++) order[k] = J->ccol[r * W + k];
        qsort(order, (size_t)cnt, sizeof(int), cmp_int);
        for (int k = 0; k < cnt; k++) {
            double val = 0.0;
            for (int q = 0; q < cnt; q++)
                if (J->ccol[r * W + q] == order[k]) val = J->cval[r * W + q];
            colind[nnz] = order[k];
            vals[nnz] = val;
            if ((size_t)order[k] == r) { ndiag++; if (val < dmin) dmin = val; if (val > dmax) dmax = val; }
            nnz++;
        }
    }
    rowptr[n] = (int)nnz;
    MatDestroy(&J);
    g_t_jac += wall() - t_jac0;
    const int jacobi = !strcmp(pc->type, PCJACOBI);
    if (jacobi && !(ndiag == n && dmin * dmax > 0.0)) {
        free(rowptr); free(colind); free(vals);
        SHIM_ERR(71, "-pc_type jacobi: zero (or missing) diagonal entry");
    }
    p4b_sell *A = NULL, *Dinv = NULL;
    int prc = p4b_sell_create(g_ctx, (int)n, rowptr, colind, vals, &A);
    if (!prc && jacobi) {
        /* [PETSc] PCApply_Jacobi z = D^-1 r as the SpMV of the diagonal matrix D^-1 (the library has no pointwise Vec product) */
        for (size_t r = 0; r < n; r++) {
            double d = 1.0;
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++)
                if ((size_t)colind[k] == r) d = vals[k];
            vals[r] = 1.0 / d;                      /* (vals[r] with r <= rowptr[r] is no longer needed: rows are non-empty) */
        }
        for (size_t r = 0; r <= n; r++) rowptr[r] = (int)r;
        for (size_t r = 0; r < n; r++) colind[r] = (int)r;
        prc = p4b_sell_create(g_ctx, (int)n, rowptr, colind, vals, &Dinv);
    }
    free(rowptr); free(colind); free(vals);
    if (prc) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, prc, p4b_last_error());
    if (snes->fd_color || snes->sym_check) PetscCall(ksponly_matrix_checks(snes, NULL, A, u, F));
    /* [PETSc] KSPSolve_CG on A y = F0 from y = 0 */
    Vec R = NULL, Z = NULL, P = NULL, Wv = NULL;
    PetscCall(vec_new(dm, &R)); PetscCall(vec_new(dm, &Z)); PetscCall(vec_new(dm, &P)); PetscCall(vec_new(dm, &Wv));
    PetscCall(vec_to_dev(F)); PetscCall(vec_to_dev(Y)); PetscCall(vec_to_dev(R)); PetscCall(vec_to_dev(Z));
    PetscCall(vec_to_dev(P)); PetscCall(vec_to_dev(Wv));
    const double t_ksp0 = wall();
    P4B(p4b_vec_aypx(g_ctx, n, 0.0, F->d, R->d));                    /* r = b */
    P4B(p4b_vec_aypx(g_ctx, n, -1.0, Y->d, Y->d));                   /* y = y - y = 0 */
    if (jacobi) P4B(p4b_sell_spmv(Dinv, R->d, Z->d));                /* z = M^-1 r */
    else P4B(p4b_vec_aypx(g_ctx, n, 0.0, R->d, Z->d));
    P4B(p4b_vec_aypx(g_ctx, n, 0.0, Z->d, P->d));
    double beta = 0.0, dp = 0.0;
    P4B(p4b_vec_dot(g_ctx, n, Z->d, R->d, &beta));
    P4B(p4b_vec_norm2(g_ctx, n, Z->d, &dp));
    const double ttol = ksp->rtol * dp > ksp->abstol ? ksp->rtol * dp : ksp->abstol;
    int its = 0, reason = P4B_DIVERGED_ITS;
    if (ksp->monitor_flag) printf("    %d KSP Residual norm %14.12e\n", 0, dp);
    const double dp0 = dp;
    if (dp <= ttol) reason = P4B_CONVERGED_ATOL;
    else if (beta <= 0.0) reason = P4B_DIVERGED_INDEFINITE_PC;       /* [PETSc] KSPSolve_CG: (z, r) must be positive */
    while (reason == P4B_DIVERGED_ITS && its < ksp->max_it) {
        double pw = 0.0, bnew = 0.0;
        P4B(p4b_sell_spmv(A, P->d, Wv->d));
        P4B(p4b_vec_dot(g_ctx, n, P->d, Wv->d, &pw));
        if (pw <= 0.0) { reason = P4B_DIVERGED_INDEFINITE_MAT; break; }      /* any coefficients can arrive here */
        const double a = beta / pw;
        P4B(p4b_vec_axpy(g_ctx, n, a, P->d, Y->d));
        P4B(p4b_vec_axpy(g_ctx, n, -a, Wv->d, R->d));
        if (jacobi) P4B(p4b_sell_spmv(Dinv, R->d, Z->d));
        else P4B(p4b_vec_aypx(g_ctx, n, 0.0, R->d, Z->d));
        P4B(p4b_vec_norm2(g_ctx, n, Z->d, &dp));
        its++;
        if (ksp->monitor_flag) printf("    %d KSP Residual norm %14.12e\n", its, dp);
        if (!(dp == dp)) { reason = P4B_DIVERGED_NAN; break; }
        if (dp <= ttol) { reason = dp <= ksp->abstol ? P4B_CONVERGED_ATOL : P4B_CONVERGED_RTOL; break; }
        if (dp >= 1.0e5 * dp0) { reason = P4B_DIVERGED_DTOL; break; }        /* [PETSc] -ksp_divtol default */
        P4B(p4b_vec_dot(g_ctx, n, Z->d, R->d, &bnew));
        if (bnew <= 0.0) { reason = P4B_DIVERGED_INDEFINITE_PC; break; }
        P4B(p4b_vec_aypx(g_ctx, n, bnew / beta, Z->d, P->d));          /* p = z + (beta_new / beta) p */
        beta = bnew;
    }
    g_t_ksp += wall() - t_ksp0;
    Y->valid = LOC_DEV;
    ksp->its = its;
    ksp->reason = reason;
    if (ksp->converged_reason_flag) {
        if (reason > 0) printf("    Linear solve converged due to %s iterations %d\n", reason_name(reason), its);
        else printf("    Linear solve did not converge due to %s iterations %d\n", reason_name(reason), its);
    }
    p4b_sell_destroy(A);
    if (Dinv) p4b_sell_destroy(Dinv);
    vec_free(R); vec_free(Z); vec_free(P); vec_free(Wv);
    /* u = u0 - y, the post-solve norm, the reason line: as the structured path */
    PetscCall(VecAXPY(u, -1.0, Y));
    snes->its = 1;
    if (snes->monitor_short || snes->monitor) {
        double fnorm;
        PetscCall(compute_function(snes, u, F));
        PetscCall(VecNorm(F, NORM_2, &fnorm));
        print_snes_norm(snes, 1, fnorm);
    }
    if (snes->converged_reason_flag) printf("Nonlinear solve converged due to CONVERGED_ITS iterations 1\n");
    return 0;
}

/* ---- -p4b_gpus N: the KSP solve of the unchanged fish.c on N GPUs of this box -------------------------------------
 * fish.c creates its DMDA on PETSC_COMM_WORLD and the reference runs it under `mpiexec -n P` (c/testit.sh:22,
 * c/ch6/makefile:21,30).  There is no MPI here; SURVEY.md 5 proposed "one host thread drives all GPUs; callbacks see a
 * single logical rank owning the whole grid".  So the callbacks (F(u0), the level Jacobians) run once on the host over
 * the whole grid, and the solve -- where the time is -- runs on z-slabs: one host thread per GPU, each with its own
 * context, stream and slab of the hierarchy, talking through the library's communicator (peer memory over NVLink fused
 * into the kernels, csrc/comm.h; the threads map each other's memory directly instead of through CUDA IPC). */
#include <pthread.h>
struct mg_worker {
    int rank, nranks, dim, nlev, pct, max_it, rc;
    unsigned char id[128];
    p4b_grid g;
    p4b_mg_opts o;
    const double *coef, *F;
    double *Y, rtol, abstol;
    p4b_ksp_result res;
    char err[512];
};
static void *mg_worker_main(void *arg) {
    struct mg_worker *w = (struct mg_worker *)arg;
    p4b_ctx *ctx = NULL;
    p4b_mg *mg = NULL;
    double *b = NULL, *x = NULL;
    int zs = 0, zm = 0;
    size_t nloc = 0;
#define WK(call) do { if (!w->rc && (call)) { w->rc = 70; snprintf(w->err, sizeof w->err, "%s", p4b_last_error()); } } while (0)
    WK(p4b_ctx_create_own_stream(w->rank, &ctx));
    WK(p4b_comm_init(ctx, w->id, w->rank, w->nranks));
    WK(p4b_mg_create_stencil(ctx, &w->g, &w->o, w->coef, w->nlev, &mg));
    if (!w->rc) WK(p4b_mg_local_range(mg, &zs, &zm, &nloc));
    if (!w->rc) {
        const size_t plane = w->dim == 3 ? (size_t)w->g.mx * w->g.my : (size_t)w->g.mx;      /* nodes per slab index */
        const size_t off = (size_t)zs * plane;
        WK(p4b_malloc(ctx, nloc * sizeof(double), (void **)&b));
        WK(p4b_malloc(ctx, nloc * sizeof(double), (void **)&x));
        WK(p4b_memcpy_h2d(ctx, b, w->F + off, nloc * sizeof(double)));
        WK(p4b_cg_solve(mg, w->pct, b, x, w->rtol, w->abstol, w->max_it, &w->res));
        WK(p4b_memcpy_d2h(ctx, w->Y + off, x, nloc * sizeof(double)));
    }
    /* collective teardown: every thread gets here (a failed rank would otherwise leave the others waiting) */
    if (mg) p4b_mg_destroy(mg);
    if (b) p4b_free(ctx, b);
    if (x) p4b_free(ctx, x);
    if (ctx) p4b_ctx_destroy(ctx);
#undef WK
    return NULL;
}
static PetscErrorCode ksp_solve_multi_gpu(int ngpu, const p4b_grid *g, const p4b_mg_opts *o, const double *coef, int nlev,
                                          int pct, const double *F, double *Y, double rtol, double abstol, int max_it,
                                          p4b_ksp_result *res) {
    struct mg_worker *w = (struct mg_worker *)calloc((size_t)ngpu, sizeof *w);
    pthread_t *th = (pthread_t *)calloc((size_t)ngpu, sizeof *th);
    if (!w || !th) { free(w); free(th); SHIM_ERR(55, "out of memory"); }
    unsigned char id[128];
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);      /* NCCL's banner ("NCCL version ...") must not land in the driver's stdout */
    if (p4b_comm_unique_id(id)) { free(w); free(th); SHIM_ERR(70, p4b_last_error()); }
    for (int r = 0; r < ngpu; r++) {
        w[r].rank = r; w[r].nranks = ngpu; w[r].dim = g->dim; w[r].nlev = nlev; w[r].pct = pct; w[r].max_it = max_it;
        memcpy(w[r].id, id, sizeof id);
        w[r].g = *g; w[r].o = *o; w[r].coef = coef; w[r].F = F; w[r].Y = Y; w[r].rtol = rtol; w[r].abstol = abstol;
        if (pthread_create(&th[r], NULL, mg_worker_main, &w[r])) { free(w); free(th); SHIM_ERR(55, "cannot create a host thread"); }
    }
    for (int r = 0; r < ngpu; r++) pthread_join(th[r], NULL);
    char msg[640] = "";
    for (int r = 0; r < ngpu && !msg[0]; r++)
        if (w[r].rc) snprintf(msg, sizeof msg, "-p4b_gpus %d, rank %d: %s", ngpu, r, w[r].err);
    *res = w[0].res;                             /* the Krylov scalars are all-reduced: identical on every rank */
    for (int r = 1; r < ngpu; r++)
        if (w[r].res.solve_ms > res->solve_ms) res->solve_ms = w[r].res.solve_ms;
    free(w); free(th);
    if (msg[0]) SHIM_ERR(70, msg);
    return 0;
}

/* -p4b_gpus N for the unchanged pattern.c: the time stepping of the recognised model on y-slabs, one host thread per GPU
 * (p4b_pattern_solve_from on a communicator: ring ghost rows, all-reduced dots).  Rank 0 prints. */
struct ts_worker {
    int rank, nranks, m, rc;
    unsigned char id[128];
    p4b_pattern_opts o;
    double *Y;                   /* host: the whole grid, 2 m m doubles; every rank reads and writes its rows */
    p4b_pattern_result res;
    char err[512];
};
static void newton_line_fwd(const char *line, void *ctx) { (void)ctx; puts(line); }
static void *ts_worker_main(void *arg) {
    struct ts_worker *w = (struct ts_worker *)arg;
    p4b_ctx *ctx = NULL;
    double *y = NULL;
    const size_t rows = (size_t)(w->m / w->nranks), nloc = 2 * (size_t)w->m * rows, off = nloc * (size_t)w->rank;
#define WK(call) do { if (!w->rc && (call)) { w->rc = 70; snprintf(w->err, sizeof w->err, "%s", p4b_last_error()); } } while (0)
    WK(p4b_ctx_create_own_stream(w->rank, &ctx));
    WK(p4b_comm_init(ctx, w->id, w->rank, w->nranks));
    WK(p4b_malloc(ctx, nloc * sizeof(double), (void **)&y));
    WK(p4b_memcpy_h2d(ctx, y, w->Y + off, nloc * sizeof(double)));
    WK(p4b_pattern_solve_from(ctx, &w->o, y, w->rank == 0 ? newton_line_fwd : NULL, NULL, y, nloc, &w->res));
    WK(p4b_memcpy_d2h(ctx, w->Y + off, y, nloc * sizeof(double)));
    if (y) p4b_free(ctx, y);
    if (ctx) p4b_ctx_destroy(ctx);
#undef WK
    return NULL;
}
static PetscErrorCode ts_solve_multi_gpu(int ngpu, int m, const p4b_pattern_opts *o, double *Y, p4b_pattern_result *res) {
    if (m % ngpu || (m / ngpu) % 2) {
        char msg[256];
        snprintf(msg, sizeof msg, "-p4b_gpus %d: the %d rows of the grid must split into an even number of rows per GPU", ngpu, m);
        SHIM_ERR(60, msg);
    }
    struct ts_worker *w = (struct ts_worker *)calloc((size_t)ngpu, sizeof *w);
    pthread_t *th = (pthread_t *)calloc((size_t)ngpu, sizeof *th);
    if (!w || !th) { free(w); free(th); SHIM_ERR(55, "out of memory"); }
    unsigned char id[128];
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    if (p4b_comm_unique_id(id)) { free(w); free(th); SHIM_ERR(70, p4b_last_error()); }
    for (int r = 0; r < ngpu; r++) {
        w[r].rank = r; w[r].nranks = ngpu; w[r].m = m; w[r].o = *o; w[r].Y = Y;
        memcpy(w[r].id, id, sizeof id);
        if (pthread_create(&th[r], NULL, ts_worker_main, &w[r])) { free(w); free(th); SHIM_ERR(55, "cannot create a host thread"); }
    }
    for (int r = 0; r < ngpu; r++) pthread_join(th[r], NULL);
    char msg[640] = "";
    for (int r = 0; r < ngpu && !msg[0]; r++)
        if (w[r].rc) snprintf(msg, sizeof msg, "-p4b_gpus %d, rank %d: %s", ngpu, r, w[r].err);
    *res = w[0].res;
    free(w); free(th);
    if (msg[0]) SHIM_ERR(70, msg);
    return 0;
}

PetscErrorCode SNESSolve(SNES snes, Vec b, Vec x) {
    if (b) SHIM_ERR(56, "SNESSolve with a right-hand side is not provided");
    if (snes->dm && (snes->dm->b[0] == DM_BOUNDARY_PERIODIC || snes->dm->b[1] == DM_BOUNDARY_PERIODIC))
        SHIM_ERR(56, "SNESSolve is provided on DM_BOUNDARY_NONE DMDAs (fish.c:203-213, minimal.c:130-134)");
    if (!strcmp(snes->type, SNESNEWTONLS)) return snes_solve_newtonls(snes, x);
    if (strcmp(snes->type, SNESKSPONLY))
        SHIM_ERR(56, "-snes_type: ksponly (fish.c:230-231) and newtonls (minimal.c) are provided by the shim");
    if (strcmp(snes->ksp.type, KSPCG)) SHIM_ERR(56, "only -ksp_type cg is provided on the device path");
    KSP ksp = &snes->ksp;
    PC pc = &ksp->pc;
    if (!pc->type[0])
        SHIM_ERR(56, "PETSc's default PC (ILU(0) on one rank) is sequential and not provided on the device: "
                     "pass -pc_type mg (or jacobi, none)");
    if (!strcmp(pc->type, PCMG) && strcmp(pc->levels_pc, "jacobi"))
        SHIM_ERR(56, "PCMG's default level smoother PC (SOR) is sequential and not provided on the device: "
                     "pass -mg_levels_pc_type jacobi");
    DM dm = snes->dm;
    if (!dm || !dm->jac) SHIM_ERR(73, "no Jacobian callback: call DMDASNESSetJacobianLocal()");
    PetscCall(ensure_ctx());
    const double t_snes0 = wall();

    /* the solution vector is owned by the SNES (fish.c:248) */
    if (!snes->sol) PetscCall(vec_new(dm, &snes->sol));
    PetscCall(vec_to_host(x));
    memcpy(snes->sol->h, x->h, x->n * sizeof(double));
    snes->sol->valid = LOC_HOST;
    Vec u = snes->sol, F = NULL, Y = NULL;
    PetscCall(vec_new(dm, &F));
    PetscCall(vec_new(dm, &Y));

    /* F0 = F(u0) */
    PetscCall(compute_function(snes, u, F));
    double fnorm;
    PetscCall(VecNorm(F, NORM_2, &fnorm));
    print_snes_norm(snes, 0, fnorm);

    {
        const char *mt = opt_value("-mat_type");
        if (mt && !strcmp(mt, MATSELLCUDA)) {
            PetscErrorCode rcs = ksponly_solve_assembled(snes, u, F, Y, fnorm);
            vec_free(F);
            vec_free(Y);
            if (rcs) return rcs;
            PetscCall(vec_to_host(u));
            memcpy(x->h, u->h, x->n * sizeof(double));
            x->valid = LOC_HOST;
            g_t_snes += wall() - t_snes0;
            fflush(stdout);
            return 0;
        }
    }
    /* Jacobian on every level by rediscretisation: the user's callback on the coarsened DMDAs (PCSetUp_MG) */
    const double t_jac0 = wall();
    int nlev = 1;
    int Ml[P4B_MAX_LEVELS][3];
    memcpy(Ml[0], dm->M, sizeof(int) * 3);
    if (!strcmp(pc->type, PCMG)) {
        /* [PETSc] PCSetUp_MG on a DMDA: without -pc_mg_levels there are refine+1 levels, the DMDA as created (before
         * -da_refine) being the coarsest */
        const int want = pc->levels > 0 ? pc->levels : dm->refine + 1;
        for (;;) {
            if (nlev >= want) break;
            int ok = 1;
            for (int d = 0; d < dm->dim; d++)
                if (Ml[nlev - 1][d] <= 3 || (Ml[nlev - 1][d] - 1) % 2) ok = 0;
            if (!ok || nlev >= P4B_MAX_LEVELS) break;
            for (int d = 0; d < 3; d++) Ml[nlev][d] = d < dm->dim ? (Ml[nlev - 1][d] - 1) / 2 + 1 : 1;
            nlev++;
        }
        if (pc->levels > 0 && nlev < pc->levels) SHIM_ERR(62, "cannot coarsen to the requested -pc_mg_levels");
    }
    double coef[P4B_MAX_LEVELS * 4];
    for (int l = 0; l < nlev; l++) {
        struct _p_DM cd = *dm;                    /* coarsened DMDA: same box, fewer nodes */
        memcpy(cd.M, Ml[l], sizeof(int) * 3);
        memset(cd.pool, 0, sizeof cd.pool);
        DMDALocalInfo info;
        PetscCall(DMDAGetLocalInfo(&cd, &info));
        Mat J;
        PetscCall(mat_new(&cd, &J));
        Vec ul = NULL;
        void *au = NULL;
        if (l == 0) ul = u;
        else PetscCall(vec_new(&cd, &ul));        /* PETSc hands the callback a state on that level */
        PetscCall(DMDAVecGetArrayRead(&cd, ul, &au));
        PetscErrorCode rc = dm->jac(&info, au, J, J, dm->jacctx);
        PetscCall(DMDAVecRestoreArrayRead(&cd, ul, &au));
        if (l) vec_free(ul);
        if (rc) { MatDestroy(&J); vec_free(F); vec_free(Y); return rc; }
        if (J->general || !J->have_diag || J->rows_set != (long long)da_n(&cd)) {
            char msg[640];
            snprintf(msg, sizeof msg,
                     "Jacobian on level %d is not the constant-coefficient Dirichlet stencil the device Mat type "
                     "\"stencilcuda\" represents (%s; %lld of %zu rows set); pass -mat_type sellcuda for the assembled device path",
                     l, J->general ? J->why : "incomplete", J->rows_set, da_n(&cd));
            MatDestroy(&J);
            vec_free(F);
            vec_free(Y);
            SHIM_ERR(56, msg);
        }
        coef[4 * l + 0] = J->diag;
        coef[4 * l + 1] = J->have_c[0] ? J->c[0] : 0.0;
        coef[4 * l + 2] = J->have_c[1] ? J->c[1] : 0.0;
        coef[4 * l + 3] = J->have_c[2] ? J->c[2] : 0.0;
        /* grids with a single interior node per direction never show an off-diagonal: harmless, it is unused */
        MatDestroy(&J);
    }
    g_t_jac += wall() - t_jac0;

    /* KSPSolve: J y = F0 on the device */
    p4b_grid g;
    {
        struct _p_Mat tmp;
        memset(&tmp, 0, sizeof tmp);
        tmp.dm = dm;
        PetscCall(mat_to_grid(&tmp, &g));
    }
    p4b_mg_opts o;
    p4b_mg_default_opts(&o);
    o.levels = nlev;
    o.cycle = pc->cycle; o.smoother = pc->smoother; o.smooth_its = pc->smooth_its;
    if (pc->have_eig) { o.emin = pc->emin; o.emax = pc->emax; }
    o.est_lo = pc->est_lo; o.est_hi = pc->est_hi; o.fuse = pc->fuse;
    const int pct = !strcmp(pc->type, PCMG) ? P4B_PC_MG : (!strcmp(pc->type, PCJACOBI) ? P4B_PC_JACOBI : P4B_PC_NONE);
    int ngpu = 1;
    {
        const char *v = opt_value("-p4b_gpus");
        if (!v) v = getenv("P4B_GPUS");
        if (v) ngpu = atoi(v);
        if (ngpu < 1) SHIM_ERR(62, "-p4b_gpus must be at least 1");
        if (ngpu > 1 && dm->dim == 1) SHIM_ERR(62, "-p4b_gpus: 1-D grids are not distributed");
    }
    p4b_mg *mg = NULL;
    p4b_ksp_result res;
    if (ngpu > 1) {
        /* N GPUs, one host thread each (see ksp_solve_multi_gpu); the matrix checks below need the one-GPU hierarchy */
        if (snes->fd_color || snes->sym_check) {
            P4B(p4b_mg_create_stencil(g_ctx, &g, &o, coef, nlev, &mg));
            PetscCall(ksponly_matrix_checks(snes, mg, NULL, u, F));
            P4B(p4b_mg_destroy(mg));
            mg = NULL;
        }
        PetscCall(vec_to_host(F));
        PetscCall(vec_to_host(Y));
        const double t_ksp0 = wall();
        PetscCall(ksp_solve_multi_gpu(ngpu, &g, &o, coef, nlev, pct, F->h, Y->h, ksp->rtol, ksp->abstol, ksp->max_it, &res));
        g_t_ksp += wall() - t_ksp0;
        g_t_ksp_dev_ms += res.solve_ms;
        Y->valid = LOC_HOST;
    } else {
    P4B(p4b_mg_create_stencil(g_ctx, &g, &o, coef, nlev, &mg));
    if (snes->fd_color || snes->sym_check) PetscCall(ksponly_matrix_checks(snes, mg, NULL, u, F));
    PetscCall(vec_to_dev(F));
    PetscCall(vec_to_dev(Y));
    const double t_ksp0 = wall();
    P4B(p4b_cg_solve(mg, pct, F->d, Y->d, ksp->rtol, ksp->abstol, ksp->max_it, &res));
    g_t_ksp += wall() - t_ksp0;
    g_t_ksp_dev_ms += res.solve_ms;
    Y->valid = LOC_DEV;
    }
    ksp->its = res.its;
    ksp->reason = res.reason;
    if (ksp->monitor_flag)
        for (int i = 0; i < res.nhist; i++) printf("    %d KSP Residual norm %14.12e\n", i, res.hist[i]);
    if (ksp->converged_reason_flag) {
        if (res.reason > 0) printf("    Linear solve converged due to %s iterations %d\n", reason_name(res.reason), res.its);
        else printf("    Linear solve did not converge due to %s iterations %d\n", reason_name(res.reason), res.its);
    }
    if (mg) P4B(p4b_mg_destroy(mg));

    /* u = u0 - y ; then the post-solve function norm PETSc's KSPONLY monitor prints */
    PetscCall(VecAXPY(u, -1.0, Y));
    snes->its = 1;
    if (snes->monitor_short || snes->monitor) {
        PetscCall(compute_function(snes, u, F));
        PetscCall(VecNorm(F, NORM_2, &fnorm));
        print_snes_norm(snes, 1, fnorm);
    }
    if (snes->converged_reason_flag)
        printf("Nonlinear solve converged due to CONVERGED_ITS iterations 1\n");
    vec_free(F);
    vec_free(Y);
    /* PETSc's SNESSolve leaves the solution in x as well */
    PetscCall(vec_to_host(u));
    memcpy(x->h, u->h, x->n * sizeof(double));
    x->valid = LOC_HOST;
    g_t_snes += wall() - t_snes0;
    fflush(stdout);
    return 0;
}

PetscErrorCode SNESDestroy(SNES *snes) {
    if (!snes || !*snes) return 0;
    for (int i = 0; i < (*snes)->nmon; i++)
        if ((*snes)->mon[i].destroy) (*snes)->mon[i].destroy(&(*snes)->mon[i].ctx);
    vec_free((*snes)->sol);
    if ((*snes)->dm) { DM d = (*snes)->dm; DMDestroy(&d); }
    free(*snes);
    *snes = NULL;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* TS: the unchanged c/ch5/pattern.c                                                                 */
/* ------------------------------------------------------------------------------------------------ */
/* pattern.c registers four host callbacks on a periodic 2-D DMDA with two components (pattern.c:103-114) and calls
 * TSSolve.  The device path for this system is p4b_pattern_solve_from (include/p4b200.h): TSARKIMEX3 / TSTHETA with the
 * residuals and the matrix-free stage Jacobian as CUDA kernels of the two-species reaction-diffusion model
 *     F(Y, Ydot) = Ydot - D_c L9 Y / (6 h^2),    G(Y) = (-u v^2 + phi (1 - u),  u v^2 - (phi + kappa) v).
 * TSSolve therefore (1) IDENTIFIES the model's five numbers from the registered callbacks (three probe evaluations on
 * the host: the box side from the DMDA, phi and kappa from G at constant states, D_u and D_v from F of a unit pulse),
 * (2) VERIFIES that the callbacks ARE that model: F and G evaluated by the user's functions on the host and by the
 * device kernels at a generic state must agree to rounding, and every row the user's Jacobian callbacks insert is
 * compared with the stage matrix the device applies (rd_check_rows above) -- the same recognise-or-refuse contract as
 * the "stencilcuda" Mat type of fish.c, and (3) runs the solve on the device from the caller's initial state.
 * Callbacks that are anything else are refused with the measured deviation: there is no general assembled
 * block-stencil TS path (and no CPU fallback).  Only the callbacks PETSc itself would call for the chosen -ts_type
 * are invoked (RHSJacobian is never called for the IMEX default), so a driver's own call-back report (pattern.c:127-135)
 * prints what it prints under PETSc. */
struct _p_TS {
    DM dm;
    char type[32];
    void *appctx;
    double t0, max_time, dt, rtol, atol;
    int max_steps, monitor, eft;
    SNES snes;               /* holds the KSP / PC / SNES options of the stage solves */
    int nmon;                /* TSMonitorSet monitors (heat.c:72): run at step 0 and after every accepted step */
    struct { PetscErrorCode (*f)(TS, PetscInt, PetscReal, Vec, void *); void *ctx; PetscErrorCode (*destroy)(void **); } mon[5];
    double tcur;             /* time of the monitor call in progress (TSGetTime); t0 outside a solve */
    int in_solve;
};

PetscErrorCode DMDASetFieldName(DM da, PetscInt nf, const char name[]) { (void)da; (void)nf; (void)name; return 0; }
PetscErrorCode DMDAGetCoordinateArray(DM da, void *xc) {
    if (da->dim != 2) SHIM_ERR(56, "DMDAGetCoordinateArray: 2-D DMDAs only");
    const int mx = da->M[0], my = da->M[1];
    /* [PETSc] DMDASetUniformCoordinates: periodic directions have mx cells, the others mx - 1 */
    const double hx = (da->cmax[0] - da->cmin[0]) / (da->b[0] == DM_BOUNDARY_PERIODIC ? mx : mx - 1);
    const double hy = (da->cmax[1] - da->cmin[1]) / (da->b[1] == DM_BOUNDARY_PERIODIC ? my : my - 1);
    char *blk = (char *)malloc(sizeof(DMDACoor2d *) * (size_t)my + sizeof(DMDACoor2d) * (size_t)mx * my);
    if (!blk) SHIM_ERR(55, "out of host memory for the coordinate array");
    DMDACoor2d **rows = (DMDACoor2d **)blk;
    DMDACoor2d *c = (DMDACoor2d *)(blk + sizeof(DMDACoor2d *) * (size_t)my);
    for (int j = 0; j < my; j++) {
        rows[j] = c + (size_t)j * mx;
        for (int i = 0; i < mx; i++) { rows[j][i].x = da->cmin[0] + i * hx; rows[j][i].y = da->cmin[1] + j * hy; }
    }
    *(void **)xc = rows;
    return 0;
}
PetscErrorCode DMDARestoreCoordinateArray(DM da, void *xc) {
    (void)da;
    free(*(void **)xc);
    *(void **)xc = NULL;
    return 0;
}
PetscErrorCode DMDATSSetRHSFunctionLocal(DM dm, InsertMode imode, DMDATSRHSFunctionLocal func, void *ctx) {
    (void)imode; dm->rhsfunc = func; dm->rhsfuncctx = ctx; return 0;
}
PetscErrorCode DMDATSSetRHSJacobianLocal(DM dm, DMDATSRHSJacobianLocal func, void *ctx) {
    dm->rhsjac = func; dm->rhsjacctx = ctx; return 0;
}
PetscErrorCode DMDATSSetIFunctionLocal(DM dm, InsertMode imode, DMDATSIFunctionLocal func, void *ctx) {
    (void)imode; dm->ifunc = func; dm->ifuncctx = ctx; return 0;
}
PetscErrorCode DMDATSSetIJacobianLocal(DM dm, DMDATSIJacobianLocal func, void *ctx) {
    dm->ijac = func; dm->ijacctx = ctx; return 0;
}

PetscErrorCode TSCreate(MPI_Comm comm, TS *ts) {
    TS t = (TS)calloc(1, sizeof *t);
    snprintf(t->type, 32, "%s", TSBEULER);                  /* [PETSc] the default TS type is backward Euler */
    t->max_time = 5.0; t->dt = 0.1; t->max_steps = 5000; t->rtol = t->atol = 1.0e-4;     /* [PETSc] TS defaults */
    t->eft = TS_EXACTFINALTIME_UNSPECIFIED;
    PetscCall(SNESCreate(comm, &t->snes));
    *ts = t;
    return 0;
}
PetscErrorCode TSSetProblemType(TS ts, TSProblemType type) { (void)ts; (void)type; return 0; }
PetscErrorCode TSSetDM(TS ts, DM dm) { ts->dm = dm; dm->refct++; return 0; }
PetscErrorCode TSSetApplicationContext(TS ts, void *usrP) { ts->appctx = usrP; return 0; }
PetscErrorCode TSSetType(TS ts, TSType type) { snprintf(ts->type, 32, "%s", type); return 0; }
PetscErrorCode TSGetType(TS ts, TSType *type) { *type = ts->type; return 0; }
PetscErrorCode TSSetTime(TS ts, PetscReal t) { ts->t0 = t; return 0; }
PetscErrorCode TSSetMaxTime(TS ts, PetscReal maxtime) { ts->max_time = maxtime; return 0; }
PetscErrorCode TSSetTimeStep(TS ts, PetscReal time_step) { ts->dt = time_step; return 0; }
PetscErrorCode TSSetExactFinalTime(TS ts, TSExactFinalTimeOption eftopt) { ts->eft = (int)eftopt; return 0; }
PetscErrorCode TSSetFromOptions(TS ts) {
    const char *v;
    if ((v = opt_value("-ts_type"))) snprintf(ts->type, 32, "%s", v);
    if ((v = opt_value("-ts_dt"))) ts->dt = strtod(v, NULL);
    if ((v = opt_value("-ts_max_time"))) ts->max_time = strtod(v, NULL);
    if ((v = opt_value("-ts_max_steps"))) ts->max_steps = atoi(v);
    if ((v = opt_value("-ts_rtol"))) ts->rtol = strtod(v, NULL);
    if ((v = opt_value("-ts_atol"))) ts->atol = strtod(v, NULL);
    if ((v = opt_value("-ts_exact_final_time"))) {
        if (!strcmp(v, "matchstep")) ts->eft = TS_EXACTFINALTIME_MATCHSTEP;
        else if (!strcmp(v, "stepover")) ts->eft = TS_EXACTFINALTIME_STEPOVER;
        else if (!strcmp(v, "interpolate")) ts->eft = TS_EXACTFINALTIME_INTERPOLATE;
        else SHIM_ERR(62, "-ts_exact_final_time: stepover, interpolate or matchstep");
    }
    ts->monitor = opt_has("-ts_monitor") && !binary_target("-ts_monitor");      /* a binary viewer prints nothing */
    /* KSP / PC / SNES options of the stage solves.  -snes_fd_color (pattern.test3) asks PETSc to difference the residual
     * instead of calling the Jacobian callbacks: the device stage operator IS the analytic Jacobian those differences
     * approximate, so the option only means that the Jacobian callbacks are not called (nor checked) */
    return SNESSetFromOptions(ts->snes);
}
PetscErrorCode TSDestroy(TS *ts) {
    if (!ts || !*ts) return 0;
    if ((*ts)->dm) { DM d = (*ts)->dm; DMDestroy(&d); }
    SNESDestroy(&(*ts)->snes);
    free(*ts);
    *ts = NULL;
    return 0;
}

/* ghosted (width 1) host copy of a field with dof components and a[j][i] views of it that are valid for j, i in [-1, m]:
 * what DMGlobalToLocal + DMDAVecGetArray hand a callback.  Ghost nodes of a periodic direction hold the wrapped values; a
 * DM_BOUNDARY_NONE direction has no ghost nodes in PETSc -- the cells exist here and hold zeros, which no correct callback
 * reads (heat.c:152-155 mirrors at i = 0 and i = mx-1 instead). */
struct ghosted { double *buf; double **rows; void *a; };
static int ghosted_make(const double *Y, int mx, int my, int dof, int px, int py, struct ghosted *g) {
    const int gx = mx + 2, gy = my + 2;
    g->buf = (double *)calloc((size_t)gx * gy * dof, sizeof(double));
    g->rows = (double **)malloc(sizeof(double *) * (size_t)gy);
    if (!g->buf || !g->rows) return 1;
    for (int jj = 0; jj < gy; jj++) {
        double *row = g->buf + (size_t)jj * gx * dof;
        g->rows[jj] = row + dof;                     /* a[j][0] is the first owned node; a[j][-1] the left ghost */
        if (!py && (jj == 0 || jj == gy - 1)) continue;
        const int j = (jj - 1 + my) % my;
        memcpy(row + dof, Y + (size_t)j * mx * dof, sizeof(double) * (size_t)mx * dof);
        if (px) {
            memcpy(row, Y + ((size_t)j * mx + (mx - 1)) * dof, sizeof(double) * dof);
            memcpy(row + (size_t)(mx + 1) * dof, Y + (size_t)j * mx * dof, sizeof(double) * dof);
        }
    }
    g->a = g->rows + 1;                              /* a[-1] is the lower ghost row */
    return 0;
}
#define DM_PX(dm) ((dm)->b[0] == DM_BOUNDARY_PERIODIC)
#define DM_PY(dm) ((dm)->b[1] == DM_BOUNDARY_PERIODIC)
static void ghosted_free(struct ghosted *g) { free(g->buf); free(g->rows); g->buf = NULL; g->rows = NULL; }

/* the user's G(t, Y) and F(t, Y, Ydot) on host arrays of the whole grid */
static PetscErrorCode ts_eval_rhs(DM dm, double t, const double *Y, double *G) {
    DMDALocalInfo info;
    struct ghosted gY;
    PetscCall(DMDAGetLocalInfo(dm, &info));
    if (ghosted_make(Y, dm->M[0], dm->M[1], dm->dof, DM_PX(dm), DM_PY(dm), &gY)) SHIM_ERR(55, "out of host memory");
    void *aG = make_tables_dof(2, dm->M, dm->dof, G);
    const double t0 = wall();
    PetscErrorCode rc = dm->rhsfunc(&info, t, gY.a, aG, dm->rhsfuncctx);
    g_t_func += wall() - t0;
    free(aG);
    ghosted_free(&gY);
    return rc;
}
static PetscErrorCode ts_eval_ifunc(DM dm, double t, const double *Y, const double *Ydot, double *F) {
    DMDALocalInfo info;
    struct ghosted gY, gD;
    if (!dm->ifunc) {                        /* no IFunction registered: [PETSc] F = Ydot, the system is Ydot = G(t, Y) */
        memcpy(F, Ydot, sizeof(double) * (size_t)dm->M[0] * dm->M[1] * dm->dof);
        return 0;
    }
    PetscCall(DMDAGetLocalInfo(dm, &info));
    if (ghosted_make(Y, dm->M[0], dm->M[1], dm->dof, DM_PX(dm), DM_PY(dm), &gY)) SHIM_ERR(55, "out of host memory");
    if (ghosted_make(Ydot, dm->M[0], dm->M[1], dm->dof, DM_PX(dm), DM_PY(dm), &gD)) SHIM_ERR(55, "out of host memory");
    void *aF = make_tables_dof(2, dm->M, dm->dof, F);
    const double t0 = wall();
    PetscErrorCode rc = dm->ifunc(&info, t, gY.a, gD.a, aF, dm->ifuncctx);
    g_t_func += wall() - t0;
    free(aF);
    ghosted_free(&gY);
    ghosted_free(&gD);
    return rc;
}

static double maxabs_diff(const double *a, const double *b, size_t n, double *scale) {
    double d = 0.0, s = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double e = fabs(a[i] - b[i]);
        if (!(e <= d)) d = e;                        /* NaN propagates into the deviation */
        if (fabs(b[i]) > s) s = fabs(b[i]);
    }
    *scale = s;
    return d;
}

/* everything TSSolve allocates for its probes, so that one release covers every exit */
struct ts_work {
    double *Y, *D, *Fu, *Fd;                     /* host: probe state, probe Ydot, user's result, device result */
    double *dY, *dD, *dF;                        /* device copies */
    struct ghosted gY, gD;
    p4b_pattern_result *R;
};
static void ts_work_release(struct ts_work *W) {
    free(W->Y); free(W->D); free(W->Fu); free(W->Fd);
    if (g_ctx) {
        if (W->dY) p4b_free(g_ctx, W->dY);
        if (W->dD) p4b_free(g_ctx, W->dD);
        if (W->dF) p4b_free(g_ctx, W->dF);
    }
    ghosted_free(&W->gY);
    ghosted_free(&W->gD);
    free(W->R);
    memset(W, 0, sizeof *W);
}

/* TSSolve for callbacks that are NOT the model the library has as kernels: p4b_ts2d_solve -- the same time steppers with
 * F and G evaluated by the user's host callbacks and the stage operator the differenced residual ([PETSc] MatMFFD), vectors
 * and algebra on the device.  No Jacobian matrix means no multigrid (-pc_type none), and the differencing assumes
 * F = M Ydot + f(Y) with a constant M, which is checked here on the callbacks (five evaluations). */
static int ts_ifn(void *user, int m, double t, const double *Y, const double *Ydot, double *F) {
    TS ts = (TS)user;
    (void)m;
    return ts_eval_ifunc(ts->dm, t, Y, Ydot, F) != 0;
}
static int ts_gfn(void *user, int m, double t, const double *Y, double *G) {
    TS ts = (TS)user;
    (void)m;
    return ts_eval_rhs(ts->dm, t, Y, G) != 0;
}
static PetscErrorCode ts_solve_general(TS ts, Vec x, struct ts_work *W, p4b_pattern_opts *o, const char *why) {
    DM dm = ts->dm;
    KSP ksp = &ts->snes->ksp;
    const size_t n = x->n;
    char msg[768];
    const int pattern_shape = dm->dof == 2 && DM_PX(dm) && DM_PY(dm) && dm->M[0] == dm->M[1];
    if (o->pc_type != 0 && o->ts_type != 4) {
        snprintf(msg, sizeof msg, "%s.  Arbitrary callbacks run through the matrix-free route (stage operator = differenced "
                 "residual, no Jacobian matrix, hence no multigrid): pass -pc_type none", why);
        SHIM_ERR(56, msg);
    }
    if (o->ts_type == 4 && dm->ifunc)
        SHIM_ERR(56, "-ts_type rk integrates Ydot = G(t, Y): an IFunction is registered ([PETSc] TSRK refuses implicit systems too)");
    /* F(Y, a D) - F(Y, 0) must be a (F(Y, D) - F(Y, 0)) and must not depend on Y */
    double *Y = W->Y, *D = W->D, *F0 = W->Fu, *F1 = W->Fd;
    double *F2 = (double *)malloc(sizeof(double) * n), *T = (double *)malloc(sizeof(double) * n);
    if (!F2 || !T) { free(F2); free(T); SHIM_ERR(55, "out of host memory"); }
    PetscErrorCode rc = 0;
    double worst = 0.0, scale = 1.0;
    if (dm->ifunc) do {
        memset(T, 0, sizeof(double) * n);
        if ((rc = ts_eval_ifunc(dm, 0.0, Y, T, F0))) break;
        if ((rc = ts_eval_ifunc(dm, 0.0, Y, D, F1))) break;
        for (size_t i = 0; i < n; i++) T[i] = 2.0 * D[i];
        if ((rc = ts_eval_ifunc(dm, 0.0, Y, T, F2))) break;
        for (size_t i = 0; i < n; i++) {
            const double e = fabs((F2[i] - F0[i]) - 2.0 * (F1[i] - F0[i]));
            if (!(e <= worst)) worst = e;
            if (fabs(F1[i] - F0[i]) > scale) scale = fabs(F1[i] - F0[i]);
            F1[i] -= F0[i];                                           /* F1 := M D at Y */
        }
        for (size_t i = 0; i < n; i++) T[i] = Y[i] + 0.1 * D[i];
        memset(F2, 0, sizeof(double) * n);
        if ((rc = ts_eval_ifunc(dm, 0.0, T, F2, F0))) break;           /* F(Yb, 0) */
        if ((rc = ts_eval_ifunc(dm, 0.0, T, D, F2))) break;            /* F(Yb, D) */
        for (size_t i = 0; i < n; i++) {
            const double e = fabs((F2[i] - F0[i]) - F1[i]);
            if (!(e <= worst)) worst = e;
        }
    } while (0);
    free(F2);
    free(T);
    if (rc) return rc;
    if (!(worst <= 1.0e-10 * scale)) {
        snprintf(msg, sizeof msg, "TSSolve: the registered IFunction is not of the form M Ydot + f(Y) with a constant M "
                 "(deviation %.3e): the matrix-free route cannot difference it", worst);
        SHIM_ERR(56, msg);
    }
    ts_work_release(W);
    o->call_back_report = 0;
    o->no_rhsjacobian = 0;
    o->grid_x = dm->M0[0]; o->grid_y = dm->M0[1]; o->refine = dm->refine;
    o->ts_dt = ts->dt; o->ts_max_time = ts->max_time; o->ts_max_steps = ts->max_steps;
    o->ts_rtol = ts->rtol; o->ts_atol = ts->atol; o->ts_monitor = ts->monitor;
    o->snes_rtol = ts->snes->rtol; o->snes_stol = ts->snes->stol; o->snes_atol = ts->snes->atol; o->snes_max_it = ts->snes->max_it;
    o->ksp_rtol = ksp->rtol; o->ksp_max_it = ksp->max_it; o->gmres_restart = ts->snes->gmres_restart;
    o->snes_converged_reason = ts->snes->converged_reason_flag;
    o->ksp_converged_reason = ksp->converged_reason_flag;
    PetscCall(vec_to_host(x));
    p4b_pattern_result *R = W->R = (p4b_pattern_result *)calloc(1, sizeof *R);
    if (!R) SHIM_ERR(55, "out of host memory");
    fflush(stdout);
    int prc = pattern_shape ? p4b_ts2d_solve(g_ctx, o, ts_ifn, ts_gfn, ts, x->h, n, newton_line, NULL, R)
                            : p4b_ts_solve_callbacks(g_ctx, o, ts_ifn, ts_gfn, ts, x->h, n, newton_line, NULL, R);
    fflush(stdout);
    if (prc) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, prc, p4b_last_error());
    x->valid = LOC_HOST;
    g_ts_general++;
    return 0;
}

/* [PETSc] binary viewers on the TS monitors (c/ch5/MOVIES.md:44: -ts_monitor binary:t.dat -ts_monitor_solution binary:u.dat):
 * big-endian records, class id first -- Real 1211213 + one double per step into the first file, Vec 1211214 + length +
 * values into the second; what $PETSC_DIR/lib/petsc/bin/PetscBinaryIO.py (c/ch5/plotTS.py:44-46) reads. */
static struct { FILE *ft, *fu; } g_tsbin;
static void put_be32(FILE *f, int v) {
    unsigned char b[4] = {(unsigned char)((unsigned)v >> 24), (unsigned char)((unsigned)v >> 16), (unsigned char)((unsigned)v >> 8), (unsigned char)v};
    fwrite(b, 1, 4, f);
}
static void put_be64(FILE *f, const double *v, size_t n) {
    for (size_t i = 0; i < n; i++) {
        unsigned long long u;
        unsigned char b[8];
        memcpy(&u, v + i, 8);
        for (int k = 0; k < 8; k++) b[k] = (unsigned char)(u >> (56 - 8 * k));
        fwrite(b, 1, 8, f);
    }
}
static int ts_binary_monitor(void *user, int step, double t, const double *Y, size_t n) {
    (void)user; (void)step;
    if (g_tsbin.ft) { put_be32(g_tsbin.ft, 1211213); put_be64(g_tsbin.ft, &t, 1); }
    if (g_tsbin.fu) { put_be32(g_tsbin.fu, 1211214); put_be32(g_tsbin.fu, (int)n); put_be64(g_tsbin.fu, Y, n); }
    return 0;
}

/* Largest deviation, relative to the size of the values, between the registered IFunction / RHSFunction callbacks and the
 * device kernels of the identified model at (t, Y, Ydot) (host arrays, n doubles).  Own buffers: also used after the solve. */
static PetscErrorCode ts_model_deviation(DM dm, const p4b_pattern_opts *o, int m, double t, const double *Y, const double *D,
                                         size_t n, double *devF, double *devG) {
    double *Fu = (double *)malloc(sizeof(double) * n), *Fd = (double *)malloc(sizeof(double) * n);
    double *dY = NULL, *dD = NULL, *dF = NULL, scale;
    PetscErrorCode rc = 0;
    if (!Fu || !Fd) { free(Fu); free(Fd); SHIM_ERR(55, "out of host memory"); }
    if (p4b_malloc(g_ctx, n * sizeof(double), (void **)&dY) || p4b_malloc(g_ctx, n * sizeof(double), (void **)&dD) ||
        p4b_malloc(g_ctx, n * sizeof(double), (void **)&dF)) rc = 55;
    if (!rc) rc = p4b_memcpy_h2d(g_ctx, dY, Y, n * sizeof(double)) || p4b_memcpy_h2d(g_ctx, dD, D, n * sizeof(double));
    if (!rc) rc = ts_eval_ifunc(dm, t, Y, D, Fu);
    if (!rc) rc = p4b_pattern_ifunction(g_ctx, m, m, o->L, o->Du, o->Dv, dY, dD, dF) || p4b_memcpy_d2h(g_ctx, Fd, dF, n * sizeof(double));
    if (!rc) { *devF = maxabs_diff(Fu, Fd, n, &scale); *devF /= (scale > 1.0 ? scale : 1.0); }
    if (!rc) rc = ts_eval_rhs(dm, t, Y, Fu);
    if (!rc) rc = p4b_pattern_rhsfunction(g_ctx, m, m, o->phi, o->kappa, dY, dF) || p4b_memcpy_d2h(g_ctx, Fd, dF, n * sizeof(double));
    if (!rc) { *devG = maxabs_diff(Fu, Fd, n, &scale); *devG /= (scale > 1.0 ? scale : 1.0); }
    if (dY) p4b_free(g_ctx, dY);
    if (dD) p4b_free(g_ctx, dD);
    if (dF) p4b_free(g_ctx, dF);
    free(Fu); free(Fd);
    if (rc) SHIM_ERR(rc, "TSSolve: comparing the callbacks with the device kernels failed");
    return 0;
}

/* max |G_user(Y) - G_kernel(Y)| relative to max |G_user| (>= 1): is the registered RHSFunction the library's heat kernel? */
static PetscErrorCode ts_heat_deviation(DM dm, double D0, const double *Y, double *Fu, double *Fd, size_t n, double *dev) {
    double *dY = NULL, *dG = NULL, scale = 0.0;
    PetscCall(ts_eval_rhs(dm, 0.0, Y, Fu));
    P4B(p4b_malloc(g_ctx, n * sizeof(double), (void **)&dY));
    if (p4b_malloc(g_ctx, n * sizeof(double), (void **)&dG)) { p4b_free(g_ctx, dY); SHIM_ERR(55, "out of device memory"); }
    int rc = p4b_memcpy_h2d(g_ctx, dY, Y, n * sizeof(double));
    if (!rc) rc = p4b_heat_rhs(g_ctx, dm->M[0], dm->M[1], D0, dY, dG);
    if (!rc) rc = p4b_memcpy_d2h(g_ctx, Fd, dG, n * sizeof(double));
    p4b_free(g_ctx, dY);
    p4b_free(g_ctx, dG);
    if (rc) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, p4b_last_error());
    *dev = maxabs_diff(Fu, Fd, n, &scale);
    *dev /= scale > 1.0 ? scale : 1.0;
    if (!(*dev == *dev)) *dev = 1.0;
    return 0;
}
/* TSSolve on a 2-D DMDA that is not pattern.c's (any dof, any boundary types) -- c/ch5/heat.c: one component, Neumann in x,
 * periodic in y, RHSFunction + RHSJacobian, no IFunction.  The library has no kernels for such a system, so this is the
 * callback route from the start: G (and F, if registered) are the user's host callbacks on ghosted a[j][i] views, the
 * integrators and all vector algebra run on the device, implicit stages are solved matrix-free (Newton + GMRES on the
 * differenced residual: -pc_type none; the registered Jacobian callback is not called).  -ts_type rk is [PETSc]'s
 * default TSRK scheme 3bs with its adaptivity (c/ch5/output/heat.test2). */
static PetscErrorCode ts_solve_any_dmda(TS ts, Vec x, struct ts_work *W) {
    DM dm = ts->dm;
    KSP ksp = &ts->snes->ksp;
    PC pc = &ksp->pc;
    char msg[256];
    if (!dm->rhsfunc) SHIM_ERR(73, "TSSolve: call DMDATSSetRHSFunctionLocal()");
    p4b_pattern_opts o;
    P4B(p4b_pattern_default_opts(&o));
    if (!strcmp(ts->type, TSARKIMEX)) o.ts_type = 0;
    else if (!strcmp(ts->type, TSBEULER)) o.ts_type = 1;
    else if (!strcmp(ts->type, TSCN)) o.ts_type = 2;
    else if (!strcmp(ts->type, TSBDF)) o.ts_type = 3;
    else if (!strcmp(ts->type, TSRK)) o.ts_type = 4;
    else {
        snprintf(msg, sizeof msg, "-ts_type %s is not provided on the device path (arkimex, beuler, cn, bdf, rk are)", ts->type);
        SHIM_ERR(56, msg);
    }
    if (o.ts_type == 4 && opt_value("-ts_rk_type") && strcmp(opt_value("-ts_rk_type"), "3bs"))
        SHIM_ERR(56, "-ts_rk_type: 3bs ([PETSc]'s default) is the explicit scheme provided");
    if (ts->t0 != 0.0) SHIM_ERR(56, "TSSolve: the device path starts at t = 0 (heat.c:77)");
    if (ts->eft != TS_EXACTFINALTIME_MATCHSTEP)
        SHIM_ERR(56, "TSSolve: TS_EXACTFINALTIME_MATCHSTEP is what the device path provides (heat.c:80)");
    if (o.ts_type != 4) {
        if (!pc->type[0])
            SHIM_ERR(56, "PETSc's default PC (ILU(0) on one rank) is sequential and not provided on the device: pass -pc_type none "
                         "(the stage solves of this route are matrix-free)");
        if (!strcmp(pc->type, PCNONE)) o.pc_type = 0;
        else SHIM_ERR(56, "TSSolve on this DMDA: the stage operator is the differenced residual (no matrix): -pc_type none");
        if (strcmp(ksp->type, KSPGMRES)) SHIM_ERR(56, "TSSolve: the stage solves are GMRES ([PETSc] default)");
    } else o.pc_type = 0;
    PetscCall(ensure_ctx());
    const size_t n = x->n;
    PetscCall(vec_to_host(x));
    W->Y = (double *)calloc(n, sizeof(double)); W->D = (double *)calloc(n, sizeof(double));
    W->Fu = (double *)malloc(sizeof(double) * n); W->Fd = (double *)malloc(sizeof(double) * n);
    if (!W->Y || !W->D || !W->Fu || !W->Fd) SHIM_ERR(55, "out of host memory");
    unsigned long long lcg = 0x9E3779B97F4A7C15ULL;
    for (size_t i = 0; i < n; i++) {             /* a generic state and Ydot for the M Ydot + f(Y) check of an IFunction */
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        W->Y[i] = x->h[i] + 0.05 * ((double)(lcg >> 11) / 9007199254740992.0 - 0.5);
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        W->D[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
    }
    const double t_start = wall();
    /* heat.c's system is one the library does have as a kernel (p4b_heat_rhs): if the registered RHSFunction IS that function
     * -- D0 identified from one unit pulse, then compared at a generic state, and again at the final state -- the whole
     * run is device-resident (p4b_heat_solve: G and the stage operator are kernels).  -p4b_recognise_residual 0 switches
     * the substitution off. */
    {
        const char *v = opt_value("-p4b_recognise_residual");
        const int mx = dm->M[0], my = dm->M[1];
        double D0 = 0.0;
        int is_heat = (!v || atol(v) != 0) && dm->dof == 1 && !DM_PX(dm) && DM_PY(dm) && !dm->ifunc && mx >= 3 && my >= 3;
        if (is_heat) {
            const size_t pn = (size_t)1 * mx + 1;                 /* node (1, 1): an interior row of the Laplacian */
            const double hx = 1.0 / (mx - 1), hy = 1.0 / my;
            memset(W->Fd, 0, sizeof(double) * n);
            PetscCall(ts_eval_rhs(dm, 0.0, W->Fd, W->Fu));        /* G(0) */
            const double g0 = W->Fu[pn];
            W->Fd[pn] = 1.0;
            PetscCall(ts_eval_rhs(dm, 0.0, W->Fd, W->Fu));        /* G(e_pn) */
            D0 = -(W->Fu[pn] - g0) / (2.0 / (hx * hx) + 2.0 / (hy * hy));
            is_heat = D0 > 0.0 && D0 == D0;
        }
        if (is_heat) {
            double dev = 0.0;
            PetscCall(ts_heat_deviation(dm, D0, W->Y, W->Fu, W->Fd, n, &dev));
            is_heat = dev <= 1.0e-11;
        }
        if (is_heat) {
            o.grid_x = dm->M0[0]; o.grid_y = dm->M0[1]; o.refine = dm->refine;
            o.ts_dt = ts->dt; o.ts_max_time = ts->max_time; o.ts_max_steps = ts->max_steps;
            o.ts_rtol = ts->rtol; o.ts_atol = ts->atol; o.ts_monitor = ts->monitor;
            o.snes_rtol = ts->snes->rtol; o.snes_stol = ts->snes->stol; o.snes_atol = ts->snes->atol; o.snes_max_it = ts->snes->max_it;
            o.ksp_rtol = ksp->rtol; o.ksp_max_it = ksp->max_it; o.gmres_restart = ts->snes->gmres_restart;
            o.snes_converged_reason = ts->snes->converged_reason_flag;
            o.ksp_converged_reason = ksp->converged_reason_flag;
            p4b_pattern_result *R = (p4b_pattern_result *)calloc(1, sizeof *R);
            if (!R) SHIM_ERR(55, "out of host memory");
            fflush(stdout);
            int prc = p4b_heat_solve(g_ctx, &o, mx, my, D0, x->h, newton_line, NULL, R);
            const double tf = R->t_final;
            free(R);
            fflush(stdout);
            if (prc) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, prc, p4b_last_error());
            x->valid = LOC_HOST;
            /* once more where the solve ended up (as for pattern.c: the trajectory has been printed, so a mismatch is an
             * error, not a silent re-run) */
            double dev = 1.0;
            PetscCall(ts_heat_deviation(dm, D0, x->h, W->Fu, W->Fd, n, &dev));
            if (!(dev <= 1.0e-10)) {
                char m2[512];
                snprintf(m2, sizeof m2, "TSSolve: the registered RHSFunction matched the library's heat kernel at the probes but "
                         "not at the final state (t = %g, deviation %.3e): the trajectory above is the KERNEL's, not the "
                         "callback's; run again with -p4b_recognise_residual 0 (host callback)", tf, dev);
                SHIM_ERR(56, m2);
            }
            g_t_snes += wall() - t_start;
            fprintf(stderr, "[p4b200] TS: the registered FormRHSFunctionLocal equals the library's heat-equation kernel (D0 = %g; "
                            "probed at a generic state, re-verified at the final state): time stepping on the device.  "
                            "-p4b_recognise_residual 0 keeps it a host callback.\n", D0);
            return 0;
        }
    }
    PetscErrorCode rc = ts_solve_general(ts, x, W, &o, "this DMDA is not pattern.c's");
    g_t_snes += wall() - t_start;
    if (!rc)
        fprintf(stderr, "[p4b200] TS: callbacks evaluated on the host (no kernels for this system), integrator and vector algebra "
                        "on the device%s.\n", o.ts_type == 4 ? "" : ", matrix-free stage operator");
    return rc;
}

static PetscErrorCode ts_solve(TS ts, Vec x, struct ts_work *W) {
    DM dm = ts->dm;
    KSP ksp = &ts->snes->ksp;
    PC pc = &ksp->pc;
    char msg[512];
    if (!dm) SHIM_ERR(73, "TSSolve: call TSSetDM() first");
    if (dm->dim != 2) SHIM_ERR(56, "TSSolve is provided for 2-D DMDAs (pattern.c:79-84, heat.c:60-65)");
    if (dm->dof != 2 || !DM_PX(dm) || !DM_PY(dm)) return ts_solve_any_dmda(ts, x, W);
    if (dm->M[0] != dm->M[1]) SHIM_ERR(56, "TSSolve: the device path needs mx == my (pattern.c:89)");
    if (!dm->ifunc || !dm->rhsfunc) SHIM_ERR(73, "TSSolve: call DMDATSSetIFunctionLocal() and DMDATSSetRHSFunctionLocal()");
    if (!dm->ijac && !ts->snes->fd_color)
        SHIM_ERR(56, "TSSolve: no IJacobian callback registered (-ptn_no_ijacobian): the device path checks its stage matrix "
                     "against the registered callback; pass -snes_fd_color to say that differencing the residual is meant");
    p4b_pattern_opts o;
    P4B(p4b_pattern_default_opts(&o));
    if (!strcmp(ts->type, TSARKIMEX)) o.ts_type = 0;
    else if (!strcmp(ts->type, TSBEULER)) o.ts_type = 1;
    else if (!strcmp(ts->type, TSCN)) o.ts_type = 2;
    else if (!strcmp(ts->type, TSBDF)) o.ts_type = 3;
    else {
        snprintf(msg, sizeof msg, "-ts_type %s is not provided on the device path (arkimex, beuler, cn, bdf are)", ts->type);
        SHIM_ERR(56, msg);
    }
    if (ts->t0 != 0.0) SHIM_ERR(56, "TSSolve: the device path starts at t = 0 (pattern.c:116)");
    if (ts->eft != TS_EXACTFINALTIME_MATCHSTEP)
        SHIM_ERR(56, "TSSolve: TS_EXACTFINALTIME_MATCHSTEP is what the device path provides (pattern.c:119)");
    if (!pc->type[0])
        SHIM_ERR(56, "PETSc's default PC (ILU(0) on one rank) is sequential and not provided on the device: "
                     "pass -pc_type mg or -pc_type none");
    if (!strcmp(pc->type, PCMG)) o.pc_type = 1;
    else if (!strcmp(pc->type, PCNONE)) o.pc_type = 0;
    else SHIM_ERR(56, "TSSolve: -pc_type mg and -pc_type none are provided");
    if (o.pc_type == 1 && strcmp(pc->levels_pc, "jacobi"))
        SHIM_ERR(56, "PCMG's default level smoother PC (SOR) is sequential and not provided on the device: "
                     "pass -mg_levels_pc_type jacobi");
    if (strcmp(ksp->type, KSPGMRES)) SHIM_ERR(56, "TSSolve: the stage solves are GMRES ([PETSc] default)");
    PetscCall(ensure_ctx());
    {
        const char *v = opt_value("-p4b_gmres_cgs");
        P4B(p4b_tune("gmres_cgs", v ? atol(v) : 0));
    }
    const double t_start = wall();

    const int m = dm->M[0];
    const size_t n = x->n;
    const double Lbox = dm->cmax[0] - dm->cmin[0];
    if (fabs((dm->cmax[1] - dm->cmin[1]) - Lbox) > 1e-14 * fabs(Lbox) || !(Lbox > 0.0))
        SHIM_ERR(56, "TSSolve: the device path needs a square box (DMDASetUniformCoordinates, pattern.c:92)");
    const double h = Lbox / m;
    PetscCall(vec_to_host(x));
    double *Y = W->Y = (double *)calloc(n, sizeof(double)), *D = W->D = (double *)calloc(n, sizeof(double));
    double *Fu = W->Fu = (double *)malloc(sizeof(double) * n), *Fd = W->Fd = (double *)malloc(sizeof(double) * n);
    if (!Y || !D || !Fu || !Fd) SHIM_ERR(55, "out of host memory");

    /* (1) identify: G at (u,v) = (0,0) is (phi, 0), at (0,1) it is (phi, -(phi+kappa)); F of unit pulses with Ydot = 0
     *     is 20 C_c at the pulse, C_c = D_c / (6 h^2) */
    Y[2 * 1 + 1] = 1.0;                                          /* node 1: (u, v) = (0, 1); every other node (0, 0) */
    PetscCall(ts_eval_rhs(dm, 0.0, Y, Fu));
    o.phi = Fu[0];
    o.kappa = -Fu[2 * 1 + 1] - o.phi;
    Y[2 * 1 + 1] = 0.0;
    const size_t pn = (size_t)1 * m + 1;                         /* node (1, 1) */
    Y[2 * pn] = 1.0; Y[2 * pn + 1] = 1.0;
    PetscCall(ts_eval_ifunc(dm, 0.0, Y, D, Fu));
    o.Du = Fu[2 * pn] / 20.0 * 6.0 * h * h;
    o.Dv = Fu[2 * pn + 1] / 20.0 * 6.0 * h * h;
    o.L = Lbox;

    /* (2) verify at a generic state: the caller's initial state plus a fixed pseudo-random perturbation */
    unsigned long long lcg = 0x9E3779B97F4A7C15ULL;
    for (size_t i = 0; i < n; i++) {
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        Y[i] = x->h[i] + 0.05 * ((double)(lcg >> 11) / 9007199254740992.0 - 0.5);
        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
        D[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
    }
    P4B(p4b_malloc(g_ctx, n * sizeof(double), (void **)&W->dY));
    P4B(p4b_malloc(g_ctx, n * sizeof(double), (void **)&W->dD));
    P4B(p4b_malloc(g_ctx, n * sizeof(double), (void **)&W->dF));
    double *dY = W->dY, *dD = W->dD, *dF = W->dF;
    P4B(p4b_memcpy_h2d(g_ctx, dY, Y, n * sizeof(double)));
    P4B(p4b_memcpy_h2d(g_ctx, dD, D, n * sizeof(double)));
    double dev, scale;
    PetscCall(ts_eval_ifunc(dm, 0.0, Y, D, Fu));
    P4B(p4b_pattern_ifunction(g_ctx, m, m, o.L, o.Du, o.Dv, dY, dD, dF));
    P4B(p4b_memcpy_d2h(g_ctx, Fd, dF, n * sizeof(double)));
    int is_model = 1;
    dev = maxabs_diff(Fu, Fd, n, &scale);
    if (!(dev <= 1.0e-11 * (scale > 1.0 ? scale : 1.0))) {
        is_model = 0;
        snprintf(msg, sizeof msg, "TSSolve: the registered IFunction is not F = Ydot - D L9 Y / (6 h^2) with the identified "
                 "D_u = %g, D_v = %g (max deviation %.3e): not the reaction-diffusion model the device path has as kernels", o.Du,
                 o.Dv, dev);
    }
    if (is_model) {
        PetscCall(ts_eval_rhs(dm, 0.0, Y, Fu));
        P4B(p4b_pattern_rhsfunction(g_ctx, m, m, o.phi, o.kappa, dY, dF));
        P4B(p4b_memcpy_d2h(g_ctx, Fd, dF, n * sizeof(double)));
        dev = maxabs_diff(Fu, Fd, n, &scale);
        if (!(dev <= 1.0e-11 * (scale > 1.0 ? scale : 1.0))) {
            is_model = 0;
            snprintf(msg, sizeof msg, "TSSolve: the registered RHSFunction is not G = (-u v^2 + phi (1-u), u v^2 - (phi+kappa) v) "
                     "with the identified phi = %g, kappa = %g (max deviation %.3e): not the model the device path has as kernels",
                     o.phi, o.kappa, dev);
        }
    }
    /* (2b) the same at later times and at states far from the initial one: a forcing that depends on t, or a term that
     *      only acts for larger values, must not slip through one probe at t = 0 */
    for (int probe = 1; probe <= 2 && is_model; probe++) {
        const double tp = probe == 1 ? 0.37 * ts->max_time : ts->max_time, amp = probe == 1 ? 0.6 : 1.5;
        for (size_t i = 0; i < n; i++) {
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            Y[i] = (probe == 1 ? x->h[i] : 0.0) + amp * ((double)(lcg >> 11) / 9007199254740992.0 - (probe == 1 ? 0.5 : 0.0));
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            D[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5;
        }
        double dF = 0.0, dG = 0.0;
        PetscCall(ts_model_deviation(dm, &o, m, tp, Y, D, n, &dF, &dG));
        if (!(dF <= 1.0e-11) || !(dG <= 1.0e-11)) {
            is_model = 0;
            snprintf(msg, sizeof msg, "TSSolve: the registered callbacks equal the reaction-diffusion model at t = 0 near the "
                     "initial state but not at t = %g, amplitude %g (deviation F %.3e, G %.3e): not the model the device path "
                     "has as kernels", tp, amp, dF, dG);
        }
    }
    if (is_model && opt_value("-p4b_recognise_residual") && !atol(opt_value("-p4b_recognise_residual"))) {
        is_model = 0;                  /* A/B: run the model through the general route as well */
        snprintf(msg, sizeof msg, "TSSolve: -p4b_recognise_residual 0");
    }
    if (!is_model) return ts_solve_general(ts, x, W, &o, msg);
    /* the Jacobian callbacks PETSc would call for this type: IJacobian always, RHSJacobian for the fully implicit types */
    {
        DMDALocalInfo info;
        struct rd_check chk;
        struct _p_Mat P;
        const double t_jac0 = wall();
        PetscCall(DMDAGetLocalInfo(dm, &info));
        if (ghosted_make(Y, m, m, 2, 1, 1, &W->gY) || ghosted_make(D, m, m, 2, 1, 1, &W->gD)) SHIM_ERR(55, "out of host memory");
        memset(&chk, 0, sizeof chk);
        chk.m = m; chk.shift = 1.0 / ts->dt; chk.C[0] = o.Du / (6.0 * h * h); chk.C[1] = o.Dv / (6.0 * h * h);
        chk.phi = o.phi; chk.kappa = o.kappa; chk.Y = Y;
        memset(&P, 0, sizeof P);
        P.dm = dm; P.type = MATSTENCILCUDA; P.rd = &chk;
        for (int mode = 1; mode <= 2; mode++) {
            if (ts->snes->fd_color) break;                       /* PETSc would not call them either */
            if (mode == 2 && (!dm->rhsjac || o.ts_type == 0)) break;
            chk.mode = mode; chk.maxdev = 0.0; chk.rows = 0; chk.bad_structure = 0;
            PetscErrorCode rc = mode == 1 ? dm->ijac(&info, 0.0, W->gY.a, W->gD.a, chk.shift, &P, &P, dm->ijacctx)
                                          : dm->rhsjac(&info, 0.0, W->gY.a, &P, &P, dm->rhsjacctx);
            if (rc) return rc;
            const double tol = 1.0e-11 * (fabs(chk.shift) + 20.0 * chk.C[0] + 20.0 * chk.C[1] + 1.0);
            if (chk.bad_structure || chk.rows != (long long)n || !(chk.maxdev <= tol)) {
                snprintf(msg, sizeof msg, "TSSolve: the registered %s does not insert the stage matrix the device path applies "
                         "(%lld of %zu rows, max deviation %.3e%s)", mode == 1 ? "IJacobian" : "RHSJacobian", chk.rows, n,
                         chk.maxdev, chk.bad_structure ? ", entries outside the expected pattern" : "");
                SHIM_ERR(56, msg);
            }
        }
        g_t_jac += wall() - t_jac0;
    }

    /* (3) the solve, on the device, from the caller's state */
    o.no_rhsjacobian = dm->rhsjac ? 0 : 1;
    o.call_back_report = 0;
    o.grid_x = dm->M0[0]; o.grid_y = dm->M0[1]; o.refine = dm->refine;
    o.ts_dt = ts->dt; o.ts_max_time = ts->max_time; o.ts_max_steps = ts->max_steps;
    o.ts_rtol = ts->rtol; o.ts_atol = ts->atol; o.ts_monitor = ts->monitor;
    o.smooth_its = pc->smooth_its;
    {
        const char *v = opt_value("-p4b_mg_rscale");
        if (v) o.mg_rscale = strtod(v, NULL);
    }
    o.snes_rtol = ts->snes->rtol; o.snes_stol = ts->snes->stol; o.snes_atol = ts->snes->atol; o.snes_max_it = ts->snes->max_it;
    o.ksp_rtol = ksp->rtol; o.ksp_max_it = ksp->max_it; o.gmres_restart = ts->snes->gmres_restart;
    o.snes_converged_reason = ts->snes->converged_reason_flag;
    o.ksp_converged_reason = ksp->converged_reason_flag;
    PetscCall(vec_to_dev(x));
    ts_work_release(W);                          /* the probes are done: give their buffers back before the solve */
    p4b_pattern_result *R = W->R = (p4b_pattern_result *)calloc(1, sizeof *R);
    if (!R) SHIM_ERR(55, "out of host memory");
    fflush(stdout);
    fprintf(stderr, "[p4b200] TS: the registered callbacks equal the library's reaction-diffusion model (F, G at three "
                    "times and states, every Jacobian row; re-verified at the final state): time stepping on the device.  "
                    "-p4b_recognise_residual 0 -pc_type none runs the host callbacks.\n");
    int ngpu = 1;
    {
        const char *v = opt_value("-p4b_gpus");
        if (!v) v = getenv("P4B_GPUS");
        if (v) ngpu = atoi(v);
        if (ngpu < 1) SHIM_ERR(62, "-p4b_gpus must be at least 1");
        if (ngpu > 1 && (g_tsbin.ft || g_tsbin.fu))
            SHIM_ERR(56, "-ts_monitor[_solution] binary: with -p4b_gpus > 1 is not provided (every GPU holds its rows only)");
    }
    if (ngpu > 1) {
        PetscCall(vec_to_host(x));
        fflush(stdout);
        PetscCall(ts_solve_multi_gpu(ngpu, m, &o, x->h, R));
        fflush(stdout);
        x->valid = LOC_HOST;
    } else {
    int rc = p4b_pattern_solve_from(g_ctx, &o, x->d, newton_line, NULL, x->d, n, R);
    fflush(stdout);
    if (rc) return PetscShimError(PETSC_COMM_SELF, __LINE__, __func__, __FILE__, rc, p4b_last_error());
    x->valid = LOC_DEV;
    }
    {   /* (4) once more where the solve ended up: the final state at the final time.  The probes are a finite sample;
         *     what they could not see is reported loudly here (the trajectory has been printed already, so this is an
         *     error, not a silent re-run) */
        double *Yf = (double *)malloc(sizeof(double) * n), *Df = (double *)malloc(sizeof(double) * n), dF = 0.0, dG = 0.0;
        if (!Yf || !Df) { free(Yf); free(Df); SHIM_ERR(55, "out of host memory"); }
        PetscCall(vec_to_host(x));
        unsigned long long l2 = 0xD1B54A32D192ED03ULL;
        for (size_t i = 0; i < n; i++) {
            Yf[i] = x->h[i];
            l2 = l2 * 6364136223846793005ULL + 1442695040888963407ULL;
            Df[i] = (double)(l2 >> 11) / 9007199254740992.0 - 0.5;
        }
        PetscErrorCode rv = ts_model_deviation(dm, &o, m, R->t_final, Yf, Df, n, &dF, &dG);
        free(Yf); free(Df);
        if (rv) return rv;
        if (!(dF <= 1.0e-10) || !(dG <= 1.0e-10)) {
            snprintf(msg, sizeof msg, "TSSolve: the registered callbacks matched the device kernels at every probe but not at "
                     "the final state (t = %g, deviation F %.3e, G %.3e): the trajectory above is the MODEL's, not the "
                     "callbacks'; run again with -p4b_recognise_residual 0 -pc_type none (host callbacks)", R->t_final, dF, dG);
            SHIM_ERR(56, msg);
        }
    }
    g_t_snes += wall() - t_start;
    return 0;
}

/* [PETSc] TSMonitorSet (heat.c:72): user monitors run at step 0 and after every accepted step with the step number, the time
 * and the solution; inside one, TSGetDM / TSGetTime / TSGetTimeStep answer for the solve in progress (heat.c:117,133). */
PetscErrorCode TSMonitorSet(TS ts, PetscErrorCode (*monitor)(TS, PetscInt, PetscReal, Vec, void *), void *mctx,
                            PetscErrorCode (*mdestroy)(void **)) {
    if (ts->nmon >= 5) SHIM_ERR(63, "too many monitors set");
    ts->mon[ts->nmon].f = monitor;
    ts->mon[ts->nmon].ctx = mctx;
    ts->mon[ts->nmon].destroy = mdestroy;
    ts->nmon++;
    return 0;
}
PetscErrorCode TSGetTime(TS ts, PetscReal *t) { *t = ts->in_solve ? ts->tcur : ts->t0; return 0; }
PetscErrorCode TSGetMaxTime(TS ts, PetscReal *maxtime) { *maxtime = ts->max_time; return 0; }
PetscErrorCode TSGetTimeStep(TS ts, PetscReal *dt) { *dt = ts->in_solve ? p4b_ts_time_step() : ts->dt; return 0; }
PetscErrorCode TSGetDM(TS ts, DM *dm) { *dm = ts->dm; return 0; }

static int ts_step_dispatch(void *user, int step, double t, const double *Y, size_t n) {
    TS ts = (TS)user;
    if ((g_tsbin.ft || g_tsbin.fu) && ts_binary_monitor(NULL, step, t, Y, n)) return 1;
    if (!ts->nmon) return 0;
    struct _p_Vec v;
    memset(&v, 0, sizeof v);
    v.n = n; v.dm = ts->dm; v.h = (double *)Y; v.valid = LOC_HOST;
    ts->tcur = t;
    PetscErrorCode rc = 0;
    for (int i = 0; i < ts->nmon && !rc; i++) rc = ts->mon[i].f(ts, step, t, &v, ts->mon[i].ctx);
    free(v.tables);
    fflush(stdout);
    return rc != 0;
}

PetscErrorCode TSSolve(TS ts, Vec x) {
    struct ts_work W;
    memset(&W, 0, sizeof W);
    {
        const char *ft = binary_target("-ts_monitor"), *fu = binary_target("-ts_monitor_solution");
        if (opt_value("-ts_monitor_solution") && !fu)
            SHIM_ERR(56, "-ts_monitor_solution: the binary viewer is provided (binary:FILE); draw / ascii viewers are not");
        g_tsbin.ft = ft ? fopen(ft, "wb") : NULL;
        g_tsbin.fu = fu ? fopen(fu, "wb") : NULL;
        if ((ft && !g_tsbin.ft) || (fu && !g_tsbin.fu)) SHIM_ERR(65, "cannot open the binary viewer's file for writing");
        if (g_tsbin.ft || g_tsbin.fu || ts->nmon) P4B(p4b_set_ts_step_monitor(ts_step_dispatch, ts));
    }
    ts->in_solve = 1;
    PetscErrorCode rc = ts_solve(ts, x, &W);
    ts->in_solve = 0;
    if (g_tsbin.ft || g_tsbin.fu || ts->nmon) {
        p4b_set_ts_step_monitor(NULL, NULL);
        if (g_tsbin.ft) fclose(g_tsbin.ft);
        if (g_tsbin.fu) fclose(g_tsbin.fu);
        g_tsbin.ft = g_tsbin.fu = NULL;
    }
    ts_work_release(&W);                         /* also on every error return of ts_solve */
    return rc;
}
