/* refwrap.c -- includes the reference's c/ch6/fish.c VERBATIM (path given by -DREF_FISH_C=...) so that its
 * static function tables (g_bdry_ptr, f_rhs_ptr: fish.c:115-123) are visible, and exposes the reference's
 * own callbacks through a plain C interface for the oracle tests.  TEST INFRASTRUCTURE ONLY. */
#define main fish_reference_main
#include REF_FISH_C
#undef main
#include <stdlib.h>

DM refstub_dm(int dim, const int *M, const double *L);
Vec refstub_vec(DM dm, double *data);
Mat refstub_mat(DM dm);
long refstub_mat_nnz(Mat A);
void refstub_mat_copy(Mat A, int *row, int *col, double *val);
void refstub_mat_free(Mat A);
void *refstub_tables(int dim, const int *M, double *base);

static void make_user(PoissonCtx *user, int dim, int problem, const double *L, const double *c) {
    user->Lx = L[0]; user->Ly = L[1]; user->Lz = L[2];
    user->cx = c[0]; user->cy = c[1]; user->cz = c[2];
    user->g_bdry = g_bdry_ptr[dim - 1][problem];
    user->f_rhs = f_rhs_ptr[dim - 1][problem];
    user->addctx = NULL;
}

/* F = Poisson{1,2,3}DFunctionLocal(u) of the reference */
int ref_function(int dim, const int *M, const double *L, const double *c, int problem, double *u, double *F) {
    PoissonCtx user;
    DMDALocalInfo info;
    make_user(&user, dim, problem, L, c);
    DM dm = refstub_dm(dim, M, L);
    DMDAGetLocalInfo(dm, &info);
    void *au = refstub_tables(dim, info.dim == 1 ? M : M, u), *aF = refstub_tables(dim, M, F);
    int rc = residual_ptr[dim - 1](&info, au, aF, &user);
    if (dim > 1) { free(au); free(aF); }
    free(dm);
    return rc;
}

/* COO triplets inserted by Poisson{1,2,3}DJacobianLocal of the reference; returns nnz (or -1) */
long ref_jacobian(int dim, const int *M, const double *L, const double *c, long cap, int *row, int *col, double *val) {
    PoissonCtx user;
    DMDALocalInfo info;
    make_user(&user, dim, 0, L, c);
    DM dm = refstub_dm(dim, M, L);
    DMDAGetLocalInfo(dm, &info);
    Mat J = refstub_mat(dm);
    int rc = jacobian_ptr[dim - 1](&info, NULL, J, J, &user);
    long n = refstub_mat_nnz(J);
    if (rc || n > cap) n = -1;
    else refstub_mat_copy(J, row, col, val);
    refstub_mat_free(J);
    free(dm);
    return n;
}

/* InitialState (poissonfunctions.c:260-346) with ZEROS */
int ref_initial_state(int dim, const int *M, const double *L, int problem, int gonboundary, double *u) {
    PoissonCtx user;
    const double c[3] = {1, 1, 1};
    make_user(&user, dim, problem, L, c);
    DM dm = refstub_dm(dim, M, L);
    Vec v = refstub_vec(dm, u);
    int rc = InitialState(dm, ZEROS, gonboundary ? PETSC_TRUE : PETSC_FALSE, v, &user);
    free(v);
    free(dm);
    return rc;
}

/* Form{1,2,3}DUExact (fish.c:288-340) */
int ref_uexact(int dim, const int *M, const double *L, int problem, double *u) {
    PoissonCtx user;
    DMDALocalInfo info;
    const double c[3] = {1, 1, 1};
    make_user(&user, dim, problem, L, c);
    DM dm = refstub_dm(dim, M, L);
    DMDAGetLocalInfo(dm, &info);
    Vec v = refstub_vec(dm, u);
    int rc = getuexact_ptr[dim - 1](&info, v, &user);
    free(v);
    free(dm);
    return rc;
}
