/* stub.c -- host-only, minimal implementations of the few PETSc calls that the reference's CALLBACKS make
 * (c/ch6/poissonfunctions.c), so that file and c/ch6/fish.c can be compiled unchanged, from where they lie
 * under /root/reference, into oracle/_ref/libfishref.so.  TEST INFRASTRUCTURE: it validates the oracle's
 * restatement of the discretisation against the reference's own code.  It contains no solver. */
#define _POSIX_C_SOURCE 200809L
#include <petsc.h>
#include <stdarg.h>
#include <stdlib.h>

struct _p_DM { int dim, M[3], dof; double cmin[3], cmax[3]; };
struct _p_Vec { size_t n; double *h; DM dm; void *tables; };
struct _p_Mat { long cap, n; int *row, *col; double *val; DM dm; };

PetscErrorCode PetscShimError(MPI_Comm comm, int line, const char *func, const char *file, PetscErrorCode code,
                              const char *msg) {
    (void)comm;
    fprintf(stderr, "[refstub] %s (%s:%d %s)\n", msg, file, line, func);
    return code ? code : 1;
}
PetscErrorCode PetscLogFlops(PetscLogDouble f) { (void)f; return 0; }
PetscErrorCode DMGetBoundingBox(DM dm, PetscReal gmin[], PetscReal gmax[]) {
    for (int i = 0; i < dm->dim; i++) { gmin[i] = dm->cmin[i]; gmax[i] = dm->cmax[i]; }
    return 0;
}
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info) {
    memset(info, 0, sizeof *info);
    info->da = da; info->dim = da->dim; info->dof = 1; info->sw = 1;
    info->mx = da->M[0]; info->my = da->M[1]; info->mz = da->M[2];
    info->xm = da->M[0]; info->ym = da->M[1]; info->zm = da->M[2];
    info->gxm = da->M[0]; info->gym = da->M[1]; info->gzm = da->M[2];
    return 0;
}
static void *tables(int dim, const int *M, double *base) {
    if (dim == 1) return base;
    if (dim == 2) {
        double **rows = malloc(sizeof(double *) * M[1]);
        for (int j = 0; j < M[1]; j++) rows[j] = base + (size_t)j * M[0];
        return rows;
    }
    char *blk = malloc(sizeof(double **) * M[2] + sizeof(double *) * (size_t)M[2] * M[1]);
    double ***pl = (double ***)blk;
    double **rows = (double **)(blk + sizeof(double **) * M[2]);
    for (int k = 0; k < M[2]; k++) {
        pl[k] = rows + (size_t)k * M[1];
        for (int j = 0; j < M[1]; j++) pl[k][j] = base + ((size_t)k * M[1] + j) * M[0];
    }
    return pl;
}
PetscErrorCode DMDAVecGetArray(DM da, Vec v, void *array) {
    v->tables = tables(da->dim, da->M, v->h);
    *(void **)array = v->tables;
    return 0;
}
PetscErrorCode DMDAVecRestoreArray(DM da, Vec v, void *array) {
    if (da->dim > 1) free(v->tables);
    v->tables = NULL;
    *(void **)array = NULL;
    return 0;
}
PetscErrorCode VecSet(Vec x, PetscScalar a) { for (size_t i = 0; i < x->n; i++) x->h[i] = a; return 0; }
PetscErrorCode VecSetRandom(Vec x, PetscRandom r) { (void)x; (void)r; return 56; }
PetscErrorCode PetscRandomCreate(MPI_Comm c, PetscRandom *r) { (void)c; *r = NULL; return 0; }
PetscErrorCode PetscRandomDestroy(PetscRandom *r) { *r = NULL; return 0; }
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil im[], PetscInt n, const MatStencil in[],
                                   const PetscScalar v[], InsertMode addv) {
    (void)addv;
    const int *M = A->dm->M, dim = A->dm->dim;
    for (int r = 0; r < m; r++)
        for (int c = 0; c < n; c++) {
            if (A->n == A->cap) {
                A->cap = A->cap ? 2 * A->cap : 1024;
                A->row = realloc(A->row, sizeof(int) * A->cap);
                A->col = realloc(A->col, sizeof(int) * A->cap);
                A->val = realloc(A->val, sizeof(double) * A->cap);
            }
            const int rk = dim >= 3 ? im[r].k : 0, rj = dim >= 2 ? im[r].j : 0;
            const int ck = dim >= 3 ? in[c].k : 0, cj = dim >= 2 ? in[c].j : 0;
            const int dof = A->dm->dof > 1 ? A->dm->dof : 1;
            if (dof > 1) {      /* multi-component periodic grids (pattern.c): wrap the stencil indices */
                const int ri = (im[r].i + M[0]) % M[0], rjj = (rj + M[1]) % M[1];
                const int ci = (in[c].i + M[0]) % M[0], cjj = (cj + M[1]) % M[1];
                A->row[A->n] = (rjj * M[0] + ri) * dof + im[r].c;
                A->col[A->n] = (cjj * M[0] + ci) * dof + in[c].c;
                A->val[A->n] = v[r * n + c];
                A->n++;
                continue;
            }
            A->row[A->n] = (rk * M[1] + rj) * M[0] + im[r].i;
            A->col[A->n] = (ck * M[1] + cj) * M[0] + in[c].i;
            A->val[A->n] = v[r * n + c];
            A->n++;
        }
    return 0;
}
PetscErrorCode MatZeroEntries(Mat A) { A->n = 0; return 0; }
PetscErrorCode VecScale(Vec x, PetscScalar a) { for (size_t i = 0; i < x->n; i++) x->h[i] *= a; return 0; }
void refstub_set_dof(DM dm, int dof) { dm->dof = dof; }
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }

/* helpers for refwrap.c */
DM refstub_dm(int dim, const int *M, const double *L) {
    DM d = calloc(1, sizeof *d);
    d->dim = dim;
    for (int i = 0; i < 3; i++) { d->M[i] = i < dim ? M[i] : 1; d->cmin[i] = 0; d->cmax[i] = L[i]; }
    return d;
}
Vec refstub_vec(DM dm, double *data) {
    Vec v = calloc(1, sizeof *v);
    v->n = (size_t)dm->M[0] * dm->M[1] * dm->M[2];
    v->h = data;
    v->dm = dm;
    return v;
}
Mat refstub_mat(DM dm) { Mat A = calloc(1, sizeof *A); A->dm = dm; return A; }
long refstub_mat_nnz(Mat A) { return A->n; }
void refstub_mat_copy(Mat A, int *row, int *col, double *val) {
    memcpy(row, A->row, sizeof(int) * A->n);
    memcpy(col, A->col, sizeof(int) * A->n);
    memcpy(val, A->val, sizeof(double) * A->n);
}
void refstub_mat_free(Mat A) { free(A->row); free(A->col); free(A->val); free(A); }
void *refstub_tables(int dim, const int *M, double *base) { return tables(dim, M, base); }
