/* refwrap_pattern.c -- includes the reference's c/ch5/pattern.c VERBATIM (main renamed; its InitialState renamed so
 * it can share a library with fish's) and exposes its callbacks for the oracle tests.  TEST INFRASTRUCTURE ONLY. */
#define main pattern_reference_main
#define InitialState pattern_InitialState
#include REF_PATTERN_C
#undef main
#undef InitialState
#include <stdlib.h>

DM refstub_dm(int dim, const int *M, const double *L);
long refstub_mat_nnz(Mat A);
Mat refstub_mat(DM dm);
void refstub_mat_copy(Mat A, int *row, int *col, double *val);
void refstub_mat_free(Mat A);
void refstub_set_dof(DM dm, int dof);

/* periodic ghosted views: aY[j][i] valid for i,j in [-1, m] (pattern.c:252-257 reads them) */
static Field **ghosted(int mx, int my, const double *Y, Field **store) {
    Field *buf = malloc(sizeof(Field) * (size_t)(mx + 2) * (my + 2));
    Field **rows = malloc(sizeof(Field *) * (my + 2));
    for (int j = -1; j <= my; j++) {
        rows[j + 1] = buf + (size_t)(j + 1) * (mx + 2) + 1;
        const int jj = (j + my) % my;
        for (int i = -1; i <= mx; i++) {
            const int ii = (i + mx) % mx;
            rows[j + 1][i].u = Y[2 * (jj * mx + ii)];
            rows[j + 1][i].v = Y[2 * (jj * mx + ii) + 1];
        }
    }
    *store = buf;
    return rows + 1;
}

static void make_ctx(PatternCtx *user, double Lside, double Du, double Dv, double phi, double kappa) {
    user->L = Lside; user->Du = Du; user->Dv = Dv; user->phi = phi; user->kappa = kappa;
}
static void make_info(DMDALocalInfo *info, int mx, int my) {
    memset(info, 0, sizeof *info);
    info->dim = 2; info->dof = 2; info->sw = 1;
    info->mx = mx; info->my = my; info->mz = 1;
    info->xm = mx; info->ym = my; info->zm = 1;
}

int ref_pattern_ifunction(int mx, int my, double Lside, double Du, double Dv, const double *Y, const double *Ydot, double *F) {
    PatternCtx user;
    DMDALocalInfo info;
    Field *b1, *b2, *b3;
    make_ctx(&user, Lside, Du, Dv, 0.0, 0.0);
    make_info(&info, mx, my);
    Field **aY = ghosted(mx, my, Y, &b1), **aYd = ghosted(mx, my, Ydot, &b2), **aF = ghosted(mx, my, Y, &b3);
    int rc = FormIFunctionLocal(&info, 0.0, aY, aYd, aF, &user);
    for (int j = 0; j < my; j++)
        for (int i = 0; i < mx; i++) { F[2 * (j * mx + i)] = aF[j][i].u; F[2 * (j * mx + i) + 1] = aF[j][i].v; }
    free(b1); free(b2); free(b3); free(aY - 1); free(aYd - 1); free(aF - 1);
    return rc;
}

int ref_pattern_rhsfunction(int mx, int my, double phi, double kappa, const double *Y, double *G) {
    PatternCtx user;
    DMDALocalInfo info;
    Field *b1, *b2;
    make_ctx(&user, 2.5, 0.0, 0.0, phi, kappa);
    make_info(&info, mx, my);
    Field **aY = ghosted(mx, my, Y, &b1), **aG = ghosted(mx, my, Y, &b2);
    int rc = FormRHSFunctionLocal(&info, 0.0, aY, aG, &user);
    for (int j = 0; j < my; j++)
        for (int i = 0; i < mx; i++) { G[2 * (j * mx + i)] = aG[j][i].u; G[2 * (j * mx + i) + 1] = aG[j][i].v; }
    free(b1); free(b2); free(aY - 1); free(aG - 1);
    return rc;
}

/* COO triplets of FormIJacobianLocal (periodic stencil indices are NOT wrapped: i-1 may be -1, as the reference passes
 * them to MatSetValuesStencil); rows/cols are returned as (j*mx+i)*2+c with i,j wrapped here */
long ref_pattern_ijacobian(int mx, int my, double Lside, double Du, double Dv, double shift, long cap, int *row, int *col,
                           double *val) {
    PatternCtx user;
    DMDALocalInfo info;
    const int M[3] = {mx, my, 1};
    const double L3[3] = {Lside, Lside, 1.0};
    make_ctx(&user, Lside, Du, Dv, 0.0, 0.0);
    make_info(&info, mx, my);
    DM dm = refstub_dm(2, M, L3);
    refstub_set_dof(dm, 2);
    info.da = dm;
    Mat P = refstub_mat(dm);
    int rc = FormIJacobianLocal(&info, 0.0, NULL, NULL, shift, P, P, &user);
    long n = refstub_mat_nnz(P);
    if (rc || n > cap) n = -1;
    else refstub_mat_copy(P, row, col, val);
    refstub_mat_free(P);
    free(dm);
    return n;
}
