/* refwrap_minimal.c -- includes the reference's c/ch7/minimal.c VERBATIM (main renamed) and exposes its
 * FormFunctionLocal and boundary functions for the oracle tests.  TEST INFRASTRUCTURE ONLY. */
#define main minimal_reference_main
#include REF_MINIMAL_C
#undef main
#include <stdlib.h>

DM refstub_dm(int dim, const int *M, const double *L);
void *refstub_tables(int dim, const int *M, double *base);

/* FF = FormFunctionLocal(u) of minimal.c:210-282 ; problem 0 = tent, 1 = catenoid */
int ref_minimal_function(const int *M, int problem, double q, double tent_H, double catenoid_c, double *u, double *FF) {
    PoissonCtx user;
    MinimalCtx mctx;
    DMDALocalInfo info;
    const double L[3] = {1.0, 1.0, 1.0};
    mctx.q = q; mctx.tent_H = tent_H; mctx.catenoid_c = catenoid_c; mctx.quaddegree = 3;
    user.Lx = user.Ly = user.Lz = 1.0;
    user.cx = user.cy = user.cz = 1.0;
    user.g_bdry = problem == 0 ? &g_bdry_tent : &g_bdry_catenoid;
    user.f_rhs = NULL;
    user.addctx = &mctx;
    DM dm = refstub_dm(2, M, L);
    DMDAGetLocalInfo(dm, &info);
    void *au = refstub_tables(2, M, u), *aF = refstub_tables(2, M, FF);
    int rc = FormFunctionLocal(&info, (PetscReal **)au, (PetscReal **)aF, &user);
    free(au); free(aF); free(dm);
    return rc;
}

/* g sampled at every node (FormExactFromG's loop, minimal.c:191-208) */
int ref_minimal_g(const int *M, int problem, double tent_H, double catenoid_c, double *g) {
    PoissonCtx user;
    MinimalCtx mctx;
    mctx.q = -0.5; mctx.tent_H = tent_H; mctx.catenoid_c = catenoid_c; mctx.quaddegree = 3;
    user.addctx = &mctx;
    const double hx = 1.0 / (M[0] - 1), hy = 1.0 / (M[1] - 1);
    for (int j = 0; j < M[1]; j++)
        for (int i = 0; i < M[0]; i++)
            g[j * M[0] + i] = problem == 0 ? g_bdry_tent(i * hx, j * hy, 0.0, &user) : g_bdry_catenoid(i * hx, j * hy, 0.0, &user);
    return 0;
}
