/* A small TS driver written against include/petsc.h (not from the reference): a two-species reaction-diffusion system on
 * the periodic 2-D DMDA whose callbacks can be made to DIFFER from the model the device path implements, to check that
 * TSSolve of the shim refuses them instead of silently running its own kernels.
 *   -variant 0   the model itself, with other parameter values than pattern.c's defaults (must run)
 *   -variant 1   reaction with an extra cubic term in G^u
 *   -variant 2   anisotropic diffusion in F (x and y edges weighted differently)
 *   -variant 3   IJacobian with a wrong corner weight
 *   -variant 4   RHSJacobian with a wrong off-diagonal entry (fully implicit types only)
 *   -variant 5   a forcing in G^u that grows with t (zero at t = 0: one probe at t = 0 cannot see it) */
#include <petsc.h>

typedef struct { PetscReal u, v; } Field;
typedef struct { PetscReal L, Du, Dv, phi, kappa; PetscInt variant; } Ctx;

static PetscErrorCode RHS(DMDALocalInfo *info, PetscReal t, Field **aY, Field **aG, Ctx *user) {
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            const PetscReal u = aY[j][i].u, v = aY[j][i].v, uv2 = u * v * v;
            aG[j][i].u = -uv2 + user->phi * (1.0 - u) + (user->variant == 1 ? 1.0e-3 * u * u * u : 0.0)
                         + (user->variant == 5 ? 1.0e-4 * t : 0.0);
            aG[j][i].v = uv2 - (user->phi + user->kappa) * v;
        }
    return 0;
}
static PetscErrorCode IF(DMDALocalInfo *info, PetscReal t, Field **aY, Field **aYdot, Field **aF, Ctx *user) {
    (void)t;
    const PetscReal h = user->L / (PetscReal)info->mx, Cu = user->Du / (6.0 * h * h), Cv = user->Dv / (6.0 * h * h);
    const PetscReal wy = user->variant == 2 ? 4.5 : 4.0;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            const PetscReal lapu = aY[j+1][i-1].u + wy * aY[j+1][i].u + aY[j+1][i+1].u + 4.0 * aY[j][i-1].u - (12.0 + 2.0 * wy) * aY[j][i].u
                                   + 4.0 * aY[j][i+1].u + aY[j-1][i-1].u + wy * aY[j-1][i].u + aY[j-1][i+1].u;
            const PetscReal lapv = aY[j+1][i-1].v + 4.0 * aY[j+1][i].v + aY[j+1][i+1].v + 4.0 * aY[j][i-1].v - 20.0 * aY[j][i].v
                                   + 4.0 * aY[j][i+1].v + aY[j-1][i-1].v + 4.0 * aY[j-1][i].v + aY[j-1][i+1].v;
            aF[j][i].u = aYdot[j][i].u - Cu * lapu;
            aF[j][i].v = aYdot[j][i].v - Cv * lapv;
        }
    return 0;
}
static PetscErrorCode IJac(DMDALocalInfo *info, PetscReal t, Field **aY, Field **aYdot, PetscReal shift, Mat J, Mat P, Ctx *user) {
    (void)t; (void)aY; (void)aYdot; (void)J;
    const PetscReal h = user->L / (PetscReal)info->mx;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++)
            for (PetscInt c = 0; c < 2; c++) {
                const PetscReal CC = (c == 0 ? user->Du : user->Dv) / (6.0 * h * h);
                MatStencil row, col[9];
                PetscReal val[9];
                PetscInt s = 0;
                row.i = i; row.j = j; row.c = c; row.k = 0;
                for (PetscInt dj = -1; dj <= 1; dj++)
                    for (PetscInt di = -1; di <= 1; di++) {
                        col[s].i = i + di; col[s].j = j + dj; col[s].c = c; col[s].k = 0;
                        val[s] = (!di && !dj) ? shift + 20.0 * CC : ((!di || !dj) ? -4.0 * CC : -(user->variant == 3 ? 1.25 : 1.0) * CC);
                        s++;
                    }
                PetscCall(MatSetValuesStencil(P, 1, &row, 9, col, val, INSERT_VALUES));
            }
    PetscCall(MatAssemblyBegin(P, MAT_FINAL_ASSEMBLY));
    PetscCall(MatAssemblyEnd(P, MAT_FINAL_ASSEMBLY));
    return 0;
}
static PetscErrorCode RHSJac(DMDALocalInfo *info, PetscReal t, Field **aY, Mat J, Mat P, Ctx *user) {
    (void)t; (void)J;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            const PetscReal u = aY[j][i].u, v = aY[j][i].v;
            MatStencil row, col[2];
            PetscReal val[2];
            row.i = col[0].i = col[1].i = i; row.j = col[0].j = col[1].j = j; row.k = col[0].k = col[1].k = 0;
            col[0].c = 0; col[1].c = 1;
            row.c = 0; val[0] = -v * v - user->phi; val[1] = -2.0 * u * v * (user->variant == 4 ? 1.01 : 1.0);
            PetscCall(MatSetValuesStencil(P, 1, &row, 2, col, val, INSERT_VALUES));
            row.c = 1; val[0] = v * v; val[1] = 2.0 * u * v - (user->phi + user->kappa);
            PetscCall(MatSetValuesStencil(P, 1, &row, 2, col, val, INSERT_VALUES));
        }
    PetscCall(MatAssemblyBegin(P, MAT_FINAL_ASSEMBLY));
    PetscCall(MatAssemblyEnd(P, MAT_FINAL_ASSEMBLY));
    return 0;
}

int main(int argc, char **argv) {
    Ctx user;
    DM da;
    TS ts;
    Vec x;
    Field **aY;
    DMDALocalInfo info;
    PetscReal nrm;
    PetscCall(PetscInitialize(&argc, &argv, NULL, "TS variants for the p4b200 shim\n"));
    user.L = 2.0; user.Du = 6.0e-5; user.Dv = 3.5e-5; user.phi = 0.03; user.kappa = 0.055; user.variant = 0;
    PetscOptionsBegin(PETSC_COMM_WORLD, "", "variants", "");
    PetscCall(PetscOptionsInt("-variant", "which callback deviates from the model", "ts_variants.c", user.variant, &user.variant, NULL));
    PetscOptionsEnd();
    PetscCall(DMDACreate2d(PETSC_COMM_WORLD, DM_BOUNDARY_PERIODIC, DM_BOUNDARY_PERIODIC, DMDA_STENCIL_BOX, 4, 4, PETSC_DECIDE,
                           PETSC_DECIDE, 2, 1, NULL, NULL, &da));
    PetscCall(DMSetFromOptions(da));
    PetscCall(DMSetUp(da));
    PetscCall(DMDASetUniformCoordinates(da, 0.0, user.L, 0.0, user.L, -1.0, -1.0));
    PetscCall(TSCreate(PETSC_COMM_WORLD, &ts));
    PetscCall(TSSetProblemType(ts, TS_NONLINEAR));
    PetscCall(TSSetDM(ts, da));
    PetscCall(DMDATSSetRHSFunctionLocal(da, INSERT_VALUES, (DMDATSRHSFunctionLocal)RHS, &user));
    PetscCall(DMDATSSetRHSJacobianLocal(da, (DMDATSRHSJacobianLocal)RHSJac, &user));
    PetscCall(DMDATSSetIFunctionLocal(da, INSERT_VALUES, (DMDATSIFunctionLocal)IF, &user));
    PetscCall(DMDATSSetIJacobianLocal(da, (DMDATSIJacobianLocal)IJac, &user));
    PetscCall(TSSetType(ts, TSARKIMEX));
    PetscCall(TSSetTime(ts, 0.0));
    PetscCall(TSSetMaxTime(ts, 20.0));
    PetscCall(TSSetTimeStep(ts, 2.0));
    PetscCall(TSSetExactFinalTime(ts, TS_EXACTFINALTIME_MATCHSTEP));
    PetscCall(TSSetFromOptions(ts));
    PetscCall(DMCreateGlobalVector(da, &x));
    PetscCall(DMDAGetLocalInfo(da, &info));
    PetscCall(DMDAVecGetArray(da, x, &aY));
    for (PetscInt j = 0; j < info.my; j++)
        for (PetscInt i = 0; i < info.mx; i++) {
            const PetscReal sx = PetscSinReal(2.0 * PETSC_PI * i / info.mx), sy = PetscSinReal(2.0 * PETSC_PI * j / info.my);
            aY[j][i].v = 0.25 * sx * sx * sy * sy;
            aY[j][i].u = 1.0 - 2.0 * aY[j][i].v;
        }
    PetscCall(DMDAVecRestoreArray(da, x, &aY));
    PetscCall(TSSolve(ts, x));
    PetscCall(VecNorm(x, NORM_2, &nrm));
    PetscCall(PetscPrintf(PETSC_COMM_WORLD, "done: |Y|_2 = %.10e\n", nrm));
    PetscCall(VecDestroy(&x));
    PetscCall(TSDestroy(&ts));
    PetscCall(DMDestroy(&da));
    PetscCall(PetscFinalize());
    return 0;
}
