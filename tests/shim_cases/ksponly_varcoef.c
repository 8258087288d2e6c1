/* A small linear driver written against include/petsc.h (not from the reference): -div(a(x,y) grad u) + c u = f on the unit
 * square with Dirichlet data, 5-point finite volumes, solved as fish.c solves its problem (SNESKSPONLY + KSPCG, residual and
 * Jacobian callbacks).  The coefficient varies, so the Jacobian callback's values are NOT the constant-coefficient stencil
 * the shim's structured Mat type recognises: "stencilcuda" must refuse them, the assembled type (-mat_type sellcuda: values
 * kept as inserted, SELL-32 SpMV on the device) must solve the system.  Prints sum(u) and u at the centre. */
#include <petsc.h>

static PetscReal coef(PetscReal x, PetscReal y) { return 1.0 + 0.5 * PetscSinReal(3.0 * x) * PetscCosReal(2.0 * y); }
static PetscReal gbd(PetscReal x, PetscReal y) { return PetscSinReal(x) + y * y; }
static PetscReal rhs(PetscReal x, PetscReal y) { return 1.0 + x * y; }
#define C0 2.0

static PetscErrorCode Residual(DMDALocalInfo *info, PetscReal **au, PetscReal **aF, void *ctx) {
    const PetscInt mx = info->mx, my = info->my;
    const PetscReal hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1);
    (void)ctx;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            const PetscReal x = i * hx, y = j * hy;
            if (i == 0 || j == 0 || i == mx - 1 || j == my - 1) { aF[j][i] = au[j][i] - gbd(x, y); continue; }
            const PetscReal ue = (i + 1 == mx - 1) ? gbd(x + hx, y) : au[j][i + 1], uw = (i - 1 == 0) ? gbd(x - hx, y) : au[j][i - 1];
            const PetscReal un = (j + 1 == my - 1) ? gbd(x, y + hy) : au[j + 1][i], us = (j - 1 == 0) ? gbd(x, y - hy) : au[j - 1][i];
            aF[j][i] = (hy / hx) * (coef(x + 0.5 * hx, y) * (au[j][i] - ue) + coef(x - 0.5 * hx, y) * (au[j][i] - uw))
                     + (hx / hy) * (coef(x, y + 0.5 * hy) * (au[j][i] - un) + coef(x, y - 0.5 * hy) * (au[j][i] - us))
                     + hx * hy * (C0 * au[j][i] - rhs(x, y));
        }
    return 0;
}

static PetscErrorCode Jacobian(DMDALocalInfo *info, PetscReal **au, Mat J, Mat P, void *ctx) {
    const PetscInt mx = info->mx, my = info->my;
    const PetscReal hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1);
    (void)au; (void)ctx; (void)J;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            MatStencil row, col[5];
            PetscReal v[5];
            PetscInt n = 0;
            const PetscReal x = i * hx, y = j * hy;
            row.i = i; row.j = j; row.k = 0; row.c = 0;
            col[n].i = i; col[n].j = j; col[n].k = 0; col[n].c = 0;
            if (i == 0 || j == 0 || i == mx - 1 || j == my - 1) { v[n++] = 1.0; }
            else {
                const PetscReal ae = (hy / hx) * coef(x + 0.5 * hx, y), aw = (hy / hx) * coef(x - 0.5 * hx, y);
                const PetscReal an = (hx / hy) * coef(x, y + 0.5 * hy), as = (hx / hy) * coef(x, y - 0.5 * hy);
                v[n++] = ae + aw + an + as + hx * hy * C0;
                if (i + 1 < mx - 1) { col[n].i = i + 1; col[n].j = j; col[n].k = 0; col[n].c = 0; v[n++] = -ae; }
                if (i - 1 > 0)      { col[n].i = i - 1; col[n].j = j; col[n].k = 0; col[n].c = 0; v[n++] = -aw; }
                if (j + 1 < my - 1) { col[n].i = i; col[n].j = j + 1; col[n].k = 0; col[n].c = 0; v[n++] = -an; }
                if (j - 1 > 0)      { col[n].i = i; col[n].j = j - 1; col[n].k = 0; col[n].c = 0; v[n++] = -as; }
            }
            PetscCall(MatSetValuesStencil(P, 1, &row, n, col, v, INSERT_VALUES));
        }
    PetscCall(MatAssemblyBegin(P, MAT_FINAL_ASSEMBLY));
    PetscCall(MatAssemblyEnd(P, MAT_FINAL_ASSEMBLY));
    return 0;
}

int main(int argc, char **argv) {
    DM da;
    SNES snes;
    KSP ksp;
    Vec u0, u;
    DMDALocalInfo info;
    PetscReal **a, sum = 0.0;
    PetscCall(PetscInitialize(&argc, &argv, NULL, "variable-coefficient KSPONLY driver for the p4b200 shim\n"));
    PetscCall(DMDACreate2d(PETSC_COMM_WORLD, DM_BOUNDARY_NONE, DM_BOUNDARY_NONE, DMDA_STENCIL_STAR, 5, 5, PETSC_DECIDE, PETSC_DECIDE,
                           1, 1, NULL, NULL, &da));
    PetscCall(DMSetFromOptions(da));
    PetscCall(DMSetUp(da));
    PetscCall(DMDASetUniformCoordinates(da, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0));
    PetscCall(SNESCreate(PETSC_COMM_WORLD, &snes));
    PetscCall(SNESSetDM(snes, da));
    PetscCall(DMDASNESSetFunctionLocal(da, INSERT_VALUES, (DMDASNESFunctionFn *)Residual, NULL));
    PetscCall(DMDASNESSetJacobianLocal(da, (DMDASNESJacobianFn *)Jacobian, NULL));
    PetscCall(SNESSetType(snes, SNESKSPONLY));
    PetscCall(SNESGetKSP(snes, &ksp));
    PetscCall(KSPSetType(ksp, KSPCG));
    PetscCall(SNESSetFromOptions(snes));
    PetscCall(DMGetGlobalVector(da, &u0));
    PetscCall(VecSet(u0, 0.0));
    PetscCall(SNESSolve(snes, NULL, u0));
    PetscCall(DMRestoreGlobalVector(da, &u0));
    PetscCall(SNESGetSolution(snes, &u));
    PetscCall(DMDAGetLocalInfo(da, &info));
    PetscCall(DMDAVecGetArray(da, u, &a));
    for (PetscInt j = 0; j < info.my; j++)
        for (PetscInt i = 0; i < info.mx; i++) sum += a[j][i];
    PetscCall(PetscPrintf(PETSC_COMM_WORLD, "done on %d x %d grid: sum %.12e centre %.12e\n", info.mx, info.my, sum,
                          a[info.my / 2][info.mx / 2]));
    PetscCall(DMDAVecRestoreArray(da, u, &a));
    PetscCall(DMDestroy(&da));
    PetscCall(SNESDestroy(&snes));
    PetscCall(PetscFinalize());
    return 0;
}
