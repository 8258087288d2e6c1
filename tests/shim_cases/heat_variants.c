/* A small TS driver written against include/petsc.h (not from the reference): the method-of-lines heat system of
 * c/ch5/heat.c's discretisation restated in its own way (Neumann data through mirrored values in x, periodic in y, the same
 * source and flux functions), optionally plus terms the library's kernel does not have -- to check the routes of TSSolve
 * on a DMDA that is not pattern.c's:
 *   -variant 0   the model (must be RECOGNISED: time stepping on the device, also with another diffusivity -D0)
 *   -variant 1   the model - 40 u^3 (must NOT be recognised: host callbacks, matrix-free stage operator)
 *   -variant 2   the model + a term that acts only where u > 0.02: the probes (|u| <= 0.025 around the zero initial state,
 *                and the term is switched on only after the first 3 evaluations) cannot see it, the final state does
 *                -- the re-verification there must turn the run into a loud error
 * Prints sum(u) and max|u| of the final state. */
#include <petsc.h>

typedef struct { PetscReal D0; PetscInt variant, calls; } Ctx;

static PetscErrorCode G(DMDALocalInfo *info, PetscReal t, PetscReal **au, PetscReal **aG, Ctx *user) {
    const PetscInt mx = info->mx, my = info->my;
    const PetscReal hx = 1.0 / (mx - 1), hy = 1.0 / my;
    (void)t;
    user->calls++;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++) {
        const PetscReal y = j * hy, flux = PetscSinReal(6.0 * PETSC_PI * y);
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            const PetscReal x = i * hx, c = au[j][i];
            const PetscReal west = i > 0 ? au[j][i - 1] : au[j][1] + 2.0 * hx * flux;
            const PetscReal east = i < mx - 1 ? au[j][i + 1] : au[j][mx - 2];
            const PetscReal lap = (west - 2.0 * c + east) / (hx * hx) + (au[j - 1][i] - 2.0 * c + au[j + 1][i]) / (hy * hy);
            aG[j][i] = user->D0 * lap + 3.0 * PetscExpReal(-25.0 * (x - 0.6) * (x - 0.6)) * PetscSinReal(2.0 * PETSC_PI * y);
            if (user->variant == 1) aG[j][i] -= 40.0 * c * c * c;
            if (user->variant == 2 && user->calls > 3 && c > 0.02) aG[j][i] -= 5.0 * (c - 0.02);
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    Ctx user;
    DM da;
    TS ts;
    Vec u;
    DMDALocalInfo info;
    PetscReal **a, sum = 0.0, mxv = 0.0;
    PetscCall(PetscInitialize(&argc, &argv, NULL, "TS variants on heat.c's DMDA for the p4b200 shim\n"));
    user.D0 = 1.0; user.variant = 0; user.calls = 0;
    PetscOptionsBegin(PETSC_COMM_WORLD, "", "variants", "");
    PetscCall(PetscOptionsInt("-variant", "0 = the model, 1 = model + cubic term, 2 = model + a term the probes cannot see", "heat_variants.c", user.variant, &user.variant, NULL));
    PetscCall(PetscOptionsReal("-D0", "diffusivity", "heat_variants.c", user.D0, &user.D0, NULL));
    PetscOptionsEnd();
    PetscCall(DMDACreate2d(PETSC_COMM_WORLD, DM_BOUNDARY_NONE, DM_BOUNDARY_PERIODIC, DMDA_STENCIL_STAR, 5, 4, PETSC_DECIDE, PETSC_DECIDE,
                           1, 1, NULL, NULL, &da));
    PetscCall(DMSetFromOptions(da));
    PetscCall(DMSetUp(da));
    PetscCall(DMCreateGlobalVector(da, &u));
    PetscCall(TSCreate(PETSC_COMM_WORLD, &ts));
    PetscCall(TSSetProblemType(ts, TS_NONLINEAR));
    PetscCall(TSSetDM(ts, da));
    PetscCall(DMDATSSetRHSFunctionLocal(da, INSERT_VALUES, (DMDATSRHSFunctionLocal)G, &user));
    PetscCall(TSSetType(ts, TSBEULER));
    PetscCall(TSSetTime(ts, 0.0));
    PetscCall(TSSetMaxTime(ts, 0.02));
    PetscCall(TSSetTimeStep(ts, 0.002));
    PetscCall(TSSetExactFinalTime(ts, TS_EXACTFINALTIME_MATCHSTEP));
    PetscCall(TSSetFromOptions(ts));
    PetscCall(VecSet(u, 0.0));
    PetscCall(TSSolve(ts, u));
    PetscCall(DMDAGetLocalInfo(da, &info));
    PetscCall(DMDAVecGetArray(da, u, &a));
    for (PetscInt j = 0; j < info.my; j++)
        for (PetscInt i = 0; i < info.mx; i++) { sum += a[j][i]; if (PetscAbsReal(a[j][i]) > mxv) mxv = PetscAbsReal(a[j][i]); }
    PetscCall(DMDAVecRestoreArray(da, u, &a));
    PetscCall(PetscPrintf(PETSC_COMM_WORLD, "done on %d x %d grid: sum %.12e max %.12e (%d evaluations of G on the host)\n", info.mx, info.my,
                          sum, mxv, user.calls));
    PetscCall(VecDestroy(&u));
    PetscCall(TSDestroy(&ts));
    PetscCall(DMDestroy(&da));
    PetscCall(PetscFinalize());
    return 0;
}
