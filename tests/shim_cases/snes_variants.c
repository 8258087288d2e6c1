/* A small SNES driver written against include/petsc.h (not from the reference): the minimal-surface residual of
 * c/ch7/minimal.c's discretisation restated in its own way, with Dirichlet data of its own, optionally plus a reaction term
 * that the library's kernel does not have -- to check both routes of p4b_snes2d_solve under the shim:
 *   -variant 0   the model with other boundary data and exponent (must be RECOGNISED: residual on the device)
 *   -variant 1   the model + 5 hx hy u^3 (must NOT be recognised: the callback is evaluated on the host every time)
 *   -variant 2   the model + a term that acts only where u < -0.05: the probes (interior values in [0, 0.5]) cannot see
 *                it, the converged iterate (negative near x = 1) does -- the re-verification at the converged iterate must
 *                catch it and the solve must be repeated with the callback on the host
 *   -jac k       registers a Jacobian callback as well (PETSc calls it when -snes_fd_color is absent): 1 = the 5-point
 *                Laplacian rows of the unit square with the boundary columns dropped (what the library has as a kernel: must
 *                be ACCEPTED), 2 = those rows times two, 3 = the same rows with the boundary columns kept (both must be REFUSED)
 * Prints sum(u) and max(u) of the converged iterate so that a test can compare with an independent solve. */
#include <petsc.h>

typedef struct { PetscReal q; PetscInt variant, jac; } Ctx;

static PetscReal gfun(PetscReal x, PetscReal y) { return 0.4 * PetscSinReal(3.0 * x + 1.0) * PetscCosReal(2.0 * y) + 0.2 * x * y; }
static PetscReal diffusivity(PetscReal ux, PetscReal uy, PetscReal q) { return PetscPowReal(1.0 + ux * ux + uy * uy, q); }

static PetscErrorCode Residual(DMDALocalInfo *info, PetscReal **au, PetscReal **aF, Ctx *user) {
    const PetscInt mx = info->mx, my = info->my;
    const PetscReal hx = 1.0 / (mx - 1), hy = 1.0 / (my - 1);
    /* node value with the Dirichlet data substituted on the boundary (so that interior rows do not depend on boundary unknowns) */
#define VAL(ii, jj) (((ii) == 0 || (jj) == 0 || (ii) == mx - 1 || (jj) == my - 1) ? gfun((ii) * hx, (jj) * hy) : au[jj][ii])
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            if (i == 0 || j == 0 || i == mx - 1 || j == my - 1) { aF[j][i] = au[j][i] - gfun(i * hx, j * hy); continue; }
            const PetscReal c = au[j][i], e = VAL(i + 1, j), w = VAL(i - 1, j), n = VAL(i, j + 1), s = VAL(i, j - 1);
            const PetscReal ne = VAL(i + 1, j + 1), nw = VAL(i - 1, j + 1), se = VAL(i + 1, j - 1), sw = VAL(i - 1, j - 1);
            const PetscReal De = diffusivity((e - c) / hx, (n + ne - s - se) / (4.0 * hy), user->q);
            const PetscReal Dw = diffusivity((c - w) / hx, (nw + n - sw - s) / (4.0 * hy), user->q);
            const PetscReal Dn = diffusivity((e + ne - w - nw) / (4.0 * hx), (n - c) / hy, user->q);
            const PetscReal Ds = diffusivity((e + se - w - sw) / (4.0 * hx), (c - s) / hy, user->q);
            aF[j][i] = -(hy / hx) * (De * (e - c) - Dw * (c - w)) - (hx / hy) * (Dn * (n - c) - Ds * (c - s));
            if (user->variant == 1) aF[j][i] += 5.0 * hx * hy * c * c * c;
            if (user->variant == 2 && c < -0.05) aF[j][i] += 40.0 * hx * hy * (c + 0.05) * (c + 0.05);
        }
#undef VAL
    return 0;
}

/* the Laplacian as an approximate Jacobian of the residual above: one MatSetValuesStencil per row */
static PetscErrorCode LaplaceRows(DMDALocalInfo *info, PetscReal **au, Mat J, Mat P, Ctx *user) {
    const PetscInt mx = info->mx, my = info->my;
    const PetscReal rx = (PetscReal)(mx - 1) / (my - 1), ry = 1.0 / rx, f = user->jac == 2 ? 2.0 : 1.0;
    (void)au;
    for (PetscInt j = info->ys; j < info->ys + info->ym; j++)
        for (PetscInt i = info->xs; i < info->xs + info->xm; i++) {
            MatStencil row, col[5];
            PetscReal v[5];
            PetscInt n = 0;
            const PetscInt di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
            row.i = i; row.j = j; row.k = 0; row.c = 0;
            col[n] = row; v[n++] = f * 2.0 * (rx + ry);
            if (i > 0 && j > 0 && i < mx - 1 && j < my - 1)
                for (PetscInt d = 0; d < 4; d++) {
                    const PetscInt ii = i + di[d], jj = j + dj[d];
                    const int bd = ii == 0 || jj == 0 || ii == mx - 1 || jj == my - 1;
                    if (bd && user->jac != 3) continue;
                    col[n] = row; col[n].i = ii; col[n].j = jj;
                    v[n++] = -f * (d < 2 ? rx : ry);
                }
            PetscCall(MatSetValuesStencil(P, 1, &row, n, col, v, INSERT_VALUES));
        }
    PetscCall(MatAssemblyBegin(P, MAT_FINAL_ASSEMBLY));
    PetscCall(MatAssemblyEnd(P, MAT_FINAL_ASSEMBLY));
    if (J != P) {
        PetscCall(MatAssemblyBegin(J, MAT_FINAL_ASSEMBLY));
        PetscCall(MatAssemblyEnd(J, MAT_FINAL_ASSEMBLY));
    }
    return 0;
}

int main(int argc, char **argv) {
    Ctx user;
    DM da;
    SNES snes;
    Vec u0, u;
    DMDALocalInfo info;
    PetscReal **a, sum = 0.0, mxv = -1.0e300;
    PetscCall(PetscInitialize(&argc, &argv, NULL, "SNES variants for the p4b200 shim\n"));
    user.q = -0.35; user.variant = 0; user.jac = 0;
    PetscOptionsBegin(PETSC_COMM_WORLD, "", "variants", "");
    PetscCall(PetscOptionsInt("-variant", "0 = the model, 1 = model + reaction, 2 = model + a term the probes cannot see", "snes_variants.c", user.variant, &user.variant, NULL));
    PetscCall(PetscOptionsInt("-jac", "0 = no Jacobian callback, 1 = Laplacian rows, 2 = twice those, 3 = boundary columns kept", "snes_variants.c", user.jac, &user.jac, NULL));
    PetscCall(PetscOptionsReal("-q", "exponent of the diffusivity", "snes_variants.c", user.q, &user.q, NULL));
    PetscOptionsEnd();
    PetscCall(DMDACreate2d(PETSC_COMM_WORLD, DM_BOUNDARY_NONE, DM_BOUNDARY_NONE, DMDA_STENCIL_BOX, 5, 5, PETSC_DECIDE, PETSC_DECIDE,
                           1, 1, NULL, NULL, &da));
    PetscCall(DMSetFromOptions(da));
    PetscCall(DMSetUp(da));
    PetscCall(DMDASetUniformCoordinates(da, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0));
    PetscCall(SNESCreate(PETSC_COMM_WORLD, &snes));
    PetscCall(SNESSetDM(snes, da));
    PetscCall(DMDASNESSetFunctionLocal(da, INSERT_VALUES, (DMDASNESFunctionFn *)Residual, &user));
    if (user.jac) PetscCall(DMDASNESSetJacobianLocal(da, (DMDASNESJacobianFn *)LaplaceRows, &user));
    PetscCall(SNESSetFromOptions(snes));
    PetscCall(DMGetGlobalVector(da, &u0));
    PetscCall(VecSet(u0, 0.1));
    PetscCall(SNESSolve(snes, NULL, u0));
    PetscCall(DMRestoreGlobalVector(da, &u0));
    PetscCall(DMDestroy(&da));
    PetscCall(SNESGetDM(snes, &da));
    PetscCall(SNESGetSolution(snes, &u));
    PetscCall(DMDAGetLocalInfo(da, &info));
    PetscCall(DMDAVecGetArray(da, u, &a));
    for (PetscInt j = 0; j < info.my; j++)
        for (PetscInt i = 0; i < info.mx; i++) { sum += a[j][i]; if (a[j][i] > mxv) mxv = a[j][i]; }
    PetscCall(DMDAVecRestoreArray(da, u, &a));
    PetscCall(PetscPrintf(PETSC_COMM_WORLD, "done on %d x %d grid: sum %.12e max %.12e\n", info.mx, info.my, sum, mxv));
    PetscCall(SNESDestroy(&snes));
    PetscCall(PetscFinalize());
    return 0;
}
