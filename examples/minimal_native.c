/* minimal_native.c -- the whole c/ch7/minimal.c run from a C host through ONE call of the C ABI (include/p4b200.h).
 *
 * What a maintainer binds instead of PETSc's SNESSolve: the options are minimal.c's own (c/ch7/minimal.c:69-103) plus
 * the PETSc ones its makefile / c/ch8/cluster.sh:70 pass.  Build (see p4pdes_b200/build.py:build_examples):
 *   gcc -std=c99 -pedantic -I include examples/minimal_native.c -L p4pdes_b200/lib -lp4b200 -Wl,-rpath,... -o minimal_native
 * Run:   ./minimal_native [-ms_problem tent|catenoid] [-ms_catenoid_c c] [-da_grid_x n] [-da_grid_y n] [-da_refine r]
 *                         [-snes_grid_sequence k] [-ksp_type gmres|cg] [-pc_type mg|none] [-snes_monitor_short]
 *                         [-snes_converged_reason] [-ksp_converged_reason]
 * e.g.   ./minimal_native -snes_converged_reason -snes_monitor_short -ms_problem catenoid -ms_catenoid_c 2.0 -da_refine 1
 *        prints what c/ch7/output/minimal.test1 holds (same first and last line; the Newton path in between is
 *        rounding-noise limited, see tests/test_minimal_oracle.py). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p4b200.h"

static void print_line(const char *line, void *ctx) { (void)ctx; puts(line); }

int main(int argc, char **argv) {
    p4b_minimal_opts o;
    p4b_minimal_result r;
    p4b_ctx *ctx = NULL;
    int i;
    if (p4b_minimal_default_opts(&o)) return 1;
    for (i = 1; i < argc; i++) {
        const char *a = argv[i], *v = i + 1 < argc ? argv[i + 1] : "";
        if (!strcmp(a, "-ms_problem")) { o.problem = strcmp(v, "tent") ? 1 : 0; i++; }
        else if (!strcmp(a, "-ms_q")) { o.q = atof(v); i++; }
        else if (!strcmp(a, "-ms_catenoid_c")) { o.catenoid_c = atof(v); i++; }
        else if (!strcmp(a, "-ms_tent_H")) { o.tent_H = atof(v); i++; }
        else if (!strcmp(a, "-da_grid_x")) { o.grid_x = atoi(v); i++; }
        else if (!strcmp(a, "-da_grid_y")) { o.grid_y = atoi(v); i++; }
        else if (!strcmp(a, "-da_refine")) { o.refine = atoi(v); i++; }
        else if (!strcmp(a, "-snes_grid_sequence")) { o.grid_sequence = atoi(v); i++; }
        else if (!strcmp(a, "-ksp_type")) { o.ksp_type = strcmp(v, "cg") ? 0 : 1; i++; }
        else if (!strcmp(a, "-pc_type")) { o.pc_type = strcmp(v, "mg") ? 0 : 1; i++; }
        else if (!strcmp(a, "-snes_monitor_short")) o.snes_monitor = 2;
        else if (!strcmp(a, "-snes_monitor")) o.snes_monitor = 1;
        else if (!strcmp(a, "-snes_converged_reason")) o.snes_converged_reason = 1;
        else if (!strcmp(a, "-ksp_converged_reason")) o.ksp_converged_reason = 1;
        else if (!strcmp(a, "-snes_fd_color")) { /* the assembled (FD-coloured) Jacobian: the default */ }
        else if (!strcmp(a, "-snes_mf_operator")) o.mf_operator = 1;
        else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (p4b_ctx_create(0, NULL, &ctx)) { fprintf(stderr, "%s\n", p4b_last_error()); return 1; }
    if (p4b_minimal_solve(ctx, &o, print_line, NULL, NULL, 0, &r)) { fprintf(stderr, "%s\n", p4b_last_error()); return 1; }
    p4b_ctx_destroy(ctx);
    return 0;
}
