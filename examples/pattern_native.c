/* pattern_native.c -- the whole c/ch5/pattern.c run from a C host through ONE call of the C ABI (include/p4b200.h):
 * what a maintainer binds instead of PETSc's TSSolve.  Options: pattern.c's own (c/ch5/pattern.c:54-78) plus the PETSc
 * ones its makefile passes (c/ch5/makefile:49-62).
 *   ./pattern_native -da_grid_x 4 -da_grid_y 4 -da_refine 2 -ts_monitor          prints c/ch5/output/pattern.test1
 *   ./pattern_native -da_refine 2 -ts_monitor -ts_dt 1 -ts_max_time 1 -ts_type beuler -pc_type mg -snes_converged_reason \
 *                    -ksp_converged_reason -snes_rtol 1.0e-1 -ptn_no_rhsjacobian  prints c/ch5/output/pattern.test2 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p4b200.h"

static void print_line(const char *line, void *ctx) { (void)ctx; puts(line); }

int main(int argc, char **argv) {
    p4b_pattern_opts o;
    p4b_pattern_result r;
    p4b_ctx *ctx = NULL;
    int i;
    if (p4b_pattern_default_opts(&o)) return 1;
    for (i = 1; i < argc; i++) {
        const char *a = argv[i], *v = i + 1 < argc ? argv[i + 1] : "";
        if (!strcmp(a, "-da_grid_x")) { o.grid_x = atoi(v); i++; }
        else if (!strcmp(a, "-da_grid_y")) { o.grid_y = atoi(v); i++; }
        else if (!strcmp(a, "-da_refine")) { o.refine = atoi(v); i++; }
        else if (!strcmp(a, "-ts_type")) { o.ts_type = !strcmp(v, "beuler") ? 1 : (!strcmp(v, "cn") ? 2 : (!strcmp(v, "bdf") ? 3 : 0)); i++; }
        else if (!strcmp(a, "-ts_dt")) { o.ts_dt = atof(v); i++; }
        else if (!strcmp(a, "-ts_max_time")) { o.ts_max_time = atof(v); i++; }
        else if (!strcmp(a, "-pc_type")) { o.pc_type = strcmp(v, "mg") ? 0 : 1; i++; }
        else if (!strcmp(a, "-snes_rtol")) { o.snes_rtol = atof(v); i++; }
        else if (!strcmp(a, "-p4b_mg_rscale")) { o.mg_rscale = atof(v); i++; }
        else if (!strcmp(a, "-ptn_phi")) { o.phi = atof(v); i++; }
        else if (!strcmp(a, "-ptn_kappa")) { o.kappa = atof(v); i++; }
        else if (!strcmp(a, "-ptn_no_rhsjacobian")) o.no_rhsjacobian = 1;
        else if (!strcmp(a, "-ptn_call_back_report")) o.call_back_report = 1;
        else if (!strcmp(a, "-ts_monitor")) o.ts_monitor = 1;
        else if (!strcmp(a, "-snes_converged_reason")) o.snes_converged_reason = 1;
        else if (!strcmp(a, "-ksp_converged_reason")) o.ksp_converged_reason = 1;
        else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (p4b_ctx_create(0, NULL, &ctx)) { fprintf(stderr, "%s\n", p4b_last_error()); return 1; }
    if (p4b_pattern_solve(ctx, &o, print_line, NULL, NULL, 0, &r)) { fprintf(stderr, "%s\n", p4b_last_error()); return 1; }
    p4b_ctx_destroy(ctx);
    return 0;
}
