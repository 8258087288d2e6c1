// nk_host_main.cpp -- runs p4pdes_b200/csrc/nk_solver.hpp on HostOps (TEST INFRASTRUCTURE ONLY; see host_ops.hpp).
//   nk_host_test [-ms_problem tent|catenoid] [-ms_q q] [-ms_catenoid_c c] [-da_grid_x n] [-da_grid_y n] [-da_refine r]
//                [-snes_grid_sequence k] [-ksp_type gmres|cg] [-pc_type mg|none] [-pc_mg_levels n] [-monitor]
// prints the solver's lines, then one JSON line with the per-stage results and a checksum of the solution.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "nk_solver.hpp"
#include "ts_solver.hpp"
#include "host_ops.hpp"

static void print_line(const char *s, void *) { printf("%s\n", s); }

// nk_host_test -pattern [-da_grid_x n] [-da_grid_y n] [-da_refine r] [-ts_type arkimex|beuler|cn] [-ts_dt h] [-ts_max_time T]
//              [-pc_type mg|none] [-snes_rtol r] [-ptn_no_rhsjacobian] [-ptn_call_back_report] [-p4b_mg_rscale s]
//              [-ts_monitor] [-snes_converged_reason] [-ksp_converged_reason]
static int pattern_main(int argc, char **argv) {
    using namespace p4b::nk;
    PatternOpts o;
    default_opts(&o);
    bool cgs = false;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-da_grid_x") o.grid_x = atoi(next());
        else if (a == "-da_grid_y") o.grid_y = atoi(next());
        else if (a == "-da_refine") o.refine = atoi(next());
        else if (a == "-ts_type") { const std::string v = next(); o.ts_type = v == "beuler" ? TS_BEULER : (v == "cn" ? TS_CN : (v == "bdf" ? TS_BDF : TS_ARKIMEX)); }
        else if (a == "-ts_dt") o.ts_dt = atof(next());
        else if (a == "-ts_max_time") o.ts_max_time = atof(next());
        else if (a == "-pc_type") o.pc_type = std::string(next()) == "mg" ? PC_MG : PC_NONE;
        else if (a == "-snes_rtol") o.snes_rtol = atof(next());
        else if (a == "-ksp_rtol") o.ksp_rtol = atof(next());
        else if (a == "-gmres_cgs") cgs = true;
        else if (a == "-p4b_mg_rscale") o.mg_rscale = atof(next());
        else if (a == "-ptn_no_rhsjacobian") o.no_rhsjacobian = 1;
        else if (a == "-ptn_call_back_report") o.call_back_report = 1;
        else if (a == "-ts_monitor") o.ts_monitor = 1;
        else if (a == "-snes_converged_reason") o.snes_converged_reason = 1;
        else if (a == "-ksp_converged_reason") o.ksp_converged_reason = 1;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    HostOps ops;
    ops.cgs = cgs;
    PatternResult R;
    double *Y = nullptr;
    Printer pr{print_line, nullptr};
    const int rc = pattern_solve(&ops, o, pr, &Y, &R);
    if (rc) { fprintf(stderr, "pattern_solve failed: %d\n", rc); return 1; }
    double sum = 0.0, vmax = 0.0;
    for (int k = 0; k < R.m * R.m; k++) { sum += Y[2 * k] + 3.0 * Y[2 * k + 1]; vmax = std::max(vmax, Y[2 * k + 1]); }
    printf("{\"m\": %d, \"nsteps\": %d, \"rejected\": %d, \"ksp_its_total\": %lld, \"newton_its_total\": %lld, \"t_final\": %.17g, "
           "\"sum\": %.17g, \"vmax\": %.17g, \"allocs\": %lld, \"frees\": %lld, \"step_newton\": [",
           R.m, R.nsteps, R.rejected, R.ksp_its_total, R.newton_its_total, R.t_final, sum, vmax, ops.allocs, ops.frees + 1);
    for (int k = 0; k < R.nsteps && k < MAX_TS_STEPS_KEPT; k++) printf("%s%d", k ? ", " : "", R.step_newton[k]);
    printf("]}\n");
    ops.release(Y);
    return 0;
}

// -callback <libfishref.so>: the residual is the REFERENCE's own FormFunctionLocal (c/ch7/minimal.c:210-282, compiled
// unchanged by oracle/refstub into oracle/_ref/libfishref.so), reached through the p4b_residual2d_fn contract
typedef int (*ref_minimal_fn)(const int *M, int problem, double q, double tent_H, double catenoid_c, double *u, double *FF);
struct RefUser { ref_minimal_fn f; int problem; double q, H, c; };
static int ref_residual(void *user, int mx, int my, const double *u, double *F) {
    RefUser *r = (RefUser *)user;
    const int M[3] = {mx, my, 1};
    return r->f(M, r->problem, r->q, r->H, r->c, const_cast<double *>(u), F);
}

int main(int argc, char **argv) {
    using namespace p4b::nk;
    if (argc > 1 && std::string(argv[1]) == "-pattern") return pattern_main(argc, argv);
    MinimalOpts o;
    default_opts(&o);
    const char *cb_lib = nullptr;
    bool cgs = false, fd_color = false, mf_poisson = false;
    for (int i = 1; i < argc; i++) {
        if (std::string(argv[i]) == "-callback" && i + 1 < argc) { cb_lib = argv[i + 1]; for (int k = i; k + 2 < argc; k++) argv[k] = argv[k + 2]; argc -= 2; break; }
    }
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-ms_problem") o.problem = std::string(next()) == "tent" ? 0 : 1;
        else if (a == "-ms_q") o.q = atof(next());
        else if (a == "-ms_catenoid_c") o.catenoid_c = atof(next());
        else if (a == "-ms_tent_H") o.tent_H = atof(next());
        else if (a == "-da_grid_x") o.grid_x = atoi(next());
        else if (a == "-da_grid_y") o.grid_y = atoi(next());
        else if (a == "-da_refine") o.refine = atoi(next());
        else if (a == "-snes_grid_sequence") o.grid_sequence = atoi(next());
        else if (a == "-ksp_type") o.ksp_type = std::string(next()) == "cg" ? KSP_CG : KSP_GMRES;
        else if (a == "-pc_type") o.pc_type = std::string(next()) == "mg" ? PC_MG : PC_NONE;
        else if (a == "-pc_mg_levels") o.mg_levels = atoi(next());
        else if (a == "-snes_fd_color") fd_color = true;
        else if (a == "-snes_mf_operator") o.mf_operator = 1;
        else if (a == "-p4b_mf_pmat") mf_poisson = std::string(next()) == "poisson";
        else if (a == "-snes_max_it") o.snes_max_it = atoi(next());
        else if (a == "-gmres_cgs") cgs = true;
        else if (a == "-monitor") { o.snes_monitor = 2; o.snes_converged_reason = 1; o.ksp_converged_reason = 1; }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    // the matrix: FD-coloured, or the Poisson one minimal.c registers (no -snes_fd_color; under -snes_mf_operator on request)
    o.jacobian = (!fd_color && !o.mf_operator) || (o.mf_operator && mf_poisson) ? 1 : 0;
    MinimalResult R;
    double *u = nullptr;
    Printer pr{print_line, nullptr};
    if (cb_lib) {
        void *h = dlopen(cb_lib, RTLD_NOW);
        if (!h) { fprintf(stderr, "%s\n", dlerror()); return 3; }
        RefUser ru{(ref_minimal_fn)dlsym(h, "ref_minimal_function"), o.problem, o.q, o.tent_H, o.catenoid_c};
        if (!ru.f) { fprintf(stderr, "ref_minimal_function not found\n"); return 3; }
        // the caller's part of minimal.c:main: the grid and the initial iterate (InitialState: g on the boundary, 0 inside)
        int mx = o.grid_x, my = o.grid_y;
        for (int r = 0; r < o.refine; r++) { mx = 2 * mx - 1; my = 2 * my - 1; }
        HostOps tmp;
        std::vector<double> g((size_t)mx * my), u0((size_t)mx * my);
        tmp.minimal_sample(mx, my, o.problem, o.tent_H, o.catenoid_c, g.data());
        tmp.initial_state2d(mx, my, g.data(), u0.data());
        HostCallbackOps cops;
        cops.cgs = cgs;
        cops.fn = ref_residual;
        cops.user = &ru;
        const int rc = minimal_solve(&cops, o, pr, &u, &R, u0.data(), false);
        if (rc) { fprintf(stderr, "minimal_solve (callback) failed: %d\n", rc); return 1; }
        // the caller's error report (minimal.c:166-181)
        std::vector<double> gf((size_t)R.mx * R.my);
        tmp.minimal_sample(R.mx, R.my, o.problem, o.tent_H, o.catenoid_c, gf.data());
        double e = 0.0, sum = 0.0;
        for (int n = 0; n < R.mx * R.my; n++) { e = std::max(e, fabs(u[n] - gf[n])); sum += u[n]; }
        printf("{\"mx\": %d, \"my\": %d, \"errinf\": %.17g, \"sum\": %.17g, \"callbacks\": %lld, \"stages\": [", R.mx, R.my, e, sum, cops.callbacks);
        for (int s = 0; s < R.nstages; s++) {
            const StageResult &S = R.stage[s];
            printf("%s{\"mx\": %d, \"its\": %d, \"reason\": \"%s\", \"ksp_its\": [", s ? ", " : "", S.mx, S.its, snes_reason_name(S.reason));
            for (int k = 0; k < S.its; k++) printf("%s%d", k ? ", " : "", S.ksp_its[k]);
            printf("]}");
        }
        printf("]}\n");
        cops.release(u);
        return 0;
    }
    HostOps ops;
    ops.cgs = cgs;
    const int rc = minimal_solve(&ops, o, pr, &u, &R);
    if (rc) { fprintf(stderr, "minimal_solve failed: %d\n", rc); return 1; }
    double sum = 0.0, sum2 = 0.0;
    for (int n = 0; n < R.mx * R.my; n++) { sum += u[n]; sum2 += u[n] * u[n] * (1 + n % 7); }
    printf("{\"mx\": %d, \"my\": %d, \"errinf\": %.17g, \"sum\": %.17g, \"wsum2\": %.17g, \"allocs\": %lld, \"frees\": %lld, \"stages\": [",
           R.mx, R.my, R.errinf, sum, sum2, ops.allocs, ops.frees + 1);
    for (int s = 0; s < R.nstages; s++) {
        const StageResult &S = R.stage[s];
        printf("%s{\"mx\": %d, \"its\": %d, \"reason\": \"%s\", \"ksp_its\": [", s ? ", " : "", S.mx, S.its, snes_reason_name(S.reason));
        for (int k = 0; k < S.its; k++) printf("%s%d", k ? ", " : "", S.ksp_its[k]);
        printf("], \"lambda\": [");
        for (int k = 0; k < S.its; k++) printf("%s%.17g", k ? ", " : "", S.lambda[k]);
        printf("], \"fnorm\": [");
        for (int k = 0; k <= S.its; k++) printf("%s%.17g", k ? ", " : "", S.fnorm[k]);
        printf("]}");
    }
    printf("]}\n");
    ops.release(u);
    return 0;
}
