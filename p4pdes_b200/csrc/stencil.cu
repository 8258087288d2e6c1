// stencil.cu -- the 3-/5-/7-point stencil family (K1, K1r, K2 of SURVEY.md 2.3).
//
// The operator is the Jacobian the reference assembles in Poisson{1,2,3}DJacobianLocal
// (c/ch6/poissonfunctions.c:117-258), applied matrix-free:
//     boundary row      (A u)_p = diag * u_p
//     interior row      (A u)_p = diag * u_p - sum_d sc_d * (u_{p-e_d} + u_{p+e_d}),
//                       where a neighbour that is a boundary node is dropped (:133-138,:172-181,:221-245)
// [PETSc] MatMult / MatResidual / KSPSolve_Chebyshev + PCApply_Jacobi become one kernel each:
//     out = [ca*pm1 +] cb*u + cg*(b - A u)      or      out = A u [, (u, A u)].
#include "kernels.h"

namespace p4b {

// ---------------------------------------------------------------------------------------------
// generic kernel: one thread per point, linear index over the local slab.  Used for every level
// the plane-marching kernel does not take (small, 1-D and 2-D grids); neighbours come through L1/L2.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) stencil_generic_kernel(const LevelDesc L, const StencilOp op, double *partials,
                                                               unsigned int *ticket) {
    const long long nloc = L.nlocal();
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    double dv[2] = {0.0, 0.0};
    const PortSpan span = port_span_linear(op.port, (long long)L.nx * L.ny, nloc, 256);
    port_wait(op.port, span.boundary);
    if (n < nloc) {
        const int plane = L.nx * L.ny;
        const int kl = (int)(n / plane);
        const int rem = (int)(n - (long long)kl * plane);
        const int j = rem / L.nx;
        const int i = rem - j * L.nx;
        const int k = kl + L.zs;
        const double *__restrict__ u = op.u;
        const double uc = u[n];
        const bool bd = (L.ax && (i == 0 || i == L.nx - 1)) || (L.ay && (j == 0 || j == L.ny - 1)) ||
                        (L.az && (k == 0 || k == L.nz - 1));
        double Au = L.diag * uc;
        if (!bd) {
            if (L.ax) {
                const double uw = (i - 1 > 0) ? u[n - 1] : 0.0;
                const double ue = (i + 1 < L.nx - 1) ? u[n + 1] : 0.0;
                Au -= L.cx * (uw + ue);
            }
            if (L.ay) {
                const double us = (j - 1 > 0) ? u[n - L.nx] : 0.0;
                const double un = (j + 1 < L.ny - 1) ? u[n + L.nx] : 0.0;
                Au -= L.cy * (us + un);
            }
            if (L.az) {
                const double ud = (k - 1 > 0) ? u[n - plane] : 0.0;
                const double uu = (k + 1 < L.nz - 1) ? u[n + plane] : 0.0;
                Au -= L.cz * (uu + ud);
            }
        }
        double o;
        if (MODE == ST_APPLY || MODE == ST_APPLY_DOT) {
            o = Au;
            if (MODE == ST_APPLY_DOT) dv[0] = uc * Au;
        } else if (MODE == ST_LIN_BU) {
            o = op.cb * uc + op.cg * (uc - Au);
        } else {
            const double bv = op.b[n];
            o = op.cb * uc + op.cg * (bv - Au);
            if (MODE == ST_LIN_PM1 || MODE == ST_LIN_PM1_DOT2) o += op.ca * op.pm1[n];
            if (MODE == ST_LIN_PM1_DOT2) { dv[0] = o * o; dv[1] = o * bv; }
        }
        op.out[n] = o;
        port_store(op.port, n, o);
    }
    port_signal(op.port, span.boundary, span.nboundary);
    if (MODE == ST_APPLY_DOT) {
        double d1[1] = {dv[0]};
        grid_sum_finalize<1, 256>(d1, partials, ticket, op.dot_out);
    }
    if (MODE == ST_LIN_PM1_DOT2) grid_sum_finalize<2, 256>(dv, partials, ticket, op.dot_out);
}

template <int MODE>
static int launch_generic(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red) {
    const long long nloc = L.nlocal();
    if (nloc <= 0) return 0;
    const long long nb = (nloc + 255) / 256;
    if ((MODE == ST_APPLY_DOT || MODE == ST_LIN_PM1_DOT2) && nb > red.max_blocks)
        return fail(63, "stencil dot: %lld blocks exceed the reducer scratch (%d)", nb, red.max_blocks);
    stencil_generic_kernel<MODE><<<(unsigned)nb, 256, 0, st>>>(L, op, red.partials, red.ticket);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_stencil_fast(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red);

int launch_stencil_generic(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red) {
    switch (op.mode) {
        case ST_APPLY: return launch_generic<ST_APPLY>(st, L, op, red);
        case ST_APPLY_DOT: return launch_generic<ST_APPLY_DOT>(st, L, op, red);
        case ST_LIN: return launch_generic<ST_LIN>(st, L, op, red);
        case ST_LIN_PM1: return launch_generic<ST_LIN_PM1>(st, L, op, red);
        case ST_LIN_BU: return launch_generic<ST_LIN_BU>(st, L, op, red);
        case ST_LIN_PM1_DOT2: return launch_generic<ST_LIN_PM1_DOT2>(st, L, op, red);
    }
    return fail(62, "unknown stencil mode %d", op.mode);
}

int launch_stencil(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red) {
    if (stencil_fast_eligible(L)) return launch_stencil_fast(st, L, op, red);
    return launch_stencil_generic(st, L, op, red);
}

}  // namespace p4b
