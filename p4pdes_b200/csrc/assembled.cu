// assembled.cu -- assembled Jacobians of the 2-D DMDA drivers as device matrices, and their finite-difference assembly.
//
// [PETSc] SNESComputeJacobianDefaultColor (-snes_fd_color, c/ch7/makefile:16,25, c/ch8/cluster.sh:70) builds the
// Jacobian of FormFunctionLocal (c/ch7/minimal.c:210-282) column group by column group: the DMDA BOX stencil of width
// 1 is coloured with 3 x 3 = 9 colours, colour(i,j) = (i mod 3) + 3 (j mod 3), no row of the matrix meets two columns
// of one colour, so ONE residual evaluation with every node of a colour perturbed yields all those columns:
//     dx_m   = eps * x_m            if |x_m| >= umin,   eps * umin * sign(x_m) otherwise   (eps = sqrt(DBL_EPSILON),
//     J_nm   = (F(x + dx e_colour)_n - F(x)_n) / dx_m                                       umin = 1e-6: MatFDColoring "ds")
// Here that is nine launches of the device residual plus a perturb and an extract kernel each; nothing leaves HBM.
//
// The matrix is kept in the layout a structured-grid Jacobian has on a GPU ("stencil9"): nine coefficient planes
// vals[s*N + n], s = 3 (dj+1) + (di+1), n = j mx + i -- 72 B/row of coefficients and perfectly coalesced, against
// 108 B/row for column-indexed CSR/SELL.  y = A x and the fused smoother step out = ca*pm1 + cb*u + cg*B(b - A u)
// (B = I or diag(A)^-1: Richardson / Chebyshev + Jacobi, [PETSc] KSPSolve_Chebyshev + PCApply_Jacobi) stream it once.
// The column-indexed SELL-32 copy ([PETSc] MATSELL / MatMult_SeqAIJ, p4b_sell_*) of the same matrix is built by the
// host side from these planes when a caller wants an AIJ-style Mat (p4pdes_b200/minimal.py:stencil9_to_csr).
#include <float.h>
#include <math.h>

#include "kernels.h"

namespace p4b {

__device__ __forceinline__ double fd_dx(double x) {
    const double eps = 1.4901161193847656e-08, umin = 1.0e-6;      // sqrt(DBL_EPSILON), MatFDColoring defaults
    double dx = x;
    if (fabs(dx) < umin) dx = (dx < 0.0 ? -1.0 : 1.0) * umin;
    return dx * eps;
}

__global__ void __launch_bounds__(256) fd_perturb_kernel(int mx, int my, int ci, int cj, const double *__restrict__ u,
                                                          double *__restrict__ up) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * my) return;
    const int j = n / mx, i = n - j * mx;
    const double x = u[n];
    up[n] = (i % 3 == ci && j % 3 == cj) ? x + fd_dx(x) : x;
}

// the column of colour (ci, cj) that row n meets is its neighbour (i + di, j + dj) with (i+di) mod 3 = ci, ...
__global__ void __launch_bounds__(256) fd_extract_kernel(int mx, int my, int ci, int cj, const double *__restrict__ u,
                                                          const double *__restrict__ F0, const double *__restrict__ Fp,
                                                          double *__restrict__ vals) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int N = mx * my;
    if (n >= N) return;
    const int j = n / mx, i = n - j * mx;
    int di = ci - i % 3, dj = cj - j % 3;           // in {-2..2}; bring into {-1, 0, 1}
    if (di > 1) di -= 3;
    if (di < -1) di += 3;
    if (dj > 1) dj -= 3;
    if (dj < -1) dj += 3;
    const int ii = i + di, jj = j + dj;
    if (ii < 0 || ii >= mx || jj < 0 || jj >= my) return;
    const int m = jj * mx + ii;
    const double vscale = 1.0 / fd_dx(u[m]);
    vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n] = (Fp[n] - F0[n]) * vscale;
}

int fd_jacobian_minimal(cudaStream_t st, int mx, int my, double q, const double *u, const double *g, const double *F0,
                        double *vals, double *up, double *Fp) {
    const int N = mx * my;
    const unsigned nb = (unsigned)((N + 255) / 256);
    P4B_CUDA(cudaMemsetAsync(vals, 0, sizeof(double) * 9 * (size_t)N, st));
    for (int cj = 0; cj < 3; cj++)
        for (int ci = 0; ci < 3; ci++) {
            fd_perturb_kernel<<<nb, 256, 0, st>>>(mx, my, ci, cj, u, up);
            P4B_LAUNCH_CHECK();
            P4B_CHECK(launch_minimal_function(st, mx, my, 0, my, q, up, g, Fp));
            fd_extract_kernel<<<nb, 256, 0, st>>>(mx, my, ci, cj, u, F0, Fp, vals);
            P4B_LAUNCH_CHECK();
        }
    return 0;
}

// MODE 0: out = A u          MODE 1: out = ca*pm1 + cb*u + cg*B(b - A u), B = 1/diag when jacobi else 1
// (pm1 may be null (ca ignored) and may alias out; b may be null (treated as zero))
template <int MODE>
__global__ void __launch_bounds__(256) stencil9_kernel(int mx, int my, const double *__restrict__ vals,
                                                        const double *__restrict__ u, const double *b, const double *pm1,
                                                        double ca, double cb, double cg, int jacobi, double *out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int N = mx * my;
    if (n >= N) return;
    const int j = n / mx, i = n - j * mx;
    double Au = 0.0, diag = 1.0;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
#pragma unroll
        for (int di = -1; di <= 1; di++) {
            const int s = 3 * (dj + 1) + (di + 1);
            const double a = vals[(size_t)s * N + n];
            if (s == 4) diag = a;
            const int ii = i + di, jj = j + dj;
            if (ii >= 0 && ii < mx && jj >= 0 && jj < my) Au += a * u[n + dj * mx + di];
        }
    }
    if (MODE == 0) {
        out[n] = Au;
    } else {
        double r = (b ? b[n] : 0.0) - Au;
        if (jacobi) r /= diag;
        double o = cb * u[n] + cg * r;
        if (pm1) o += ca * pm1[n];
        out[n] = o;
    }
}

int launch_stencil9_apply(cudaStream_t st, int mx, int my, const double *vals, const double *x, double *y) {
    const int N = mx * my;
    if (N <= 0) return 0;
    stencil9_kernel<0><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mx, my, vals, x, nullptr, nullptr, 0, 0, 0, 0, y);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_stencil9_lin(cudaStream_t st, int mx, int my, const double *vals, const double *u, const double *b,
                        const double *pm1, double ca, double cb, double cg, int jacobi, double *out) {
    const int N = mx * my;
    if (N <= 0) return 0;
    stencil9_kernel<1><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mx, my, vals, u, b, pm1, ca, cb, cg, jacobi, out);
    P4B_LAUNCH_CHECK();
    return 0;
}

// out[n] = sum_s |a_ns| / |a_nn|: its maximum is the Gershgorin bound of lambda_max(D^-1 A) that sets the Chebyshev
// targets ([PETSc] estimates the same quantity with GMRES on a random right-hand side)
__global__ void __launch_bounds__(256) stencil9_rowratio_kernel(int mx, int my, const double *__restrict__ vals,
                                                                 double *__restrict__ out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int N = mx * my;
    if (n >= N) return;
    const int j = n / mx, i = n - j * mx;
    double s = 0.0;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++)
#pragma unroll
        for (int di = -1; di <= 1; di++) {
            const int ii = i + di, jj = j + dj;
            if (ii >= 0 && ii < mx && jj >= 0 && jj < my) s += fabs(vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n]);
        }
    out[n] = s / fabs(vals[(size_t)4 * N + n]);
}
int launch_stencil9_rowratio(cudaStream_t st, int mx, int my, const double *vals, double *out) {
    const int N = mx * my;
    if (N <= 0) return 0;
    stencil9_rowratio_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mx, my, vals, out);
    P4B_LAUNCH_CHECK();
    return 0;
}

// [PETSc] DMCreateInjection on a DMDA: coarse node (I, J) takes fine node (2I, 2J) (the iterate the coarse-level
// Jacobians are evaluated at)
__global__ void __launch_bounds__(256) inject2d_kernel(int cmx, int cmy, int fmx, const double *__restrict__ uf,
                                                        double *__restrict__ uc) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= cmx * cmy) return;
    const int J = n / cmx, I = n - J * cmx;
    uc[n] = uf[(size_t)(2 * J) * fmx + 2 * I];
}
int launch_inject2d(cudaStream_t st, int cmx, int cmy, int fmx, const double *uf, double *uc) {
    const int N = cmx * cmy;
    if (N <= 0) return 0;
    inject2d_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(cmx, cmy, fmx, uf, uc);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
