// assembled.cu -- assembled Jacobians of the 2-D DMDA drivers as device matrices, and their finite-difference assembly.
//
// [PETSc] SNESComputeJacobianDefaultColor (-snes_fd_color, c/ch7/makefile:16,25, c/ch8/cluster.sh:70) builds the
// Jacobian of FormFunctionLocal (c/ch7/minimal.c:210-282) column group by column group: the DMDA BOX stencil of width
// 1 is coloured with 3 x 3 = 9 colours, colour(i,j) = (i mod 3) + 3 (j mod 3), no row of the matrix meets two columns
// of one colour, so ONE residual evaluation with every node of a colour perturbed yields all those columns:
//     h      = sqrt(DBL_EPSILON) * sqrt(1 + ||x||_2)        (MatFDColoring's default differencing "wp": one step for
//     J_nm   = (F(x + h e_colour)_n - F(x)_n) * (1 / h)      every column; pinned by c/ch7/output/minimal.test1, see below)
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

// [PETSc] MatFDColoringApply with its default differencing "wp" (Walker-Pernice): ONE step for every column,
// h = sqrt(DBL_EPSILON) * sqrt(1 + ||u||_2), entries scaled by vscale = 1/h.  (Pinned by c/ch7/output/minimal.test1:
// every printed digit of its six residual norms is reproduced with this step and with no other candidate; see
// oracle/minimal_solver_oracle.py.)  ||u||_2 is a reduction: the caller takes it with the library's norm and passes h.
double fd_step_wp(double unorm) { return 1.4901161193847656e-08 * sqrt(1.0 + unorm); }

__global__ void __launch_bounds__(256) fd_perturb_kernel(int mx, int my, int ci, int cj, double h, const double *__restrict__ u,
                                                          double *__restrict__ up) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * my) return;
    const int j = n / mx, i = n - j * mx;
    const double x = u[n];
    up[n] = (i % 3 == ci && j % 3 == cj) ? x + h : x;
}

// the column of colour (ci, cj) that row n meets is its neighbour (i + di, j + dj) with (i+di) mod 3 = ci, ...
__global__ void __launch_bounds__(256) fd_extract_kernel(int mx, int my, int ci, int cj, double vscale,
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
    vals[(size_t)(3 * (dj + 1) + (di + 1)) * N + n] = (Fp[n] - F0[n]) * vscale;
}

// the two device halves of one colour, for callers whose residual is not a kernel of this library (host callbacks)
int launch_fd_perturb(cudaStream_t st, int mx, int my, int ci, int cj, double h, const double *u, double *up) {
    fd_perturb_kernel<<<(unsigned)((mx * my + 255) / 256), 256, 0, st>>>(mx, my, ci, cj, h, u, up);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_fd_extract(cudaStream_t st, int mx, int my, int ci, int cj, double h, const double *F0, const double *Fp,
                      double *vals) {
    fd_extract_kernel<<<(unsigned)((mx * my + 255) / 256), 256, 0, st>>>(mx, my, ci, cj, 1.0 / h, F0, Fp, vals);
    P4B_LAUNCH_CHECK();
    return 0;
}

int fd_jacobian_minimal(cudaStream_t st, int mx, int my, double q, double unorm, const double *u, const double *g,
                        const double *F0, double *vals, double *up, double *Fp) {
    const int N = mx * my;
    const unsigned nb = (unsigned)((N + 255) / 256);
    const double h = fd_step_wp(unorm), vscale = 1.0 / h;
    P4B_CUDA(cudaMemsetAsync(vals, 0, sizeof(double) * 9 * (size_t)N, st));
    for (int cj = 0; cj < 3; cj++)
        for (int ci = 0; ci < 3; ci++) {
            fd_perturb_kernel<<<nb, 256, 0, st>>>(mx, my, ci, cj, h, u, up);
            P4B_LAUNCH_CHECK();
            P4B_CHECK(launch_minimal_function(st, mx, my, 0, my, q, up, g, Fp));
            fd_extract_kernel<<<nb, 256, 0, st>>>(mx, my, ci, cj, vscale, F0, Fp, vals);
            P4B_LAUNCH_CHECK();
        }
    return 0;
}

// The matrix Poisson2DJacobianLocal inserts (c/ch6/poissonfunctions.c:152-193), in the stencil9 layout: minimal.c registers
// it as its Jacobian callback (c/ch7/minimal.c:142-145, "ONLY APPROXIMATE"), so it is Newton's matrix when neither
// -snes_fd_color nor -snes_mf_operator is given and the preconditioner's matrix under -snes_mf_operator.  Diagonal
// 2 (scx + scy) on every row, -scx / -scy to INTERIOR neighbours of interior rows only (boundary columns eliminated),
// scx = cx hy/hx, scy = cy hx/hy; the four corner planes are zero.
__global__ void __launch_bounds__(256) poisson_stencil9_kernel(int mx, int my, double scx, double scy,
                                                                double *__restrict__ vals) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int N = mx * my;
    if (n >= N) return;
    const int j = n / mx, i = n - j * mx;
    const bool in = i > 0 && i < mx - 1 && j > 0 && j < my - 1;
    double v[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    v[4] = 2.0 * (scx + scy);
    if (in) {
        if (i - 1 > 0) v[3] = -scx;
        if (i + 1 < mx - 1) v[5] = -scx;
        if (j - 1 > 0) v[1] = -scy;
        if (j + 1 < my - 1) v[7] = -scy;
    }
#pragma unroll
    for (int s = 0; s < 9; s++) vals[(size_t)s * N + n] = v[s];
}
int launch_poisson_stencil9(cudaStream_t st, int mx, int my, double Lx, double Ly, double cx, double cy, double *vals) {
    const double hx = Lx / (mx - 1), hy = Ly / (my - 1);
    poisson_stencil9_kernel<<<(unsigned)((mx * my + 255) / 256), 256, 0, st>>>(mx, my, cx * hy / hx, cy * hx / hy, vals);
    P4B_LAUNCH_CHECK();
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

// ---------------------------------------------------------------------------------------------------------
// pattern.c: the Jacobian of the implicit stage equation  F(t, Y, (Y - Y0)/dt) = G(t, Y)
//     J X = shift*X - C L9(X) - G'(Y) X       ([PETSc] TSComputeIJacobian: shift dF/dYdot + dF/dY, minus the RHS Jacobian)
// with L9 = [1 4 1; 4 -20 4; 1 4 1] periodic (c/ch5/pattern.c:274-318) and the 2 x 2 pointwise blocks of
// FormRHSJacobianLocal (:202-236) evaluated at the level's iterate Y (Y == nullptr: -ptn_no_rhsjacobian, the block
// is dropped).  Applied matrix-free: the operator is a constant stencil plus a pointwise block, storing it would
// only add traffic.  MODE 0: out = J X.   MODE 1: out = ca*pm1 + cb*X + cg*B(b - J X), B = diag(J)^-1 or I.
// MODE 2: out.x/.y = Gershgorin row ratios sum_j |J_nj| / |J_nn| of the two rows of the node.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) pattern_jac_kernel(int mx, int my, double Cu, double Cv, double shift, double phi,
                                                           double kappa, const double2 *__restrict__ Y,
                                                           const double2 *__restrict__ X, const double2 *b,
                                                           const double2 *pm1, double ca, double cb, double cg,
                                                           int jacobi, int ywrap, double2 *out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * my) return;
    const int j = n / mx, i = n - j * mx;
    // the RHS Jacobian block [[-v^2 - phi, -2uv], [v^2, 2uv - (phi + kappa)]] at this node
    double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
    if (Y) {
        const double2 y = Y[n];
        const double uv = y.x * y.y, v2 = y.y * y.y;
        g00 = -v2 - phi;  g01 = -2.0 * uv;  g10 = v2;  g11 = 2.0 * uv - (phi + kappa);
    }
    const double du = shift + 20.0 * Cu - g00, dv = shift + 20.0 * Cv - g11;      // diagonal of J
    if (MODE == 2) {
        out[n] = make_double2((fabs(du) + 20.0 * Cu + fabs(g01)) / fabs(du), (fabs(dv) + 20.0 * Cv + fabs(g10)) / fabs(dv));
        return;
    }
    const int iw = (i == 0) ? mx - 1 : i - 1, ie = (i == mx - 1) ? 0 : i + 1;      // periodic wrap
    // ywrap = 0: a y-slab with its ghost rows -1 and my in memory (mp_kernels.cu pattern_ifunction_kernel)
    const int js = (j == 0 && ywrap) ? my - 1 : j - 1, jn = (j == my - 1 && ywrap) ? 0 : j + 1;
    const double2 c = X[n];
    const double2 nw = X[jn * mx + iw], nn = X[jn * mx + i], ne = X[jn * mx + ie];
    const double2 ww = X[j * mx + iw], ee = X[j * mx + ie];
    const double2 sw = X[js * mx + iw], ss = X[js * mx + i], se = X[js * mx + ie];
    const double lapu = nw.x + 4.0 * nn.x + ne.x + 4.0 * ww.x - 20.0 * c.x + 4.0 * ee.x + sw.x + 4.0 * ss.x + se.x;
    const double lapv = nw.y + 4.0 * nn.y + ne.y + 4.0 * ww.y - 20.0 * c.y + 4.0 * ee.y + sw.y + 4.0 * ss.y + se.y;
    const double Ju = shift * c.x - Cu * lapu - (g00 * c.x + g01 * c.y);
    const double Jv = shift * c.y - Cv * lapv - (g10 * c.x + g11 * c.y);
    if (MODE == 0) {
        out[n] = make_double2(Ju, Jv);
    } else {
        const double2 bb = b ? b[n] : make_double2(0.0, 0.0);
        double ru = bb.x - Ju, rv = bb.y - Jv;
        if (jacobi) { ru /= du; rv /= dv; }
        double ou = cb * c.x + cg * ru, ov = cb * c.y + cg * rv;
        if (pm1) { const double2 p = pm1[n]; ou += ca * p.x; ov += ca * p.y; }
        out[n] = make_double2(ou, ov);
    }
}

// A^-1 from banded LU factors (nk_solver.hpp stencil9_band_lu): thread `col` solves L U x = e_col in place in column
// `col` of the row-major inverse.  All threads of a warp touch the same row r of consecutive columns at the same time
// (coalesced), the band entry they multiply with is a broadcast.  n = 1089, bw = 34 (the 33 x 33 base grid of
// c/ch8/cluster.sh:70): ~0.3 ms against ~50 ms for the same loops on the host.
__global__ void __launch_bounds__(128) band_inverse_kernel(int n, int bw, const double *__restrict__ B, double *Ainv) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= n) return;
    const int W = 2 * bw + 1;
    double *x = Ainv + col;                                    // x[r] = Ainv[r * n + col]
    for (int r = 0; r < n; r++) x[(size_t)r * n] = (r == col) ? 1.0 : 0.0;
    for (int r = col + 1; r < n; r++) {                        // forward: L y = e_col (y_r = 0 for r < col)
        double s = 0.0;
        const int c0 = max(col, r - bw);
        for (int c = c0; c < r; c++) s += B[(size_t)r * W + (c - r + bw)] * x[(size_t)c * n];
        x[(size_t)r * n] -= s;
    }
    for (int r = n - 1; r >= 0; r--) {                         // backward: U x = y
        double s = x[(size_t)r * n];
        const int c1 = min(n - 1, r + bw);
        for (int c = r + 1; c <= c1; c++) s -= B[(size_t)r * W + (c - r + bw)] * x[(size_t)c * n];
        x[(size_t)r * n] = s / B[(size_t)r * W + bw];
    }
}
int launch_band_inverse(cudaStream_t st, int n, int bw, const double *B, double *Ainv) {
    band_inverse_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, bw, B, Ainv);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_pattern_jac(cudaStream_t st, int mode, int mx, int my, double Cu, double Cv, double shift, double phi,
                       double kappa, const double *Y, const double *X, const double *b, const double *pm1, double ca,
                       double cb, double cg, int jacobi, double *out, int ywrap) {
    const int N = mx * my;
    if (N <= 0) return 0;
    const unsigned nb = (unsigned)((N + 255) / 256);
    const double2 *Y2 = reinterpret_cast<const double2 *>(Y), *X2 = reinterpret_cast<const double2 *>(X);
    const double2 *b2 = reinterpret_cast<const double2 *>(b), *p2 = reinterpret_cast<const double2 *>(pm1);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    if (mode == 0) pattern_jac_kernel<0><<<nb, 256, 0, st>>>(mx, my, Cu, Cv, shift, phi, kappa, Y2, X2, b2, p2, ca, cb, cg, jacobi, ywrap, o2);
    else if (mode == 1) pattern_jac_kernel<1><<<nb, 256, 0, st>>>(mx, my, Cu, Cv, shift, phi, kappa, Y2, X2, b2, p2, ca, cb, cg, jacobi, ywrap, o2);
    else pattern_jac_kernel<2><<<nb, 256, 0, st>>>(mx, my, Cu, Cv, shift, phi, kappa, Y2, X2, b2, p2, ca, cb, cg, jacobi, ywrap, o2);
    P4B_LAUNCH_CHECK();
    return 0;
}

// [PETSc] DMCreateInterpolation on a PERIODIC 2-dof DMDA (ratio 2, fine m = 2 M): fine node 2I coincides with coarse I,
// fine node 2I+1 averages coarse I and (I+1) mod M; tensor product in x and y, the same for both components.
// mode 0: restriction b_c = P^T r      mode 1: prolongation x_f += P x_c      mode 2: injection y_c(I,J) = y_f(2I,2J)
// ywrap = 0: y-slabs of a multi-GPU run (My coarse rows, 2 My fine rows of this rank): the restriction reads the fine
// ghost row -1, the prolongation the coarse ghost row My, both in memory; no wrap in y.
__global__ void __launch_bounds__(256) pattern_transfer_kernel(int mode, int Mx, int My, int ywrap,
                                                                const double2 *__restrict__ src,
                                                                double2 *__restrict__ dst) {
    const int fx = 2 * Mx, fy = 2 * My;
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (mode == 1) {
        if (n >= fx * fy) return;
        const int j = n / fx, i = n - j * fx;
        const int I0 = i >> 1, I1 = (i & 1) ? (I0 + 1 == Mx ? 0 : I0 + 1) : I0;
        const int J0 = j >> 1, J1 = (j & 1) ? ((J0 + 1 == My && ywrap) ? 0 : J0 + 1) : J0;
        const double2 a = src[J0 * Mx + I0], b = src[J0 * Mx + I1], c = src[J1 * Mx + I0], d = src[J1 * Mx + I1];
        double2 o = dst[n];
        o.x += 0.25 * ((a.x + b.x) + (c.x + d.x));
        o.y += 0.25 * ((a.y + b.y) + (c.y + d.y));
        dst[n] = o;
        return;
    }
    if (n >= Mx * My) return;
    const int J = n / Mx, I = n - J * Mx;
    if (mode == 2) {
        dst[n] = src[(2 * J) * fx + 2 * I];
        return;
    }
    double su = 0.0, sv = 0.0;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
        int jf = 2 * J + dj;
        if (ywrap) jf = jf < 0 ? jf + fy : (jf >= fy ? jf - fy : jf);
        const double wj = dj ? 0.5 : 1.0;
        double ru = 0.0, rv = 0.0;
#pragma unroll
        for (int di = -1; di <= 1; di++) {
            int i_f = 2 * I + di;
            i_f = i_f < 0 ? i_f + fx : (i_f >= fx ? i_f - fx : i_f);
            const double2 v = src[jf * fx + i_f];
            const double wi = di ? 0.5 : 1.0;
            ru += wi * v.x;
            rv += wi * v.y;
        }
        su += wj * ru;
        sv += wj * rv;
    }
    dst[n] = make_double2(su, sv);
}

int launch_pattern_transfer(cudaStream_t st, int mode, int Mx, int My, const double *src, double *dst, int ywrap) {
    const int N = (mode == 1) ? 4 * Mx * My : Mx * My;
    if (N <= 0) return 0;
    pattern_transfer_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mode, Mx, My, ywrap,
                                                                       reinterpret_cast<const double2 *>(src),
                                                                       reinterpret_cast<double2 *>(dst));
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
