// mp_kernels.cu -- device versions of the remaining DMDA callbacks on the BASELINE path, and the assembled-matrix SpMV:
//
//   minimal_function_kernel   c/ch7/minimal.c:210-282  FormFunctionLocal (9-point BOX residual of the minimal
//                             surface equation, diffusivity (1+|grad u|^2)^q at the four staggered points)     [K7]
//   minimal_sample_kernel     c/ch7/minimal.c:27-42    g_bdry_tent / g_bdry_catenoid at every node
//   pattern_rhs_kernel        c/ch5/pattern.c:185-199  FormRHSFunctionLocal (pointwise Gray-Scott reaction)      [K8]
//   pattern_ifunction_kernel  c/ch5/pattern.c:242-267  FormIFunctionLocal  F = Ydot - C L9(Y), periodic, (u,v)
//                             interleaved; with Ydot := shift*Y it is the action of FormIJacobianLocal (:274-318)
//   pattern_init_kernel       c/ch5/pattern.c:146-179  InitialState (no noise)
//   heat_rhs_kernel           c/ch5/heat.c:141-163     FormRHSFunctionLocal (5-point Laplacian, Neumann in x through
//                             mirrored ghosts, periodic in y, source f); without the data: the action of :166-208
//   sell_spmv_kernel          [PETSc] MatMult_SeqAIJ for assembled Jacobians, as SELL-32 (sliced ELLPACK)       [K6]
//
// All fp64 and HBM-bound; none is a dense contraction, so no tensor cores.
#include <vector>

#include "kernels.h"

namespace p4b {

// ---------------------------------------------------------------------------------------------- minimal.c
__global__ void __launch_bounds__(256) minimal_sample_kernel(int mx, int my, int zs, int zm, int problem, double tent_H,
                                                              double c, double *__restrict__ g) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= (long long)mx * zm) return;
    const int jl = (int)(n / mx), i = (int)(n - (long long)jl * mx), j = jl + zs;
    const double x = i * (1.0 / (mx - 1)), y = j * (1.0 / (my - 1));
    double v;
    if (problem == 0) v = (x < 1.0e-8) ? 2.0 * tent_H * (y < 0.5 ? y : 1.0 - y) : 0.0;
    else v = c * cosh(x / c) * sin(acos((y / c) / cosh(x / c)));
    g[n] = v;
}

// DD(w) = (1+w)^q (minimal.c:46-48).  The reference calls libm pow(); fp64 pow costs ~10x the rest of the
// residual, so the exponents the drivers actually use are special-cased (q = -1/2: the minimal surface
// equation, q = 0: Laplace) and the general case is exp(q log(1+w)); all agree with pow() to a few ulp.
template <int QMODE>
__device__ __forceinline__ double diffusivity(double w, double q) {
    if (QMODE == 0) return 1.0;
    if (QMODE == 1) return rsqrt(1.0 + w);
    return exp(q * log(1.0 + w));
}

// u, g: rows [zs-1, zs+zm] readable (ghost rows) when the slab is interior
template <int QMODE>
__global__ void __launch_bounds__(256) minimal_function_kernel(int mx, int my, int zs, int zm, double q,
                                                                const double *__restrict__ u,
                                                                const double *__restrict__ g,
                                                                double *__restrict__ FF) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= (long long)mx * zm) return;
    const int jl = (int)(n / mx), i = (int)(n - (long long)jl * mx), j = jl + zs;
    const double uc = u[n];
    if (j == 0 || i == 0 || i == mx - 1 || j == my - 1) {
        FF[n] = uc - g[n];                                   // unscaled boundary rows (:227)
        return;
    }
    const bool be = (i + 1 == mx - 1), bw = (i - 1 == 0), bn = (j + 1 == my - 1), bs = (j - 1 == 0);
    // a neighbour that is a boundary node takes the boundary value g (:230-256)
    const double ue = be ? g[n + 1] : u[n + 1];
    const double uw = bw ? g[n - 1] : u[n - 1];
    const double un = bn ? g[n + mx] : u[n + mx];
    const double us = bs ? g[n - mx] : u[n - mx];
    const double une = (be || bn) ? g[n + mx + 1] : u[n + mx + 1];
    const double unw = (bw || bn) ? g[n + mx - 1] : u[n + mx - 1];
    const double use = (be || bs) ? g[n - mx + 1] : u[n - mx + 1];
    const double usw = (bw || bs) ? g[n - mx - 1] : u[n - mx - 1];
    // reciprocal spacings instead of the reference's divisions (a few ulp; fp64 division is ~10x a multiply)
    const double ihx = (double)(mx - 1), ihy = (double)(my - 1), i4hx = 0.25 * ihx, i4hy = 0.25 * ihy;
    double dux, duy;
    dux = (ue - uc) * ihx;  duy = (un + une - us - use) * i4hy;
    const double De = diffusivity<QMODE>(dux * dux + duy * duy, q);
    dux = (uc - uw) * ihx;  duy = (unw + un - usw - us) * i4hy;
    const double Dw = diffusivity<QMODE>(dux * dux + duy * duy, q);
    dux = (ue + une - uw - unw) * i4hx;  duy = (un - uc) * ihy;
    const double Dn = diffusivity<QMODE>(dux * dux + duy * duy, q);
    dux = (ue + use - uw - usw) * i4hx;  duy = (uc - us) * ihy;
    const double Ds = diffusivity<QMODE>(dux * dux + duy * duy, q);
    const double hyhx = ihx / ihy, hxhy = ihy / ihx;        // hy/hx and hx/hy
    FF[n] = -hyhx * (De * (ue - uc) - Dw * (uc - uw)) - hxhy * (Dn * (un - uc) - Ds * (uc - us));
}

int launch_minimal_sample(cudaStream_t st, int mx, int my, int zs, int zm, int problem, double tent_H, double c, double *g) {
    const long long n = (long long)mx * zm;
    if (n <= 0) return 0;
    minimal_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mx, my, zs, zm, problem, tent_H, c, g);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_minimal_function(cudaStream_t st, int mx, int my, int zs, int zm, double q, const double *u, const double *g,
                            double *FF) {
    const long long n = (long long)mx * zm;
    if (n <= 0) return 0;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (q == 0.0) minimal_function_kernel<0><<<nb, 256, 0, st>>>(mx, my, zs, zm, q, u, g, FF);
    else if (q == -0.5) minimal_function_kernel<1><<<nb, 256, 0, st>>>(mx, my, zs, zm, q, u, g, FF);
    else minimal_function_kernel<2><<<nb, 256, 0, st>>>(mx, my, zs, zm, q, u, g, FF);
    P4B_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------- pattern.c
// (rows [ys, ys + ym) of the my-row grid: a y-slab of a multi-GPU run; ys = 0, ym = my is the whole grid)
__global__ void __launch_bounds__(256) pattern_init_kernel(int mx, int my, int ys, int ym, double L,
                                                            const double2 *__restrict__ noise, double level,
                                                            double2 *__restrict__ Y) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * ym) return;
    const int jl = n / mx, i = n - jl * mx, j = jl + ys;
    const double x = i * (L / mx), y = j * (L / my);
    const double ledge = (L - 0.5) / 2.0, redge = L - ledge;
    const double PI = 3.14159265358979323846264338327950288;
    // pattern.c:159-165: Y = level * random, then v += patch, u += 1 - 2 v (the v that already carries its noise)
    double u = 0.0, v = 0.0;
    if (noise) { const double2 r = noise[n]; u = level * r.x; v = level * r.y; }
    if (x >= ledge && x <= redge && y >= ledge && y <= redge) {
        const double sx = sin(4.0 * PI * x), sy = sin(4.0 * PI * y);
        v += 0.5 * sx * sx * sy * sy;
    }
    Y[n] = make_double2(u + (1.0 - 2.0 * v), v);
}

__global__ void __launch_bounds__(256) pattern_rhs_kernel(int n, double phi, double kappa, const double2 *__restrict__ Y,
                                                           double2 *__restrict__ G) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= n) return;
    const double2 y = Y[q];
    const double uv2 = y.x * y.y * y.y;
    G[q] = make_double2(-uv2 + phi * (1.0 - y.x), uv2 - (phi + kappa) * y.y);
}

// F = Ydot - C L9(Y)  (use_shift = 0)   or   J X = shift*X - C L9(X)  (use_shift = 1, Ydot ignored)
// ywrap = 1: the my rows are the whole periodic grid.  ywrap = 0: they are a y-slab whose ghost rows -1 and my are in
// memory on both sides (multi-GPU; filled by the ring exchange), so there is no wrap in y.
__global__ void __launch_bounds__(256) pattern_ifunction_kernel(int mx, int my, double Cu, double Cv, int use_shift,
                                                                 double shift, int ywrap, const double2 *__restrict__ Y,
                                                                 const double2 *__restrict__ Ydot,
                                                                 double2 *__restrict__ F) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * my) return;
    const int j = n / mx, i = n - j * mx;
    const int iw = (i == 0) ? mx - 1 : i - 1, ie = (i == mx - 1) ? 0 : i + 1;      // periodic wrap
    const int js = (j == 0 && ywrap) ? my - 1 : j - 1, jn = (j == my - 1 && ywrap) ? 0 : j + 1;
    const double2 c = Y[n];
    const double2 nw = Y[jn * mx + iw], nn = Y[jn * mx + i], ne = Y[jn * mx + ie];
    const double2 ww = Y[j * mx + iw], ee = Y[j * mx + ie];
    const double2 sw = Y[js * mx + iw], ss = Y[js * mx + i], se = Y[js * mx + ie];
    const double lapu = nw.x + 4.0 * nn.x + ne.x + 4.0 * ww.x - 20.0 * c.x + 4.0 * ee.x + sw.x + 4.0 * ss.x + se.x;
    const double lapv = nw.y + 4.0 * nn.y + ne.y + 4.0 * ww.y - 20.0 * c.y + 4.0 * ee.y + sw.y + 4.0 * ss.y + se.y;
    double2 d;
    if (use_shift) d = make_double2(shift * c.x, shift * c.y);
    else d = Ydot[n];
    F[n] = make_double2(d.x - Cu * lapu, d.y - Cv * lapv);
}

int launch_pattern_init(cudaStream_t st, int mx, int my, double L, double *Y, const double *noise, double level, int ys,
                        int ym) {
    if (ym < 0) ym = my;
    pattern_init_kernel<<<(mx * ym + 255) / 256, 256, 0, st>>>(mx, my, ys, ym, L, reinterpret_cast<const double2 *>(noise),
                                                               level, reinterpret_cast<double2 *>(Y));
    P4B_LAUNCH_CHECK();
    return 0;
}
// ---------------------------------------------------------------------------------------------- heat.c
// MODE 0: G = D0 (uxx + uyy) + f(x, y), with ul = u[i+1] + 2 hx gamma(y) at i = 0 and ur = u[i-1] at i = mx-1
//         (c/ch5/heat.c:141-163; f_source :16-19, gamma_neumann :21-23; hx = 1/(mx-1), hy = 1/my, :99-103)
// MODE 1: out = shift u - D0 (uxx + uyy) with the homogeneous mirror conditions: the stage operator shift I - dG/du,
//         i.e. the rows FormRHSJacobianLocal inserts (:166-208), applied without being stored
template <int MODE>
__global__ void __launch_bounds__(256) heat_rhs_kernel(int mx, int my, double D0, double shift, const double *__restrict__ u,
                                                        double *__restrict__ out) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= mx * my) return;
    const int j = n / mx, i = n - j * mx;
    const double hx = 1.0 / (double)(mx - 1), hy = 1.0 / (double)my;
    const double x = hx * i, y = hy * j;
    const double c = u[n];
    double ul, ur;
    if (i == 0) {
        ul = u[n + 1];
        if (MODE == 0) ul += 2.0 * hx * sin(6.0 * M_PI * y);
    } else ul = u[n - 1];
    ur = (i == mx - 1) ? u[n - 1] : u[n + 1];
    const int js = (j == 0) ? my - 1 : j - 1, jn = (j == my - 1) ? 0 : j + 1;
    const double uxx = (ul - 2.0 * c + ur) / (hx * hx);
    const double uyy = (u[js * mx + i] - 2.0 * c + u[jn * mx + i]) / (hy * hy);
    if (MODE == 0) out[n] = D0 * (uxx + uyy) + 3.0 * exp(-25.0 * (x - 0.6) * (x - 0.6)) * sin(2.0 * M_PI * y);
    else out[n] = shift * c - D0 * (uxx + uyy);
}
int launch_heat_rhs(cudaStream_t st, int mx, int my, double D0, const double *u, double *G) {
    heat_rhs_kernel<0><<<(mx * my + 255) / 256, 256, 0, st>>>(mx, my, D0, 0.0, u, G);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_heat_jac_apply(cudaStream_t st, int mx, int my, double D0, double shift, const double *X, double *out) {
    heat_rhs_kernel<1><<<(mx * my + 255) / 256, 256, 0, st>>>(mx, my, D0, shift, X, out);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_pattern_rhs(cudaStream_t st, int n, double phi, double kappa, const double *Y, double *G) {
    pattern_rhs_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, phi, kappa, reinterpret_cast<const double2 *>(Y),
                                                          reinterpret_cast<double2 *>(G));
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_pattern_ifunction(cudaStream_t st, int mx, int my, double Cu, double Cv, int use_shift, double shift,
                             const double *Y, const double *Ydot, double *F, int ywrap) {
    pattern_ifunction_kernel<<<(mx * my + 255) / 256, 256, 0, st>>>(
        mx, my, Cu, Cv, use_shift, shift, ywrap, reinterpret_cast<const double2 *>(Y),
        reinterpret_cast<const double2 *>(Ydot), reinterpret_cast<double2 *>(F));
    P4B_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------- SELL-32 SpMV
// Rows are grouped in slices of 32; inside a slice the entries are stored column-major (entry k of the 32 rows
// is contiguous), padded to the longest row of the slice with (col = row, val = 0).  One thread per row: every
// load of vals/cols is a fully coalesced 256/128-byte access, x is gathered through L1/L2.
// Algorithmic bytes: 12 * padded_nnz + 8 N (y) + 8 N (x once) ~ (12 k + 16) N for k entries per row.
__global__ void __launch_bounds__(256) sell_spmv_kernel(int nrows, const long long *__restrict__ slice_ptr,
                                                         const int *__restrict__ cols, const double *__restrict__ vals,
                                                         const double *__restrict__ x, double *__restrict__ y) {
    const int row = blockIdx.x * 256 + threadIdx.x;
    if (row >= nrows) return;
    const int slice = row >> 5, lane = row & 31;
    const long long beg = slice_ptr[slice], end = slice_ptr[slice + 1];
    double s = 0.0;
    for (long long p = beg + lane; p < end; p += 32) s += vals[p] * x[cols[p]];
    y[row] = s;
}

struct Sell {
    int nrows = 0;
    long long padded = 0, nnz = 0;
    long long *slice_ptr = nullptr;
    int *cols = nullptr;
    double *vals = nullptr;
};

int sell_build(cudaStream_t st, int nrows, const int *rowptr, const int *colind, const double *vals, Sell **out) {
    if (nrows < 0 || !rowptr || !colind || !vals) return fail(62, "sell_build: null CSR arrays");
    const int nsl = (nrows + 31) / 32;
    std::vector<long long> sp(nsl + 1, 0);
    for (int s = 0; s < nsl; s++) {
        int w = 0;
        for (int r = s * 32; r < nrows && r < s * 32 + 32; r++) w = std::max(w, rowptr[r + 1] - rowptr[r]);
        sp[s + 1] = sp[s] + 32LL * w;
    }
    const long long padded = sp[nsl];
    std::vector<int> c((size_t)padded);
    std::vector<double> v((size_t)padded, 0.0);
    for (int s = 0; s < nsl; s++) {
        const int w = (int)((sp[s + 1] - sp[s]) / 32);
        for (int l = 0; l < 32; l++) {
            const int r = s * 32 + l;
            for (int k = 0; k < w; k++) {
                const long long p = sp[s] + 32LL * k + l;
                if (r < nrows && k < rowptr[r + 1] - rowptr[r]) {
                    c[p] = colind[rowptr[r] + k];
                    v[p] = vals[rowptr[r] + k];
                } else {
                    c[p] = r < nrows ? r : 0;
                }
            }
        }
    }
    Sell *A = new Sell();
    A->nrows = nrows;
    A->padded = padded;
    A->nnz = rowptr[nrows];
    P4B_CUDA(cudaMalloc(&A->slice_ptr, sizeof(long long) * (nsl + 1)));
    P4B_CUDA(cudaMalloc(&A->cols, sizeof(int) * (size_t)std::max(padded, 1LL)));
    P4B_CUDA(cudaMalloc(&A->vals, sizeof(double) * (size_t)std::max(padded, 1LL)));
    P4B_CUDA(cudaMemcpyAsync(A->slice_ptr, sp.data(), sizeof(long long) * (nsl + 1), cudaMemcpyHostToDevice, st));
    P4B_CUDA(cudaMemcpyAsync(A->cols, c.data(), sizeof(int) * (size_t)padded, cudaMemcpyHostToDevice, st));
    P4B_CUDA(cudaMemcpyAsync(A->vals, v.data(), sizeof(double) * (size_t)padded, cudaMemcpyHostToDevice, st));
    P4B_CUDA(cudaStreamSynchronize(st));
    *out = A;
    return 0;
}
int sell_spmv(cudaStream_t st, const Sell *A, const double *x, double *y) {
    if (A->nrows == 0) return 0;
    sell_spmv_kernel<<<(A->nrows + 255) / 256, 256, 0, st>>>(A->nrows, A->slice_ptr, A->cols, A->vals, x, y);
    P4B_LAUNCH_CHECK();
    return 0;
}
void sell_free(Sell *A) {
    if (!A) return;
    cudaFree(A->slice_ptr);
    cudaFree(A->cols);
    cudaFree(A->vals);
    delete A;
}
void sell_info(const Sell *A, int *nrows, long long *nnz, long long *padded) {
    if (nrows) *nrows = A->nrows;
    if (nnz) *nnz = A->nnz;
    if (padded) *padded = A->padded;
}

}  // namespace p4b
