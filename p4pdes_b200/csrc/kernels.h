// kernels.h -- host-side launchers of the p4b200 CUDA kernels (internal; the C ABI is p4b200.h).
#pragma once
#include "common.cuh"
#include "comm.h"

namespace p4b {

// scratch for deterministic grid-wide reductions (owned by the context)
struct Reducer {
    double *partials = nullptr;       // 4 * max_blocks doubles
    unsigned int *ticket = nullptr;   // zero-initialised, re-armed by the kernels
    int max_blocks = 0;
};

// The stencil family  (A = the Jacobian of poissonfunctions.c:117-258 on this level):
//   APPLY      out = A u                                   [+ dot(u, A u) -> dot_out]
//   LIN        out = cb*u + cg*(b - A u)
//   LIN_PM1    out = ca*pm1 + cb*u + cg*(b - A u)          (out may alias pm1)
//   LIN_BU     out = cb*u + cg*(u - A u)                   (b == u, loaded once)
//   LIN_PM1_DOT2  LIN_PM1 plus (out,out) -> dot_out[0], (out,b) -> dot_out[1]   (the CG scalars of KSPSolve_CG
//              when `out` is z = M^-1 r and b is r: the last smoother step of the cycle on the finest level)
enum StencilMode { ST_APPLY = 0, ST_APPLY_DOT, ST_LIN, ST_LIN_PM1, ST_LIN_BU, ST_LIN_PM1_DOT2 };

struct StencilOp {
    int mode;
    const double *u;      // stencil operand; ghost planes readable when the slab is interior
    const double *b;
    const double *pm1;
    double *out;
    double ca, cb, cg;
    double *dot_out;      // device scalar(s) for ST_APPLY_DOT / ST_LIN_PM1_DOT2
    HaloPort port;        // fused ghost exchange of `out` / wait for the ghosts of `u` (comm.h); zero = inactive
};

int launch_stencil(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red);
// true when the TMA-staged plane-marching kernel handles this level (big 3-D grids)
bool stencil_fast_eligible(const LevelDesc &L);

int tune_march(const char *key, long v);   // 0 when the key was recognised

// transfer (DMDA Q1, R = P^T)
int launch_restrict(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *rf, double *bc,
                    const HaloPort &port = HaloPort());
int launch_prolong_add(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *xc, double *xf,
                       const HaloPort &port = HaloPort());

// vector kernels (n = local length)
int launch_dot2(cudaStream_t st, long long n, const double *x, const double *y, double *out2, const Reducer &red);
int launch_dotn(cudaStream_t st, long long n, const double *x, const double *y, double *out1, const Reducer &red);
int launch_absmax(cudaStream_t st, long long n, const double *x, double *out1, const Reducer &red);
int launch_wrms(cudaStream_t st, long long n, const double *x, const double *y, double atol, double rtol, double *out1,
                const Reducer &red);
// x += a p ; r -= a w   with a = num[0]/den[0] read on device
int launch_axpy2(cudaStream_t st, long long n, const double *num, const double *den, const double *p,
                 const double *w, double *x, double *r);
// p = z + (num/den) p ; first != 0: p = z
int launch_aypx_dev(cudaStream_t st, long long n, const double *num, const double *den, const double *z, double *p,
                    int first);
// fused CG updates: (x += a_prev p ; p = z + b p) in one pass, r -= a w, and the final x += a p
int launch_xp_update(cudaStream_t st, long long n, const double *an, const double *ad, const double *bn, const double *bd,
                     const double *z, double *p, double *x, int first, const HaloPort &port = HaloPort());
int launch_r_update(cudaStream_t st, long long n, const double *num, const double *den, const double *w, double *r,
                    const HaloPort &port = HaloPort());
int launch_x_flush(cudaStream_t st, long long n, const double *num, const double *den, const double *p, double *x);
int launch_axpy(cudaStream_t st, long long n, double a, const double *x, double *y);
int launch_aypx(cudaStream_t st, long long n, double a, const double *x, double *y);
int launch_set(cudaStream_t st, long long n, double a, double *y);
int launch_scale_copy(cudaStream_t st, long long n, double a, const double *x, double *y);   // y = a x
int launch_axpby_out(cudaStream_t st, long long n, double a, const double *x, double b, const double *y,
                     double *out);                                                              // out = a x + b y

// reduced-space VI helpers: inactive-set mask, pointwise product (op 0) / maximum (op 1)
int launch_vi_mask(cudaStream_t st, long long n, const double *u, const double *lo, const double *F, double *mask);
int launch_pointwise(cudaStream_t st, long long n, int op, const double *x, const double *y, double *out);

// dense coarse solve x = Ainv b   (n x n, row-major Ainv)
int launch_dense_matvec(cudaStream_t st, int n, const double *Ainv, const double *b, double *x);

// fish problem
int launch_fish_sample(cudaStream_t st, const LevelDesc &L, int dim, int problem, double c0, double c1, double c2,
                       double *f, double *gb);
int launch_initial_state(cudaStream_t st, const LevelDesc &L, const double *gb, int gonboundary, double *u);
int launch_poisson_function(cudaStream_t st, const LevelDesc &L, int dim, double c0, const double *u, const double *f,
                            const double *gb, double *F);

// minimal.c / pattern.c callbacks and SELL SpMV (mp_kernels.cu)
int launch_minimal_sample(cudaStream_t st, int mx, int my, int zs, int zm, int problem, double tent_H, double c, double *g);
int launch_minimal_function(cudaStream_t st, int mx, int my, int zs, int zm, double q, const double *u, const double *g,
                            double *FF);
// ys, ym: rows [ys, ys + ym) only (a y-slab; ym < 0 = the whole grid).  ywrap = 0 in the stencil / transfer launchers:
// the my rows are a slab with ghost rows -1 and my in memory (multi-GPU y-slabs of the periodic grid)
int launch_pattern_init(cudaStream_t st, int mx, int my, double L, double *Y, const double *noise = nullptr,
                        double level = 0.0, int ys = 0, int ym = -1);
int launch_pattern_rhs(cudaStream_t st, int n, double phi, double kappa, const double *Y, double *G);
int launch_heat_rhs(cudaStream_t st, int mx, int my, double D0, const double *u, double *G);
int launch_heat_jac_apply(cudaStream_t st, int mx, int my, double D0, double shift, const double *X, double *out);
int launch_pattern_ifunction(cudaStream_t st, int mx, int my, double Cu, double Cv, int use_shift, double shift,
                             const double *Y, const double *Ydot, double *F, int ywrap = 1);
struct Sell;
int sell_build(cudaStream_t st, int nrows, const int *rowptr, const int *colind, const double *vals, Sell **out);
int sell_spmv(cudaStream_t st, const Sell *A, const double *x, double *y);
void sell_free(Sell *A);
void sell_info(const Sell *A, int *nrows, long long *nnz, long long *padded);

// assembled 2-D 9-point Jacobians (assembled.cu): stencil9 layout vals[s*N + n], s = 3(dj+1) + (di+1)
// differencing step h = fd_step_wp(||u||_2), the same for every column ([PETSc] MatFDColoring "wp")
double fd_step_wp(double unorm);
int fd_jacobian_minimal(cudaStream_t st, int mx, int my, double q, double unorm, const double *u, const double *g,
                        const double *F0, double *vals, double *up, double *Fp);
int launch_fd_perturb(cudaStream_t st, int mx, int my, int ci, int cj, double h, const double *u, double *up);
int launch_fd_extract(cudaStream_t st, int mx, int my, int ci, int cj, double h, const double *F0, const double *Fp,
                      double *vals);
int launch_band_inverse(cudaStream_t st, int n, int bw, const double *B, double *Ainv);   // dense A^-1 from band LU factors
int launch_stencil9_apply(cudaStream_t st, int mx, int my, const double *vals, const double *x, double *y);
int launch_pattern_jac(cudaStream_t st, int mode, int mx, int my, double Cu, double Cv, double shift, double phi,
                       double kappa, const double *Y, const double *X, const double *b, const double *pm1, double ca,
                       double cb, double cg, int jacobi, double *out, int ywrap = 1);
int launch_pattern_transfer(cudaStream_t st, int mode, int Mx, int My, const double *src, double *dst, int ywrap = 1);
int launch_stencil9_rowratio(cudaStream_t st, int mx, int my, const double *vals, double *out);
int launch_poisson_stencil9(cudaStream_t st, int mx, int my, double Lx, double Ly, double cx, double cy, double *vals);
int launch_inject2d(cudaStream_t st, int cmx, int cmy, int fmx, const double *uf, double *uc);
int launch_stencil9_lin(cudaStream_t st, int mx, int my, const double *vals, const double *u, const double *b,
                        const double *pm1, double ca, double cb, double cg, int jacobi, double *out);

}  // namespace p4b
