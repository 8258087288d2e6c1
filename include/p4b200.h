/* p4b200.h -- thin C ABI over the sm_100a CUDA kernels of the fish/DMDA hot path.
 *
 * Plain pointers and sizes only (no torch, no C++ types).  Device pointers are raw
 * CUDA device addresses; `stream` arguments are a cudaStream_t cast to void*.
 * Every function returns 0 on success and a non-zero PetscErrorCode-style int on
 * failure (p4b_last_error() gives the text), mirroring the reference's
 * `PetscCall(...)` convention (c/ch6/fish.c:145 ff).
 *
 * What each entry point replaces in the reference run (paths under /root/reference;
 * [PETSc] = the un-vendored PETSc library reached from that call site):
 *
 *   p4b_poisson_function     c/ch6/poissonfunctions.c:4-115   Poisson{1,2,3}DFunctionLocal
 *   p4b_initial_state        c/ch6/poissonfunctions.c:260-346 InitialState
 *   p4b_fish_sample          c/ch6/fish.c:15-82               u_exact / f_rhs tables
 *   p4b_stencil_apply        [PETSc] MatMult on the matrix of poissonfunctions.c:117-258
 *   p4b_stencil_residual     [PETSc] MatResidual (PCMG, c/ch6/fish.c:239 -> SNESSolve)
 *   p4b_cheb_jacobi          [PETSc] KSPSolve_Chebyshev + PCApply_Jacobi (-mg_levels_*)
 *   p4b_restrict             [PETSc] MatRestrict with the DMDA Q1 interpolation (R = P^T)
 *   p4b_prolong_add          [PETSc] MatInterpolateAdd
 *   p4b_vec_*                [PETSc] VecAXPY/VecAYPX/VecDot/VecNorm (fish.c:254-257, KSPCG)
 *   p4b_mg_create/apply      [PETSc] PCSetUp_MG / PCApply_MG  (-pc_type mg)
 *   p4b_cg_solve             [PETSc] KSPSolve_CG              (fish.c:233 KSPSetType(ksp,KSPCG))
 *   p4b_fish_solve_host      [PETSc] SNESSolve_KSPONLY        (fish.c:231,239)
 *   p4b_minimal_* / p4b_stencil9_* / p4b_inject2d / p4b_dense_matvec
 *                            c/ch7/minimal.c:210-282 + [PETSc] -snes_fd_color, MatMult, Chebyshev/Jacobi, PCLU on the
 *                            assembled Jacobians of the Newton-Krylov-MG run of c/ch8/cluster.sh:70
 *   p4b_pattern_*            c/ch5/pattern.c:146-318 callbacks; the stage Jacobian of the implicit TS run
 *                            (c/ch5/makefile:52-53), matrix-free, and the periodic Q1 transfer
 *
 * There is no CPU fallback behind any of these: without a CUDA device they fail.
 */
#ifndef P4B200_H_
#define P4B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P4B_VERSION 100
#define P4B_MAX_LEVELS 16
#define P4B_MAX_HIST 256

typedef struct p4b_ctx p4b_ctx;   /* device + stream (+ NCCL communicator when nranks > 1) */
typedef struct p4b_mg p4b_mg;     /* level hierarchy, smoother data, CG workspace */

/* The DMDA of fish.c:199-220 plus the PoissonCtx of poissonfunctions.h:43-54. */
typedef struct {
    int dim;             /* 1, 2 or 3 */
    int mx, my, mz;      /* global node counts; unused dimensions are 1 */
    double Lx, Ly, Lz;   /* domain (0,Lx) x (0,Ly) x (0,Lz) */
    double cx, cy, cz;   /* coefficients in -cx u_xx - cy u_yy - cz u_zz = f */
} p4b_grid;

enum { P4B_CYCLE_V = 1, P4B_CYCLE_W = 2 };
enum { P4B_SMOOTH_CHEBYSHEV = 0, P4B_SMOOTH_RICHARDSON = 1 };
enum { P4B_PC_NONE = 0, P4B_PC_JACOBI = 1, P4B_PC_MG = 2 };
enum { P4B_PROBLEM_MANUPOLY = 0, P4B_PROBLEM_MANUEXP = 1, P4B_PROBLEM_ZERO = 2 };
/* [PETSc] KSPConvergedReason values: KSPSolve_CG stops with INDEFINITE_MAT when (p, A p) <= 0, with INDEFINITE_PC when
 * (z, r) <= 0, with DTOL when the residual norm exceeds dtol (1e5) times the initial one */
enum { P4B_CONVERGED_RTOL = 2, P4B_CONVERGED_ATOL = 3, P4B_DIVERGED_ITS = -3, P4B_DIVERGED_DTOL = -4,
       P4B_DIVERGED_INDEFINITE_PC = -8, P4B_DIVERGED_NAN = -9, P4B_DIVERGED_INDEFINITE_MAT = -10 };

/* -pc_mg_* and -mg_levels_* options. */
typedef struct {
    int levels;          /* -pc_mg_levels; 0 = coarsen down to the 3^d grid */
    int cycle;           /* P4B_CYCLE_V | P4B_CYCLE_W  (-pc_mg_cycle_type) */
    int smoother;        /* P4B_SMOOTH_*               (-mg_levels_ksp_type); PC is always Jacobi */
    int smooth_its;      /* -mg_levels_ksp_max_it (PETSc default 2) */
    double emin, emax;   /* -mg_levels_ksp_chebyshev_eigenvalues; emax <= 0 => estimate */
    double est_lo, est_hi; /* esteig transform applied to the analytic lambda_max (0.1, 1.1) */
    int fuse;            /* 1 = fused kernels (default); 0 = one kernel per PETSc operation */
    int use_graph;       /* 1 (default) = replay everything below the finest level as one CUDA graph; needs a
                            non-default stream, an even smooth_its and (multi-GPU) the peer-memory transport,
                            otherwise the same kernels are launched one by one */
} p4b_mg_opts;

typedef struct {
    int its;                       /* number of CG alpha updates (-ksp_converged_reason) */
    int reason;                    /* P4B_CONVERGED_* / P4B_DIVERGED_* */
    double rnorm0, rnorm;          /* preconditioned residual norms ||M^-1 r|| (first / last) */
    int nhist;                     /* entries used in hist */
    double hist[P4B_MAX_HIST];     /* ||M^-1 r_i||, what -ksp_monitor prints */
    double solve_ms;               /* device time of the solve (CUDA events on the ctx stream) */
} p4b_ksp_result;

/* kernel classes for the built-in profiler (finest-level launches only) */
enum {
    P4B_K_APPLY_DOT = 0,   /* w = A p, (p,w)                       16 N */
    P4B_K_RESIDUAL,        /* r = b - A x                          24 N */
    P4B_K_CHEB_ZERO,       /* x = q(A) b   zero-guess Chebyshev(2) 16 N */
    P4B_K_CHEB_FIRST,      /* p1 = x + s D^-1 (b - A x)            24 N */
    P4B_K_CHEB_NEXT,       /* p+ = (1-w) p- + w p + w s D^-1 (b-Ap) 32 N */
    P4B_K_RESTRICT,        /* b_c = P^T r                          8 N + 8 N_c */
    P4B_K_PROLONG,         /* x += P x_c                           16 N + 8 N_c */
    P4B_K_AXPY2,           /* x += a p ; r -= a w                  48 N */
    P4B_K_DOT2,            /* (z,z), (z,r)                         16 N */
    P4B_K_AYPX,            /* p = z + b p                          24 N */
    P4B_K_RESID_RESTRICT,  /* reserved (no fused residual+restriction kernel exists: DESIGN.md section 4) */
    P4B_K_XP_UPDATE,       /* x += a p ; p = z + b p (one pass)    40 N */
    P4B_K_R_UPDATE,        /* r -= a w                             24 N */
    /* not kernels of the roofline table: exchange / latency items, timed the same way (0 algorithmic bytes) */
    P4B_K_HALO,            /* ghost-plane exchange (DMGlobalToLocal) */
    P4B_K_GATHER,          /* completing a replicated level's right-hand side on every rank */
    P4B_K_ALLREDUCE,       /* Krylov scalars over all ranks */
    P4B_K_COARSE,          /* dense coarsest-level solve */
    P4B_K_SUBCYCLE,        /* everything below the finest level as one unit (the CUDA-graph replay) */
    P4B_K_NCLASSES
};

typedef struct {
    long long launches;    /* launches of this class on the finest level since the last reset */
    double ms;             /* summed CUDA-event time of those launches */
    double bytes;          /* summed algorithmic bytes (the table above x points of the level) */
} p4b_kernel_stat;

/* ---- library ---- */
int p4b_version(void);
const char *p4b_last_error(void);
int p4b_device_count(int *n);
/* kernel-selection knobs for tests and A/B measurements (never changes results beyond rounding):
 * "march_enabled" 0|1, "march_min_plane" nodes, "march_P", "march_NT", "march_NS";
 * "rep_points": multigrid levels with at most this many nodes are replicated on every rank (default 70^3);
 * "comm_peer" 0|1 (set before p4b_comm_init): 1 = ghost planes / allreduce / gather as peer-memory kernels
 * over CUDA IPC + NVLink (default), 0 = NCCL send/recv/allreduce/broadcast;
 * "fused_halo" 0|1: 1 (default) = on the peer path a ghost exchange costs no kernel of its own: the kernel that
 * writes a vector also stores its boundary planes into the neighbours' ghost planes and the kernel that reads them
 * waits for the neighbours' flags; 0 = one push kernel per exchange;
 * measurement-only keys: "force_mg" 1|2 (run the multi-GPU kernel variants on ONE GPU: 1 = no neighbours, 2 = scratch
 * memory on this device stands in for both neighbours) and "port_opts" (bit 0: natural chunk order and march
 * direction, bit 1: signal at the end of the boundary CTAs' work) -- the A/B switches behind DESIGN.md section 5;
 * "gmres_cgs" 0|1 (the GMRES of p4b_minimal_solve / p4b_snes2d_solve / p4b_pattern_solve / p4b_ts2d_solve): 0 (default) =
 * modified Gram-Schmidt, one dot and one read-back per basis vector; 1 = classical Gram-Schmidt ([PETSc]'s default
 * orthogonalisation), the dots of a step through p4b_vec_mdot with one read-back -- same iterates up to rounding;
 * "recognise_residual" 0|1 (p4b_snes2d_solve): 1 (default) = probe the caller's residual callback and keep the residual on
 * the device when it is the model the library has as a kernel, 0 = evaluate the callback on the host every time */
int p4b_tune(const char *key, long value);

/* ---- context ---- */
/* stream: the cudaStream_t every kernel, copy and NCCL call of this context is issued on
 * (NULL = the default stream), so the library's work is ordered with the caller's own. */
int p4b_ctx_create(int device, void *stream, p4b_ctx **ctx);
/* the same on a stream the context creates and owns (for C hosts without CUDA headers; one per GPU and host thread) */
int p4b_ctx_create_own_stream(int device, p4b_ctx **ctx);
int p4b_ctx_destroy(p4b_ctx *ctx);
int p4b_ctx_sync(p4b_ctx *ctx);
/* multi-GPU: rank 0 makes a 128-byte id, the host side ships it to the other ranks (torch.distributed
 * in this repo), every rank calls p4b_comm_init.  Halo planes and dot products then go over NCCL. */
int p4b_comm_unique_id(void *id128);
int p4b_comm_init(p4b_ctx *ctx, const void *id128, int rank, int nranks);
/* exchange statistics of the fused ghost exchange since the last reset (reset != 0 clears them after reading):
 * out[0] = boundary-CTA waits, out[1] = their summed spin time (ns), out[2] = the longest single wait (ns),
 * out[3] = system-scope fences that published peer stores, out[4] = their summed time (ns) */
int p4b_comm_stats(p4b_ctx *ctx, unsigned long long out[5], int reset);
/* DMDA-style ownership of the slowest dimension: the first (m % nranks) ranks own one more plane. */
int p4b_slab_range(int m, int nranks, int rank, int *start, int *count);

/* ---- device memory helpers for C hosts (the PETSc-shaped shim); Python uses torch tensors ---- */
int p4b_malloc(p4b_ctx *ctx, size_t bytes, void **dptr);
int p4b_free(p4b_ctx *ctx, void *dptr);
int p4b_memcpy_h2d(p4b_ctx *ctx, void *dst, const void *src, size_t bytes);
int p4b_memcpy_d2h(p4b_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- single-slab building blocks (whole grid on this device; arrays are mx*my*mz doubles) ---- */
int p4b_stencil_apply(p4b_ctx *ctx, const p4b_grid *g, const double *u, double *y);
int p4b_stencil_residual(p4b_ctx *ctx, const p4b_grid *g, const double *b, const double *u, double *r);
/* its Chebyshev/Jacobi steps on A x = b; x is updated in place; work holds one vector */
int p4b_cheb_jacobi(p4b_ctx *ctx, const p4b_grid *g, double emin, double emax, int its, int zero_guess,
                    const double *b, double *x, double *work);
/* fine grid g; coarse arrays have ((mx-1)/2+1) ... nodes per used dimension */
int p4b_restrict(p4b_ctx *ctx, const p4b_grid *gfine, const double *rfine, double *bcoarse);
int p4b_prolong_add(p4b_ctx *ctx, const p4b_grid *gfine, const double *xcoarse, double *xfine);
/* b_c = P^T (b - A x): a COMPOSITION of p4b_stencil_residual and p4b_restrict over a scratch vector (32 N + 8 N_c bytes),
 * kept as a convenience for callers; there is no fused kernel (DESIGN.md section 4 says why) and the multigrid cycle
 * does not use this entry point */
int p4b_residual_restrict(p4b_ctx *ctx, const p4b_grid *gfine, const double *b, const double *x, double *bcoarse);
int p4b_lambda_max_jacobi(const p4b_grid *g, double *lam);

/* ---- reduced-space variational inequalities: what [PETSc] SNESVINEWTONRSLS needs beyond the Poisson kernels for
 * c/ch12/obstacle.c (host logic in p4pdes_b200/obstacle.py) ----
 *   p4b_vi_inactive_mask    mask_i = 0 where the lower bound is active (u_i <= lower_i + 1e-8 and F_i > 0: the rule of
 *                           [PETSc] vi.c that obstacle.c:196-205 repeats), 1 elsewhere
 *   p4b_vec_pointwise_mult  out = x .* y (restriction of vectors and of the matrix-free Jacobian to the inactive set)
 *   p4b_vec_pointwise_max   out = max(x, y) (projection onto u >= psi: SNESVIProjectOntoBounds, the line search's path) */
int p4b_vi_inactive_mask(p4b_ctx *ctx, size_t n, const double *u, const double *lower, const double *F, double *mask);
int p4b_vec_pointwise_mult(p4b_ctx *ctx, size_t n, const double *x, const double *y, double *out);
int p4b_vec_pointwise_max(p4b_ctx *ctx, size_t n, const double *x, const double *y, double *out);

int p4b_vec_dot(p4b_ctx *ctx, size_t n, const double *x, const double *y, double *result_host);
int p4b_vec_norm2(p4b_ctx *ctx, size_t n, const double *x, double *result_host);
/* sum_i ((x_i - y_i) / (atol + rtol max(|x_i|, |y_i|)))^2: [PETSc] TSErrorWeightedNorm2 (TSAdapt's error estimate) */
int p4b_vec_wrms2(p4b_ctx *ctx, size_t n, const double *x, const double *y, double atol, double rtol,
                  double *result_host);
int p4b_vec_norminf(p4b_ctx *ctx, size_t n, const double *x, double *result_host);
/* k dot products (X[i], y), i < k <= 64, with one read-back ([PETSc] VecMDot): X is a HOST array of k device pointers */
int p4b_vec_mdot(p4b_ctx *ctx, size_t n, int k, const double *const *X, const double *y, double *results_host);
int p4b_vec_axpy(p4b_ctx *ctx, size_t n, double a, const double *x, double *y);      /* y += a x */
int p4b_vec_aypx(p4b_ctx *ctx, size_t n, double a, const double *x, double *y);      /* y = x + a y */
int p4b_vec_set(p4b_ctx *ctx, size_t n, double a, double *y);

/* ---- the fish problem on device ---- */
/* f = f_rhs, gb = g_bdry = u_exact sampled at every node (any of the three may be NULL) */
int p4b_fish_sample(p4b_ctx *ctx, const p4b_grid *g, int problem, double *f, double *gb);
/* gonboundary: bit 0 = g on the boundary nodes (-fsh_initial_gonboundary); bit 1 = keep what u holds elsewhere (the
 * RANDOM branch: the caller filled u with the VecSetRandom stream) instead of zeros */
int p4b_initial_state(p4b_ctx *ctx, const p4b_grid *g, const double *gb, int gonboundary, double *u);
int p4b_poisson_function(p4b_ctx *ctx, const p4b_grid *g, const double *u, const double *f,
                         const double *gb, double *F);

/* ---- PCMG + KSPCG ---- */
int p4b_mg_default_opts(p4b_mg_opts *o);
int p4b_mg_create(p4b_ctx *ctx, const p4b_grid *g, const p4b_mg_opts *o, p4b_mg **mg);
/* The Mat-plugin constructor: same as p4b_mg_create, but the constant-coefficient stencil of every level is
 * given explicitly, coef[4*l + {0,1,2,3}] = (diagonal, |off-diagonal| in x, y, z) for level l, FINEST FIRST,
 * as read from the values the user's FormJacobianLocal (poissonfunctions.c:117-258) inserted on that level. */
int p4b_mg_create_stencil(p4b_ctx *ctx, const p4b_grid *g, const p4b_mg_opts *o, const double *coef,
                          int nlevels_coef, p4b_mg **mg);
int p4b_mg_destroy(p4b_mg *mg);
/* Host-only: the level hierarchy and slab ownership p4b_mg_create would use on `nranks` devices.
 * Level 0 is the coarsest.  m3[3*l..] = node counts, zs/zm[l*nranks + r] = planes of the slowest dimension
 * rank r owns on level l ("coarse plane K belongs to the owner of fine plane 2K"), replicated[l] = 1 when
 * the level is computed redundantly on every rank.  Arrays sized for P4B_MAX_LEVELS levels; any may be NULL. */
int p4b_plan_levels(const p4b_grid *g, const p4b_mg_opts *o, int nranks, int *nlevels, int *m3, int *zs, int *zm,
                    int *replicated);
int p4b_mg_nlevels(p4b_mg *mg, int *nlevels);
/* level 0 = coarsest; m[3], eig[2] = (emin, emax) used by the smoother on that level */
int p4b_mg_level_info(p4b_mg *mg, int level, int *m, double *eig);
/* the slab of the finest grid this rank owns: planes [start, start+count) of the slowest dimension */
int p4b_mg_local_range(p4b_mg *mg, int *start, int *count, size_t *nlocal);
/* y = A x with the finest level's operator ([PETSc] MatMult); x, y: nlocal doubles on device */
int p4b_mg_matmult(p4b_mg *mg, const double *x, double *y);
/* z = M^-1 r (one multigrid cycle from a zero guess); r, z: nlocal doubles on device */
int p4b_mg_apply(p4b_mg *mg, const double *r, double *z);
/* solve A x = b from x = 0 with preconditioned CG; b, x: nlocal doubles on device */
int p4b_cg_solve(p4b_mg *mg, int pc_type, const double *b, double *x, double rtol, double abstol,
                 int max_it, p4b_ksp_result *res);
/* same with HOST buffers of the local slab: H2D(b), solve, D2H(x) */
int p4b_cg_solve_host(p4b_mg *mg, int pc_type, const double *b_host, double *x_host, double rtol,
                      double abstol, int max_it, p4b_ksp_result *res);
/* SNESKSPONLY on host buffers: F0 = F(u), solve J y = F0, u <- u - y.  f, gb, u: local slab on host */
int p4b_fish_solve_host(p4b_mg *mg, const double *f_host, const double *gb_host, double *u_host,
                        double rtol, double abstol, int max_it, p4b_ksp_result *res);

/* fish.c's built-in problems on this rank's slab: b = F(u0), u0 = InitialState(zeros[, g on boundary]),
 * uexact (fish.c:186-187,237-239,248-253).  Outputs are nlocal doubles on device; any may be NULL. */
int p4b_mg_fish_setup(p4b_mg *mg, int problem, int gonboundary, double *b, double *u0, double *uexact);

/* ---- callbacks of the other two DMDA drivers on the BASELINE path, as device kernels ----
 *   p4b_minimal_function        c/ch7/minimal.c:210-282  FormFunctionLocal   (u, g, FF: my*mx doubles, i fastest)
 *   p4b_minimal_sample          c/ch7/minimal.c:27-42    g_bdry_tent (problem 0) / g_bdry_catenoid (problem 1)
 *   p4b_pattern_initial_state   c/ch5/pattern.c:146-179  InitialState without noise (Y: my*mx*2, (u,v) interleaved)
 *   p4b_pattern_rhsfunction     c/ch5/pattern.c:185-199  FormRHSFunctionLocal
 *   p4b_pattern_ifunction       c/ch5/pattern.c:242-267  FormIFunctionLocal   F = Ydot - C L9(Y), periodic
 *   p4b_pattern_ijacobian_mult  c/ch5/pattern.c:274-318  action of the matrix FormIJacobianLocal assembles
 *   p4b_sell_*                  [PETSc] MatMult_SeqAIJ for assembled Jacobians: CSR (host) -> SELL-32 (device) */
int p4b_minimal_sample(p4b_ctx *ctx, int mx, int my, int problem, double tent_H, double catenoid_c, double *g);
int p4b_minimal_function(p4b_ctx *ctx, int mx, int my, double q, const double *u, const double *g, double *FF);
int p4b_pattern_initial_state(p4b_ctx *ctx, int mx, int my, double L, double *Y);
/* the same with noise (c/ch5/pattern.c:159-165, -ptn_noisy_init): noise (device, 2*mx*my doubles in [0,1), the
 * VecSetRandom stream in natural (u,v)-interleaved order) is scaled by level and the patch added on top of it */
int p4b_pattern_initial_state_noisy(p4b_ctx *ctx, int mx, int my, double L, const double *noise, double level, double *Y);

/* ---- [PETSc] PetscRandom, default type rander48 (VecSetRandom at c/ch5/pattern.c:163 and
 * c/ch6/poissonfunctions.c:268-270).  Host-only.  PETSc's generator is the 48-bit linear congruential generator of
 * drand48 -- X <- (0x5DEECE66D X + 0xB) mod 2^48, X0 = (seed << 16) | 0x330E, value X / 2^48 -- seeded with
 * 0x12345678 + 76543 * rank.  Restated from PETSc's rander48.c as remembered; PETSc is not installable here and no
 * golden of the reference uses a random vector, so equality with PETSc's stream is UNPINNED (tests pin it on glibc's
 * srand48/drand48, the same recurrence). */
unsigned long long p4b_rander48_seed(unsigned long seed);
/* n values in [0,1) into out (host); *state advances */
int p4b_rander48_fill(unsigned long long *state, size_t n, double *out);
int p4b_pattern_rhsfunction(p4b_ctx *ctx, int mx, int my, double phi, double kappa, const double *Y, double *G);
int p4b_pattern_ifunction(p4b_ctx *ctx, int mx, int my, double L, double Du, double Dv, const double *Y,
                          const double *Ydot, double *F);
int p4b_pattern_ijacobian_mult(p4b_ctx *ctx, int mx, int my, double L, double Du, double Dv, double shift,
                               const double *X, double *JX);
/* ---- assembled Jacobians of the 2-D drivers: finite-difference assembly and the solver kernels on them ----
 *   p4b_minimal_jacobian_fd  [PETSc] SNESComputeJacobianDefaultColor on c/ch7/minimal.c:210-282 (-snes_fd_color):
 *                            9 colours (DMDA BOX stencil), MatFDColoring's default "wp" differencing (one step
 *                            h = sqrt(eps) sqrt(1 + ||u||_2) for every column); F0 = F(u) already computed
 *   p4b_poisson_stencil9     the matrix c/ch6/poissonfunctions.c:152-193 (Poisson2DJacobianLocal) inserts on an
 *                            Lx x Ly rectangle, in the same layout: what c/ch7/minimal.c:142-145 registers as its
 *                            (approximate) Jacobian -- Newton's matrix without -snes_fd_color / -snes_mf_operator,
 *                            the preconditioner's under -snes_mf_operator
 *   p4b_stencil9_apply       [PETSc] MatMult on that matrix
 *   p4b_stencil9_lin         [PETSc] KSPSolve_Chebyshev/Richardson step + PCApply_Jacobi on it:
 *                            out = ca*pm1 + cb*u + cg*B(b - A u), B = diag(A)^-1 (jacobi != 0) or I
 *                            (pm1 may be NULL or alias out; b may be NULL)
 *   p4b_dense_matvec         [PETSc] PCApply_LU on the coarsest level (x = Ainv b, Ainv n x n row-major on device)
 * Matrix layout "stencil9": 9*mx*my doubles, vals[s*mx*my + j*mx + i] = dF(i,j)/du(i+di, j+dj), s = 3(dj+1) + (di+1). */
int p4b_minimal_jacobian_fd(p4b_ctx *ctx, int mx, int my, double q, const double *u, const double *g, const double *F0,
                            double *vals9);
int p4b_poisson_stencil9(p4b_ctx *ctx, int mx, int my, double Lx, double Ly, double cx, double cy, double *vals9);
int p4b_stencil9_apply(p4b_ctx *ctx, int mx, int my, const double *vals9, const double *x, double *y);
int p4b_stencil9_lin(p4b_ctx *ctx, int mx, int my, const double *vals9, const double *u, const double *b,
                     const double *pm1, double ca, double cb, double cg, int jacobi, double *out);
int p4b_dense_matvec(p4b_ctx *ctx, int n, const double *Ainv, const double *b, double *x);
/* max_n sum_s |a_ns| / |a_nn| (Gershgorin bound of lambda_max(D^-1 A), the Chebyshev target); work: mx*my doubles */
int p4b_stencil9_gershgorin(p4b_ctx *ctx, int mx, int my, const double *vals9, double *work, double *result_host);
/* [PETSc] DMCreateInjection (DMDA, ratio 2): uc(I,J) = uf(2I,2J); fine grid (2cmx-1) x (2cmy-1) */
int p4b_inject2d(p4b_ctx *ctx, int cmx, int cmy, const double *ufine, double *ucoarse);
/* ---- pattern.c implicit stage equation F(t,Y,(Y-Y0)/dt) = G(t,Y): its Jacobian, matrix-free, on a periodic level ----
 *   J X = shift*X - C L9(X) - G'(Y) X : FormIJacobianLocal (c/ch5/pattern.c:274-318) minus FormRHSJacobianLocal
 *   (:202-236) at the iterate Y ([PETSc] TSComputeIJacobian); Y == NULL drops the RHS block (-ptn_no_rhsjacobian).
 *   p4b_pattern_jac_apply     out = J X                                        [PETSc] MatMult
 *   p4b_pattern_jac_lin       out = ca*pm1 + cb*X + cg*B(b - J X), B = diag(J)^-1 (jacobi) or I   smoother step / residual
 *   p4b_pattern_jac_gershgorin max over rows of sum_j |J_nj| / |J_nn|          Chebyshev target
 *   p4b_pattern_restrict / _prolong_add / _inject   [PETSc] DMCreateInterpolation/Injection on the periodic 2-dof DMDA
 *                             (fine grid 2Mx x 2My, coarse Mx x My; R = P^T)
 * All vectors: my*mx*2 doubles, (u,v) interleaved, i fastest. */
int p4b_pattern_jac_apply(p4b_ctx *ctx, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                          double shift, const double *Y, const double *X, double *out);
int p4b_pattern_jac_lin(p4b_ctx *ctx, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                        double shift, const double *Y, const double *X, const double *b, const double *pm1, double ca,
                        double cb, double cg, int jacobi, double *out);
int p4b_pattern_jac_gershgorin(p4b_ctx *ctx, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                               double shift, const double *Y, double *work, double *result_host);
int p4b_pattern_restrict(p4b_ctx *ctx, int Mx, int My, const double *rfine, double *bcoarse);
int p4b_pattern_prolong_add(p4b_ctx *ctx, int Mx, int My, const double *xcoarse, double *xfine);
int p4b_pattern_inject(p4b_ctx *ctx, int Mx, int My, const double *yfine, double *ycoarse);
/* out = a x + b y (x, y may alias out) and a device-to-device copy: [PETSc] VecAXPBY / VecWAXPY / VecCopy */
int p4b_vec_axpby(p4b_ctx *ctx, size_t n, double a, const double *x, double b, const double *y, double *out);
int p4b_vec_copy(p4b_ctx *ctx, size_t n, const double *x, double *y);
/* ---- the whole minimal.c run in one call: [PETSc] SNESSolve for `./minimal -snes_fd_color -pc_type mg|none
 * [-snes_grid_sequence k]` (c/ch7/minimal.c:128-181, c/ch8/cluster.sh:70) -- Newton + cubic bt line search, GMRES(30) or CG,
 * V cycle on FD-coloured level Jacobians at the injected iterate, grid sequencing; host logic csrc/nk_solver.hpp, every
 * vector operation one of the kernels above.  Field-for-field the options of minimal.c:69-103 and of PETSc. ---- */
typedef struct {
    int problem;                 /* 0 tent, 1 catenoid            -ms_problem */
    double q, catenoid_c, tent_H;/*                               -ms_q -ms_catenoid_c -ms_tent_H */
    int exact_init;              /*                               -ms_exact_init */
    int grid_x, grid_y, refine, grid_sequence;   /*               -da_grid_x -da_grid_y -da_refine -snes_grid_sequence */
    int ksp_type;                /* 0 gmres, 1 cg                 -ksp_type */
    double ksp_rtol;
    int ksp_max_it, gmres_restart;
    int pc_type;                 /* 0 none, 1 mg                  -pc_type */
    int mg_levels, smooth_its;   /*                               -pc_mg_levels -mg_levels_ksp_max_it */
    double snes_rtol, snes_stol, snes_atol;
    int snes_max_it;
    int snes_monitor;            /* 0 off, 1 -snes_monitor, 2 -snes_monitor_short */
    int snes_converged_reason, ksp_converged_reason;
    int mf_operator;             /* -snes_mf_operator: J v by differencing the residual ([PETSc] MatMFFD "wp"); the
                                    assembled matrix then only builds the preconditioner */
    int jacobian;                /* the assembled matrix: 0 = FD-coloured Jacobian of the residual (-snes_fd_color);
                                    1 = what c/ch7/minimal.c:142-145 registers, Poisson2DJacobianLocal ("ONLY APPROXIMATE"):
                                    PETSc's choice when -snes_fd_color is absent, with or without -snes_mf_operator */
} p4b_minimal_opts;
typedef struct {
    int mx, my, its, reason, nksp;   /* reason: 2 FNORM_ABS, 3 FNORM_RELATIVE, 4 SNORM_RELATIVE, < 0 diverged ([PETSc] numbering) */
    int ksp_its[64];
    double lambda[64];               /* accepted line-search step of every Newton iteration */
    double fnorm[65];                /* ||F|| before the first and after every iteration */
} p4b_minimal_stage;
typedef struct {
    int mx, my, nstages;             /* final grid; grid-sequence stages, coarsest first */
    p4b_minimal_stage stage[16];
    double errinf;                   /* |u - uexact|_inf (catenoid, q = -1/2), else -1 */
    int error;
    char errmsg[256];
} p4b_minimal_result;
typedef void (*p4b_line_fn)(const char *line, void *ctx);    /* receives the lines minimal.c / PETSc would print */
int p4b_minimal_default_opts(p4b_minimal_opts *o);
/* u_out (device, may be NULL): the final iterate, u_capacity doubles available */
int p4b_minimal_solve(p4b_ctx *ctx, const p4b_minimal_opts *opts, p4b_line_fn line, void *line_ctx, double *u_out,
                      size_t u_capacity, p4b_minimal_result *result);
/* ---- the same Newton-Krylov-multigrid solve for ANY residual on a 2-D DMDA with the BOX stencil, supplied by the caller as
 * a host callback: the FormFunctionLocal contract (DMDASNESSetFunctionLocal, c/ch7/minimal.c:140-141).  The
 * callback is invoked on the host with the whole mx x my grid (one logical rank: xs = 0, xm = mx, ...), arrays in DMDA
 * natural ordering (a binding wraps DMDALocalInfo and a[j][i] pointer tables around them), on every grid of the
 * hierarchy and of the grid sequence; Jacobians by coloured finite differences of it; the algebra stays on the device.
 * opts: the solver fields of the options struct above -- grid_x/grid_y/refine = the grid u0 lives on, grid_sequence, ksp_*, pc_*,
 * mg_*, snes_* -- the -ms_* fields are ignored.  u0_host: initial iterate; u_out_host: the solution on the final grid. ---- */
typedef int (*p4b_residual2d_fn)(void *user, int mx, int my, const double *u_host, double *F_host);
int p4b_snes2d_solve(p4b_ctx *ctx, const p4b_minimal_opts *opts, p4b_residual2d_fn residual, void *user,
                     const double *u0_host, p4b_line_fn line, void *line_ctx, double *u_out_host, size_t u_capacity,
                     p4b_minimal_result *result);
/* Recognition: before the solve the residual callback is probed on every grid the solve will touch (F(0) gives the
 * Dirichlet data, a generic iterate the exponent q, a second one the check; 2 evaluations per grid + 1).  If it IS
 * c/ch7/minimal.c:210-282 -- the unchanged minimal.c under the shim -- to rounding, the solve keeps the residual on the
 * device (minimal_function_kernel) and calls the callback no more; otherwise every evaluation is a host callback (nine per
 * level Jacobian).  p4b_snes2d_last_route(): 1 = device residual after recognition, 0 = host callbacks. */
int p4b_snes2d_last_route(void);
/* the same with a caller's monitor ([PETSc] SNESMonitorSet, c/ch7/minimal.c:146-148 registers MSEMonitor :286-345): called
 * on the host before the first and after every Newton iteration of every grid-sequence stage, ahead of the -snes_monitor
 * line, with the current iterate on that stage's grid; tablevel = the stages still to come ([PETSc] PetscObjectGetTabLevel
 * of the SNES under -snes_grid_sequence).  A non-zero return aborts the solve (error 66).  monitor may be NULL. */
typedef int (*p4b_monitor2d_fn)(void *user, int mx, int my, int its, double fnorm, int tablevel, const double *u_host);
int p4b_snes2d_solve_monitored(p4b_ctx *ctx, const p4b_minimal_opts *opts, p4b_residual2d_fn residual,
                               p4b_monitor2d_fn monitor, void *user, const double *u0_host, p4b_line_fn line, void *line_ctx,
                               double *u_out_host, size_t u_capacity, p4b_minimal_result *result);
/* ---- the whole pattern.c run in one call: [PETSc] TSSolve for `./pattern [-ts_type arkimex|beuler|cn] -pc_type mg|none`
 * (c/ch5/pattern.c:99-125, c/ch5/makefile:49-62): TSARKIMEX3 + TSAdaptBasic + MATCHSTEP, TSTHETA, or TSBDF(2), Newton + bt;
 * stage solves GMRES(30) + V cycle on the matrix-free stage operator; host logic csrc/ts_solver.hpp. ---- */
typedef struct {
    double L, Du, Dv, phi, kappa;          /* -ptn_L -ptn_Du -ptn_Dv -ptn_phi -ptn_kappa (pattern.c:47-52) */
    int no_rhsjacobian, call_back_report;  /* -ptn_no_rhsjacobian -ptn_call_back_report */
    int grid_x, grid_y, refine;            /* -da_grid_x -da_grid_y -da_refine (periodic: refine doubles) */
    int ts_type;                           /* 0 arkimex (pattern.c's default), 1 beuler, 2 cn, 3 bdf (order 2) */
    double ts_dt, ts_max_time;
    int ts_max_steps;
    double ts_rtol, ts_atol;
    int ts_monitor;
    int pc_type;                           /* 0 none, 1 mg */
    int smooth_its;
    double mg_rscale;                      /* -p4b_mg_rscale: 1 = [PETSc] R = P^T, 0.25 = averaging restriction */
    double snes_rtol, snes_stol, snes_atol;
    int snes_max_it;
    double ksp_rtol;
    int ksp_max_it, gmres_restart;
    int snes_converged_reason, ksp_converged_reason;
} p4b_pattern_opts;
typedef struct {
    int m, nsteps, rejected;               /* grid m x m x 2; accepted steps; rejected step attempts (arkimex) */
    long long ksp_its_total, newton_its_total;
    double t_final, dt_last;
    double step_t[512], step_dt[512];      /* the first 512 steps: time after the step, step taken */
    int step_newton[512];                  /* Newton iterations of the step (summed over the stages for arkimex) */
    int error;
} p4b_pattern_result;
int p4b_pattern_default_opts(p4b_pattern_opts *o);
/* Y_out (device, may be NULL): the final state, 2*m*m doubles, (u,v) interleaved */
int p4b_pattern_solve(p4b_ctx *ctx, const p4b_pattern_opts *opts, p4b_line_fn line, void *line_ctx, double *Y_out,
                      size_t Y_capacity, p4b_pattern_result *result);
/* the same from the CALLER's initial state Y0 (device, 2*m*m doubles; Y_out may alias it): what TSSolve(ts, x) of the
 * PETSc-shaped shim binds (c/ch5/pattern.c:123-125).  Only the solver's own lines are reported (no banner, no call-back
 * report: the caller prints those, pattern.c:94-96,127-135).  Y0 = NULL is p4b_pattern_solve. */
int p4b_pattern_solve_from(p4b_ctx *ctx, const p4b_pattern_opts *opts, const double *Y0, p4b_line_fn line, void *line_ctx,
                           double *Y_out, size_t Y_capacity, p4b_pattern_result *result);
/* ---- c/ch7/solns/bratu2D.c: - lap u - lambda e^u = 0 by FAS multigrid + nonlinear Gauss-Seidel ----
 *   p4b_bratu_function   FormFunctionLocal, bratu2D.c:196-225 (F = residual - b when b != NULL)
 *   p4b_bratu_ngs        NonlinearGS, :229-299: `sweeps` RED-BLACK sweeps of pointwise Newton (the reference's sweep is
 *                        lexicographic, i.e. sequential; same fixed point)
 *   p4b_bratu_exact      g_liouville at every node (exact = 1; :40-45) or zeros
 *   p4b_bratu_solve      ./bratu2D -snes_type fas [-snes_fas_type full] -fas_levels_snes_type ngs -fas_coarse_snes_type ngs:
 *                        VecSet(u,0), the FAS cycle ([PETSc] SNESFAS restated with the golden's components; which parts are
 *                        pinned by c/ch7/solns/output/bratu2D.test1 is stated in oracle/bratu_oracle.py), error norm */
typedef struct {
    double lambda;                 /* -lb_lambda */
    int exact;                     /* -lb_exact */
    int grid_x, grid_y, refine;    /* -da_grid_x/_y (3), -da_refine */
    int levels;                    /* -snes_fas_levels (0 = refine + 1: down to the -da_grid base) */
    double snes_rtol;
    int snes_max_it;
    int smooth_sweeps, smooth_its; /* -fas_levels_snes_ngs_sweeps, -fas_levels_snes_max_it */
    int coarse_sweeps, coarse_its; /* -fas_coarse_snes_ngs_sweeps, -fas_coarse_snes_max_it */
    int full_cycle;                /* -snes_fas_type full (1) | multiplicative (0) */
    int monitor, converged_reason; /* -snes_monitor_short, -snes_converged_reason */
} p4b_bratu_opts;
typedef struct {
    int mx, my, its, reason, nnorm;
    double fnorm[64];
    double errinf;                 /* |u - uexact|_inf with -lb_exact, else -1 */
    long long residual_calls, ngs_calls;
    double solve_ms;               /* CUDA events around the solve */
} p4b_bratu_result;
int p4b_bratu_default_opts(p4b_bratu_opts *o);
int p4b_bratu_function(p4b_ctx *ctx, int mx, int my, double lambda, int exact, const double *u, const double *b, double *F);
int p4b_bratu_ngs(p4b_ctx *ctx, int mx, int my, double lambda, int exact, int sweeps, const double *b, double *u);
int p4b_bratu_exact(p4b_ctx *ctx, int mx, int my, int exact, double *g);
int p4b_bratu_solve(p4b_ctx *ctx, const p4b_bratu_opts *opts, p4b_line_fn line, void *line_ctx, double *u_out,
                    size_t u_capacity, p4b_bratu_result *result);

/* [PETSc] TSMonitorSet with a solution viewer (-ts_monitor binary:t.dat -ts_monitor_solution binary:u.dat, c/ch5/MOVIES.md:44):
 * a process-wide step monitor of p4b_pattern_solve / p4b_ts2d_solve.  It is called at step 0 and after every accepted
 * step with the time and a HOST copy of the state (this rank's rows on slabs); a non-zero return aborts the solve.
 * fn = NULL removes it. */
typedef int (*p4b_ts_step_fn)(void *user, int step, double t, const double *Y_host, size_t n);
int p4b_set_ts_step_monitor(p4b_ts_step_fn fn, void *user);

/* Multi-GPU (BASELINE config 5: 2048^2 on 8 GPUs): with a context that carries a communicator (p4b_comm_init)
 * p4b_pattern_solve / p4b_pattern_solve_from run on y-slabs of the periodic DMDA (c/ch5/pattern.c:79-84) -- ring
 * exchange of one ghost row per side ([PETSc] DMGlobalToLocal), all-reduced dot products, small levels replicated.
 * Y0 / Y_out are then THIS RANK's rows, 2 * m * (m / nranks) doubles; p4b_pattern_slab_plan (host only) says which
 * levels are distributed and which rows a rank owns on each (arrays of >= 32 ints; returns the number of levels,
 * finest first, or a negative error code). */
int p4b_pattern_slab_plan(int m, int grid_x, int mg, int nranks, int rank, int *level_m, int *distributed, int *ys, int *ym);
/* ---- the same time steppers for ANY two-component system on the periodic m x m DMDA given by HOST callbacks and no
 * Jacobian: F(t, Y, Ydot) and G(t, Y) of the DMDATSSet{IFunction,RHSFunction}Local contract (c/ch5/pattern.c:103-114,
 * 185-199, 242-267; arrays whole-grid, (u,v) interleaved, natural ordering).  The stage operator is the differenced residual
 * ([PETSc] MatMFFD "wp"), so there is no multigrid: opts->pc_type must be 0 (none).  F must be M Ydot + f(Y) with a constant
 * M (the caller checks; every method-of-lines system is); the callbacks are called with t = 0 (autonomous systems).  The
 * model fields of opts (L, Du, ...) are ignored.  Y_inout_host: initial state in, final state out. ---- */
typedef int (*p4b_ifunction2d_fn)(void *user, int m, double t, const double *Y_host, const double *Ydot_host, double *F_host);
typedef int (*p4b_rhsfunction2d_fn)(void *user, int m, double t, const double *Y_host, double *G_host);
int p4b_ts2d_solve(p4b_ctx *ctx, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction, p4b_rhsfunction2d_fn rhsfunction,
                   void *user, double *Y_inout_host, size_t Y_capacity, p4b_line_fn line, void *line_ctx,
                   p4b_pattern_result *result);
/* ---- ... and for the method-of-lines system of ANY DMDA driver: the state is a field of n doubles whose layout only the
 * callbacks know (m is passed as 0).  This is how the unchanged c/ch5/heat.c runs under the shim (one component, Neumann in
 * x, periodic in y, RHSFunction only: heat.c:60-75,141-163).  Same steppers, plus opts->ts_type = 4: [PETSc] TSRK "3bs"
 * (Bogacki-Shampine 3(2), explicit, TSAdaptBasic at order 3 -- pinned by c/ch5/output/heat.test2), for which the system must
 * be Ydot = G(t, Y): the IFunction is not called.  pc_type is ignored for rk and must be 0 otherwise.
 * p4b_ts_time_step(): the step the running integrator is about to take (at the last monitor call: proposes next) -- what
 * [PETSc] TSGetTimeStep answers inside a TSMonitorSet monitor (heat.c:133). ---- */
int p4b_ts_solve_callbacks(p4b_ctx *ctx, const p4b_pattern_opts *opts, p4b_ifunction2d_fn ifunction,
                           p4b_rhsfunction2d_fn rhsfunction, void *user, double *Y_inout_host, size_t n, p4b_line_fn line,
                           void *line_ctx, p4b_pattern_result *result);
double p4b_ts_time_step(void);
/* ---- c/ch5/heat.c device-resident.  p4b_heat_rhs = FormRHSFunctionLocal (heat.c:141-163): G = D0 lap_h(u) + f on the unit
 * square, mx nodes with hx = 1/(mx-1) and the Neumann data through mirrored ghost values in x, my nodes with hy = 1/my
 * periodic in y, f_source and gamma_neumann of heat.c:16-23; u, G device arrays of mx*my doubles, n = j mx + i.
 * p4b_heat_jac_apply = (shift I - dG/du) X, the rows of FormRHSJacobianLocal (heat.c:166-208) applied matrix-free.
 * p4b_heat_solve = TSSolve for that system (Ydot = G(u), heat.c:66-92) with the steppers of p4b_ts_solve_callbacks, G and
 * the stage operator being these kernels; one level, so opts->pc_type must be 0 unless ts_type is rk.  The shim calls it
 * when the registered callback IS this function (probed, and re-checked at the final state). ---- */
int p4b_heat_rhs(p4b_ctx *ctx, int mx, int my, double D0, const double *u, double *G);
int p4b_heat_jac_apply(p4b_ctx *ctx, int mx, int my, double D0, double shift, const double *X, double *JX);
int p4b_heat_solve(p4b_ctx *ctx, const p4b_pattern_opts *opts, int mx, int my, double D0, double *Y_inout_host,
                   p4b_line_fn line, void *line_ctx, p4b_pattern_result *result);
typedef struct p4b_sell p4b_sell;
int p4b_sell_create(p4b_ctx *ctx, int nrows, const int *rowptr_host, const int *colind_host, const double *vals_host,
                    p4b_sell **A);
int p4b_sell_spmv(p4b_sell *A, const double *x, double *y);
int p4b_sell_info(p4b_sell *A, int *nrows, long long *nnz, long long *padded_nnz);
int p4b_sell_destroy(p4b_sell *A);

/* ---- profiler (CUDA events around finest-level launches on the ctx stream) ---- */
/* on = 1: finest-level kernels (+ exchanges and the sub-cycle as one unit); on = 2: trace mode, every launch on
 * every level is bracketed (the CUDA graph is bypassed so that they are visible) */
int p4b_profile_enable(p4b_mg *mg, int on);
/* trace mode: the same statistics per level (0 = coarsest) */
int p4b_profile_get_level(p4b_mg *mg, int level, int kernel_class, p4b_kernel_stat *out);
int p4b_profile_reset(p4b_mg *mg);
int p4b_profile_get(p4b_mg *mg, int kernel_class, p4b_kernel_stat *out);
/* kernels launched by this library in this process so far (bench.py reports the difference) */
long long p4b_launch_count(void);
const char *p4b_kernel_name(int kernel_class);

#ifdef __cplusplus
}
#endif
#endif /* P4B200_H_ */
