// fishfn.cu -- the fish problem on the device: manufactured functions, initial state, F(u).
//
//   fish_sample_kernel       c/ch6/fish.c:15-82 (u_exact_*/f_rhs_* tables, :115-123)
//   initial_state_kernel     c/ch6/poissonfunctions.c:260-346 (InitialState, ZEROS branch)
//   poisson_function_kernel  c/ch6/poissonfunctions.c:4-115  (Poisson{1,2,3}DFunctionLocal)
//
// The reference evaluates f_rhs/g_bdry through host function pointers (poissonfunctions.h:49-51);
// on the device they are node-sampled arrays f and gb (the C shim fills them by calling the user's
// host functions once; p4b_fish_sample fills them for fish.c's three built-in problems).
#include "kernels.h"

namespace p4b {

__device__ __forceinline__ void node_of(const LevelDesc &L, long long n, int &i, int &j, int &k) {
    const int plane = L.nx * L.ny;
    const int kl = (int)(n / plane);
    const int rem = (int)(n - (long long)kl * plane);
    j = rem / L.nx;
    i = rem - j * L.nx;
    k = kl + L.zs;
}

__device__ __forceinline__ bool on_bdry(const LevelDesc &L, int i, int j, int k) {
    return (L.ax && (i == 0 || i == L.nx - 1)) || (L.ay && (j == 0 || j == L.ny - 1)) ||
           (L.az && (k == 0 || k == L.nz - 1));
}

__device__ __forceinline__ double uexact_dev(int dim, int problem, double x, double y, double z) {
    if (problem == P4B_PROBLEM_MANUPOLY) {
        double a = x * x * (1.0 - x * x);
        if (dim >= 2) a = a * y * y * (y * y - 1.0);
        if (dim >= 3) a = a * z * z * (z * z - 1.0);
        return a;
    }
    if (problem == P4B_PROBLEM_MANUEXP) {
        if (dim == 1) return -exp(x);
        if (dim == 2) return -x * exp(y);
        return -x * exp(y + z);
    }
    return 0.0;
}

__device__ __forceinline__ double frhs_dev(int dim, int problem, double x, double y, double z, double cx, double cy,
                                           double cz) {
    if (problem == P4B_PROBLEM_MANUPOLY) {
        if (dim == 1) return cx * 12.0 * x * x - 2.0;
        const double aa = x * x * (1.0 - x * x), bb = y * y * (y * y - 1.0);
        const double ddaa = 2.0 * (1.0 - 6.0 * x * x), ddbb = 2.0 * (6.0 * y * y - 1.0);
        if (dim == 2) return -(cx * ddaa * bb + cy * aa * ddbb);
        const double cc = z * z * (z * z - 1.0), ddcc = 2.0 * (6.0 * z * z - 1.0);
        return -(cx * ddaa * bb * cc + cy * aa * ddbb * cc + cz * aa * bb * ddcc);
    }
    if (problem == P4B_PROBLEM_MANUEXP) {
        if (dim == 1) return exp(x);
        if (dim == 2) return x * exp(y);
        return 2.0 * x * exp(y + z);
    }
    return 0.0;
}

__global__ void __launch_bounds__(256) fish_sample_kernel(const LevelDesc L, int dim, int problem, double cx, double cy,
                                                           double cz, double *f, double *gb) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= L.nlocal()) return;
    int i, j, k;
    node_of(L, n, i, j, k);
    // slot coordinates -> problem coordinates (2-D grids live in slots x and z)
    const double X0 = i * L.hx, X1 = j * L.hy, X2 = k * L.hz;
    const double x = X0, y = (dim == 3) ? X1 : (dim == 2 ? X2 : 0.0), z = (dim == 3) ? X2 : 0.0;
    if (f) f[n] = frhs_dev(dim, problem, x, y, z, cx, cy, cz);
    if (gb) gb[n] = uexact_dev(dim, problem, x, y, z);
}

__global__ void __launch_bounds__(256) initial_state_kernel(const LevelDesc L, const double *__restrict__ gb,
                                                             int gonboundary, double *__restrict__ u) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= L.nlocal()) return;
    int i, j, k;
    node_of(L, n, i, j, k);
    if ((gonboundary & 1) && on_bdry(L, i, j, k)) u[n] = gb[n];
    else if (!(gonboundary & 2)) u[n] = 0.0;
}

// F(u).  u must have readable ghost planes when the slab is interior (k-1, k+1 reads).
__global__ void __launch_bounds__(256) poisson_function_kernel(const LevelDesc L, int dim, double c0,
                                                                const double *__restrict__ u,
                                                                const double *__restrict__ f,
                                                                const double *__restrict__ gb,
                                                                double *__restrict__ F) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= L.nlocal()) return;
    int i, j, k;
    node_of(L, n, i, j, k);
    const int plane = L.nx * L.ny;
    const double uc = u[n];
    if (on_bdry(L, i, j, k)) {
        // poissonfunctions.c:13-14 (1-D: cx*(2/h)), :46-47, :91-92
        const double s = (dim == 1) ? c0 * (2.0 / L.hx) : L.diag;
        F[n] = (uc - gb[n]) * s;
        return;
    }
    // neighbours: g where the neighbour is a boundary node, else u  (:16-19, :49-56, :94-105)
    double uw = 0, ue = 0, us = 0, un = 0, ud = 0, uu = 0;
    if (L.ax) {
        uw = (i - 1 == 0) ? gb[n - 1] : u[n - 1];
        ue = (i + 1 == L.nx - 1) ? gb[n + 1] : u[n + 1];
    }
    if (L.ay) {
        us = (j - 1 == 0) ? gb[n - L.nx] : u[n - L.nx];
        un = (j + 1 == L.ny - 1) ? gb[n + L.nx] : u[n + L.nx];
    }
    if (L.az) {
        ud = (k - 1 == 0) ? gb[n - plane] : u[n - plane];
        uu = (k + 1 == L.nz - 1) ? gb[n + plane] : u[n + plane];
    }
    if (dim == 1) {
        F[n] = c0 * (2.0 * uc - uw - ue) / L.hx - L.hx * f[n];    // :20-21
    } else {
        double v = L.diag * uc - L.cx * (uw + ue);
        if (L.ay) v -= L.cy * (us + un);
        if (L.az) v -= L.cz * (uu + ud);
        F[n] = v - L.vol * f[n];                                   // :57-59, :106-108
    }
}

int launch_fish_sample(cudaStream_t st, const LevelDesc &L, int dim, int problem, double c0, double c1, double c2,
                       double *f, double *gb) {
    const long long n = L.nlocal();
    if (n <= 0) return 0;
    fish_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, dim, problem, c0, c1, c2, f, gb);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_initial_state(cudaStream_t st, const LevelDesc &L, const double *gb, int gonboundary, double *u) {
    const long long n = L.nlocal();
    if (n <= 0) return 0;
    initial_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, gb, gonboundary, u);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_poisson_function(cudaStream_t st, const LevelDesc &L, int dim, double c0, const double *u, const double *f,
                            const double *gb, double *F) {
    const long long n = L.nlocal();
    if (n <= 0) return 0;
    poisson_function_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, dim, c0, u, f, gb, F);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
