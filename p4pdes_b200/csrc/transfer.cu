// transfer.cu -- DMDA Q1 grid transfer (K3, K4 of SURVEY.md 2.3).
//
// [PETSc] DMCreateInterpolation on a DMDA gives the vertex-centred bi/tri-linear P (ratio 2):
// a fine node that coincides with a coarse node copies it (weight 1), a node on a coarse edge
// averages 2 (1/2 each), on a face 4 (1/4), in a cell 8 (1/8).  Boundary nodes are not special.
// PCMG restricts with R = P^T, un-normalised.  Both are matrix-free here.
#include "kernels.h"

namespace p4b {

// b_c(I,J,K) = sum_{d in {-1,0,1}^3} w(di) w(dj) w(dk) r(2I+di, 2J+dj, 2K+dk),  w(0)=1, w(+-1)=1/2,
// fine indices outside the grid skipped; inactive slots contribute offset 0 only.
__global__ void __launch_bounds__(256) restrict_kernel(const LevelDesc F, const LevelDesc C,
                                                        const double *__restrict__ rf, double *__restrict__ bc) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= C.nlocal()) return;
    const int cplane = C.nx * C.ny;
    const int Kl = (int)(n / cplane);
    const int rem = (int)(n - (long long)Kl * cplane);
    const int J = rem / C.nx;
    const int I = rem - J * C.nx;
    const int K = Kl + C.zs;
    const int fi = F.ax ? 2 * I : I, fj = F.ay ? 2 * J : J, fk = F.az ? 2 * K : K;
    const int ri = F.ax ? 1 : 0, rj = F.ay ? 1 : 0, rk = F.az ? 1 : 0;
    const long long fplane = (long long)F.nx * F.ny;
    double s = 0.0;
    for (int dk = -rk; dk <= rk; dk++) {
        const int kf = fk + dk;
        if (kf < 0 || kf >= F.nz) continue;
        const double wk = dk ? 0.5 : 1.0;
        double sk = 0.0;
        for (int dj = -rj; dj <= rj; dj++) {
            const int jf = fj + dj;
            if (jf < 0 || jf >= F.ny) continue;
            const double wj = dj ? 0.5 : 1.0;
            const double *row = rf + ((long long)(kf - F.zs) * fplane + (long long)jf * F.nx);
            double sj = row[fi];
            if (ri) {
                double e = 0.0;
                if (fi - 1 >= 0) e += row[fi - 1];
                if (fi + 1 < F.nx) e += row[fi + 1];
                sj += 0.5 * e;
            }
            sk += wj * sj;
        }
        s += wk * sk;
    }
    bc[n] = s;
}

// x_f(i,j,k) += sum over the (<= 8) coarse parents of their Q1 weights times x_c.
__global__ void __launch_bounds__(256) prolong_add_kernel(const LevelDesc F, const LevelDesc C,
                                                           const double *__restrict__ xc, double *__restrict__ xf) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    if (n >= F.nlocal()) return;
    const int fplane = F.nx * F.ny;
    const int kl = (int)(n / fplane);
    const int rem = (int)(n - (long long)kl * fplane);
    const int j = rem / F.nx;
    const int i = rem - j * F.nx;
    const int k = kl + F.zs;
    // per slot: first parent index and whether a second parent (index+1) shares the weight
    const int I0 = F.ax ? (i >> 1) : i, oi = F.ax ? (i & 1) : 0;
    const int J0 = F.ay ? (j >> 1) : j, oj = F.ay ? (j & 1) : 0;
    const int K0 = F.az ? (k >> 1) : k, ok = F.az ? (k & 1) : 0;
    const long long cplane = (long long)C.nx * C.ny;
    double s = 0.0;
    for (int dk = 0; dk <= ok; dk++) {
        double sk = 0.0;
        for (int dj = 0; dj <= oj; dj++) {
            const double *row = xc + ((long long)(K0 + dk - C.zs) * cplane + (long long)(J0 + dj) * C.nx);
            double sj = row[I0];
            if (oi) sj = 0.5 * (sj + row[I0 + 1]);
            sk += sj;
        }
        if (oj) sk *= 0.5;
        s += sk;
    }
    if (ok) s *= 0.5;
    xf[n] += s;
}

// 3-D fast path.  One thread per (i, J, K): it interpolates the x direction once per coarse row
// (a_jk = 1/2 (c[I0] + c[I1]), I1 = I0 for even i, which is exact) and then updates the up-to-four fine
// nodes (i, 2J+{0,1}, 2K+{0,1}): 8 coarse loads (L1/L2 hits, neighbouring lanes share them) per 4 fine
// nodes instead of up to 8 per node, fine accesses fully coalesced, one index division per 4 nodes.
__global__ void __launch_bounds__(256) prolong_add3d_kernel(const LevelDesc F, const LevelDesc C, int Kfirst,
                                                            const double *__restrict__ xc, double *__restrict__ xf) {
    const int task = blockIdx.x * 256 + threadIdx.x;       // over nx * cny
    if (task >= F.nx * C.ny) return;
    const int J = task / F.nx, i = task - J * F.nx;
    const int K = Kfirst + blockIdx.y;
    const int I0 = i >> 1, I1 = I0 + (i & 1);
    const int J1 = min(J + 1, C.ny - 1), K1 = min(K + 1, C.nz - 1);
    const long long cplane = (long long)C.nx * C.ny;
    const double *c00 = xc + ((long long)(K - C.zs) * cplane + (long long)J * C.nx);
    const double *c10 = xc + ((long long)(K - C.zs) * cplane + (long long)J1 * C.nx);
    const double *c01 = xc + ((long long)(K1 - C.zs) * cplane + (long long)J * C.nx);
    const double *c11 = xc + ((long long)(K1 - C.zs) * cplane + (long long)J1 * C.nx);
    const double a00 = 0.5 * (c00[I0] + c00[I1]);
    const double a10 = 0.5 * (c10[I0] + c10[I1]);
    const double a01 = 0.5 * (c01[I0] + c01[I1]);
    const double a11 = 0.5 * (c11[I0] + c11[I1]);
    const int j0 = 2 * J, k0 = 2 * K;
    const bool jok = (j0 + 1 < F.ny);
    const long long fplane = (long long)F.nx * F.ny;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int k = k0 + c;
        if (k < F.zs || k >= F.zs + F.zm) continue;        // only planes this rank owns (and k <= nz-1)
        double *row = xf + ((long long)(k - F.zs) * fplane + (long long)j0 * F.nx + i);
        const double e0 = c ? 0.5 * (a00 + a01) : a00;
        row[0] += e0;
        if (jok) {
            const double e1 = c ? 0.25 * ((a00 + a10) + (a01 + a11)) : 0.5 * (a00 + a10);
            row[F.nx] += e1;
        }
    }
}

int launch_restrict(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *rf, double *bc) {
    const long long n = C.nlocal();
    if (n <= 0) return 0;
    restrict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(F, C, rf, bc);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_prolong_add(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *xc, double *xf) {
    const long long n = F.nlocal();
    if (n <= 0) return 0;
    if (F.ax && F.ay && F.az && F.nx >= 64 && (long long)F.nx * C.ny < (1LL << 30)) {
        const int Kfirst = F.zs / 2, Klast = (F.zs + F.zm - 1) / 2;
        dim3 grid((unsigned)(((long long)F.nx * C.ny + 255) / 256), (unsigned)(Klast - Kfirst + 1));
        prolong_add3d_kernel<<<grid, 256, 0, st>>>(F, C, Kfirst, xc, xf);
        P4B_LAUNCH_CHECK();
        return 0;
    }
    prolong_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(F, C, xc, xf);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
