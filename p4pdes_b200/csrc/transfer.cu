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
    prolong_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(F, C, xc, xf);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
