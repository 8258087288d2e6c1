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
                                                        const double *__restrict__ rf, double *__restrict__ bc,
                                                        const HaloPort port) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    const PortSpan span = port_span_linear(port, C.plane(), C.nlocal(), 256);
    port_wait(port, span.boundary);
    if (n < C.nlocal()) {
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
    port_store(port, n, s);
    }
    port_signal(port, span.boundary, span.nboundary);
}

// x_f(i,j,k) += sum over the (<= 8) coarse parents of their Q1 weights times x_c.
__global__ void __launch_bounds__(256) prolong_add_kernel(const LevelDesc F, const LevelDesc C,
                                                           const double *__restrict__ xc, double *__restrict__ xf,
                                                           const HaloPort port) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    const PortSpan span = port_span_linear(port, F.plane(), F.nlocal(), 256);
    port_wait(port, span.boundary);
    if (n < F.nlocal()) {
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
    const double v = xf[n] + s;
    xf[n] = v;
    port_store(port, n, v);
    }
    port_signal(port, span.boundary, span.nboundary);
}

// 3-D fast path.  One thread per (i, J, K): it interpolates the x direction once per coarse row
// (a_jk = 1/2 (c[I0] + c[I1]), I1 = I0 for even i, which is exact) and then updates the up-to-four fine
// nodes (i, 2J+{0,1}, 2K+{0,1}): 8 coarse loads (L1/L2 hits, neighbouring lanes share them) per 4 fine
// nodes instead of up to 8 per node, fine accesses fully coalesced, one index division per 4 nodes.
template <bool MG>
__global__ void __launch_bounds__(256) prolong_add3d_kernel(const LevelDesc F, const LevelDesc C, int Kfirst,
                                                            const double *__restrict__ xc, double *__restrict__ xf,
                                                            const HaloPort port) {
    const int task = blockIdx.x * 256 + threadIdx.x;       // over nx * cny
    // the first / last coarse plane of the range are the ones that may be a neighbour's (coarse ghost read) and
    // hold the first / last owned fine plane (pushed)
    const unsigned int by = MG ? port_remap(port, blockIdx.y, gridDim.y) : blockIdx.y;
    const bool port_cta = MG && port.sync != nullptr && ((by == 0 && port.flag_lo != nullptr) ||
                                                         (by == gridDim.y - 1 && port.flag_hi != nullptr));
    if (MG) port_wait(port, port_cta);
    if (task < F.nx * C.ny) {
    const int J = task / F.nx, i = task - J * F.nx;
    const int K = Kfirst + (int)by;
    const int I0 = i >> 1, I1 = I0 + (i & 1);
    const int J1 = min(J + 1, C.ny - 1), K1 = min(K + 1, C.nz - 1);
    const long long cplane = (long long)C.nx * C.ny;
    const double *c00 = xc + ((long long)(K - C.zs) * cplane + (long long)J * C.nx);
    const double *c10 = xc + ((long long)(K - C.zs) * cplane + (long long)J1 * C.nx);
    const double *c01 = xc + ((long long)(K1 - C.zs) * cplane + (long long)J * C.nx);
    const double *c11 = xc + ((long long)(K1 - C.zs) * cplane + (long long)J1 * C.nx);
    const int j0 = 2 * J, k0 = 2 * K;
    const bool jok = (j0 + 1 < F.ny);
    const long long fplane = (long long)F.nx * F.ny;
    // the four fine values are loaded before any coarse value is needed (all 12 loads in flight together)
    double *row[2];
    bool own[2];
    double f[2][2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int k = k0 + c;
        own[c] = (k >= F.zs && k < F.zs + F.zm);            // only planes this rank owns (and k <= nz-1)
        row[c] = xf + ((long long)(k - F.zs) * fplane + (long long)j0 * F.nx + i);
        f[c][0] = own[c] ? row[c][0] : 0.0;
        f[c][1] = (own[c] && jok) ? row[c][F.nx] : 0.0;
    }
    const double a00 = 0.5 * (c00[I0] + c00[I1]);
    const double a10 = 0.5 * (c10[I0] + c10[I1]);
    const double a01 = 0.5 * (c01[I0] + c01[I1]);
    const double a11 = 0.5 * (c11[I0] + c11[I1]);
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (!own[c]) continue;
        const double e0 = c ? 0.5 * (a00 + a01) : a00;
        const double v0 = f[c][0] + e0;
        row[c][0] = v0;
        if (jok) {
            const double e1 = c ? 0.25 * ((a00 + a10) + (a01 + a11)) : 0.5 * (a00 + a10);
            const double v1 = f[c][1] + e1;
            row[c][F.nx] = v1;
        }
    }
    // first / last owned fine plane: the values also go into the neighbours' ghost planes.  Kept out of the loop above
    // (the HaloPort fields would stay live through it: 48 instead of 32 registers) -- the thread reads back what it stored.
    if (MG && port_cta && port.push) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (!own[c]) continue;
            const int k = k0 + c;
            const long long idx = row[c] - xf;
            if (port.lo_dst && k == F.zs) {
                port.lo_dst[idx] = row[c][0];
                if (jok) port.lo_dst[idx + F.nx] = row[c][F.nx];
            }
            if (port.hi_dst && k == F.zs + F.zm - 1) {
                port.hi_dst[idx - port.hi_start] = row[c][0];
                if (jok) port.hi_dst[idx - port.hi_start + F.nx] = row[c][F.nx];
            }
        }
    }
    }
    if (MG) port_signal(port, port_cta, gridDim.x * (gridDim.y == 1 ? 1u : (port.flag_lo != nullptr) + (port.flag_hi != nullptr)));
}

// 3-D fast path of the restriction.  One thread per coarse column I and strip of TJ coarse rows; it marches a
// chunk of coarse planes.  Per fine plane it loads its 2 TJ + 1 fine rows ONCE (one aligned 16-byte load per lane
// and row, the third point of the x restriction comes from the neighbouring lane by warp shuffle), restricts them
// in x and y in registers, and combines three consecutive fine planes into one coarse plane:
//     rx(row)   = r[2I] + (r[2I-1] + r[2I+1]) / 2
//     P(J, kf)  = rx(2J) + (rx(2J-1) + rx(2J+1)) / 2
//     bc(J, K)  = P(J, 2K) + (P(J, 2K-1) + P(J, 2K+1)) / 2          (missing rows/planes contribute 0)
// The formula per coarse node does not depend on strips or chunks, so any decomposition gives the same bits.
// A fine value is fetched (2 TJ + 1) / (2 TJ) x (2 KC + 1) / (2 KC) = 1.33 times through L1/L2 (TJ = 2, KC = 8),
// against 27/8 = 3.4 times with scattered 8-byte accesses for one thread per coarse node.
template <int TJ>
__device__ __forceinline__ void restrict_plane(const LevelDesc &F, const double *__restrict__ rf, int kf, int J0, int fi,
                                               bool live, int lane, int vparity, double (&P)[TJ]) {
    constexpr int NR = 2 * TJ + 1;
    const long long fplane = (long long)F.nx * F.ny;
    const long long planeoff = (long long)(kf - F.zs) * fplane;
    double rx[NR];
    double2 v[NR];
    double ex[NR];
    bool odd[NR];
    // issue every load of the plane before the first use
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int jf = 2 * J0 - 1 + r;
        const bool rowok = live && jf >= 0 && jf < F.ny;
        const long long rowoff = planeoff + (long long)jf * F.nx;
        const double *row = rf + rowoff;
        // aligned pair that contains fine node 2I: {2I, 2I+1} when the row starts on an even element of the
        // 16-byte aligned vector, {2I-1, 2I} when it starts on an odd one
        odd[r] = ((rowoff + vparity) & 1LL) != 0;
        v[r] = make_double2(0.0, 0.0);
        ex[r] = 0.0;
        if (rowok) {
            if (!odd[r]) {
                if (fi + 1 < F.nx) v[r] = *reinterpret_cast<const double2 *>(row + fi);
                else v[r].x = row[fi];
                if (lane == 0 && fi > 0) ex[r] = row[fi - 1];
            } else {
                if (fi > 0) v[r] = *reinterpret_cast<const double2 *>(row + fi - 1);
                else v[r].y = row[fi];
                if (lane == 31 && fi + 1 < F.nx) ex[r] = row[fi + 1];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
        double lft, ctr, rgt;
        if (!odd[r]) {
            ctr = v[r].x; rgt = v[r].y;
            lft = __shfl_up_sync(0xffffffffu, v[r].y, 1);
            if (lane == 0) lft = ex[r];
        } else {
            lft = v[r].x; ctr = v[r].y;
            rgt = __shfl_down_sync(0xffffffffu, v[r].x, 1);
            if (lane == 31) rgt = ex[r];
            if (fi + 1 >= F.nx) rgt = 0.0;          // the lane to the right is beyond the row
        }
        rx[r] = ctr + 0.5 * (lft + rgt);
    }
#pragma unroll
    for (int jj = 0; jj < TJ; jj++) P[jj] = rx[2 * jj + 1] + 0.5 * (rx[2 * jj] + rx[2 * jj + 2]);
}

template <int TJ, bool MG>
__global__ void __launch_bounds__(128, MG ? 8 : 0) restrict3d_kernel(const LevelDesc F, const LevelDesc C, int KC,
                                                          const double *__restrict__ rf, double *__restrict__ bc,
                                                          const HaloPort port) {
    const int I = blockIdx.x * 128 + threadIdx.x;
    // the first / last chunk of coarse planes read the fine ghost planes and hold the coarse boundary planes (pushed)
    const unsigned int bz = MG ? port_remap(port, blockIdx.z, gridDim.z) : blockIdx.z;
    const bool port_cta = MG && port.sync != nullptr && ((bz == 0 && port.flag_lo != nullptr) ||
                                                         (bz == gridDim.z - 1 && port.flag_hi != nullptr));
    if (MG) port_wait(port, port_cta);
    const int J0 = TJ * blockIdx.y;
    const int Kb = C.zs + KC * (int)bz;                          // first coarse plane of this chunk
    const int Ke = min(Kb + KC, C.zs + C.zm);
    const int lane = threadIdx.x & 31;
    if (!MG && (I & ~31) >= C.nx) return;                        // whole warp beyond the row
    const bool warp_live = (I & ~31) < C.nx;                     // (MG: it stays for the barrier in port_signal)
    const bool live = I < C.nx;
    const int fi = 2 * I;
    const int vparity = (int)(((uintptr_t)rf >> 3) & 1);
    const long long cplane = (long long)C.nx * C.ny;
    auto plane_ok = [&](int kf) { return kf >= 0 && kf < F.nz && kf >= F.zs - 1 && kf <= F.zs + F.zm; };
    double Pm[TJ], Pc[TJ], Pp[TJ];
#pragma unroll
    for (int jj = 0; jj < TJ; jj++) Pm[jj] = 0.0;
    if (warp_live && plane_ok(2 * Kb - 1)) restrict_plane<TJ>(F, rf, 2 * Kb - 1, J0, fi, live, lane, vparity, Pm);
    for (int K = Kb; K < Ke && warp_live; K++) {
        restrict_plane<TJ>(F, rf, 2 * K, J0, fi, live, lane, vparity, Pc);
        if (plane_ok(2 * K + 1)) {
            restrict_plane<TJ>(F, rf, 2 * K + 1, J0, fi, live, lane, vparity, Pp);
        } else {
#pragma unroll
            for (int jj = 0; jj < TJ; jj++) Pp[jj] = 0.0;
        }
        if (live) {
            double *out = bc + (long long)(K - C.zs) * cplane + (long long)J0 * C.nx + I;
#pragma unroll
            for (int jj = 0; jj < TJ; jj++)
                if (J0 + jj < C.ny) {
                    const double v = Pc[jj] + 0.5 * (Pm[jj] + Pp[jj]);
                    out[(long long)jj * C.nx] = v;
                }
        }
#pragma unroll
        for (int jj = 0; jj < TJ; jj++) Pm[jj] = Pp[jj];
    }
    // the coarse boundary planes of the slab also go into the neighbours' ghost planes.  After the march, reading back
    // what this thread stored: with port_store inside the loop the HaloPort fields stay live through it and the kernel
    // needs 96 registers instead of 48 (measured on a thin slab: 0.059 ms against 0.036 ms).
    if (MG && port_cta && port.push && live) {
#pragma unroll
        for (int side = 0; side < 2; side++) {
            double *dst = side ? port.hi_dst : port.lo_dst;
            const int K = side ? C.zs + C.zm - 1 : C.zs;
            if (!dst || K < Kb || K >= Ke) continue;
            const long long off = (long long)(K - C.zs) * cplane + (long long)J0 * C.nx + I;
            const double *src = bc + off;
            dst += side ? off - port.hi_start : off;
#pragma unroll
            for (int jj = 0; jj < TJ; jj++)
                if (J0 + jj < C.ny) dst[(long long)jj * C.nx] = src[(long long)jj * C.nx];
        }
    }
    if (MG) port_signal(port, port_cta, gridDim.x * gridDim.y *
                                            (gridDim.z == 1 ? 1u : (port.flag_lo != nullptr) + (port.flag_hi != nullptr)));
}

int launch_restrict(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *rf, double *bc,
                    const HaloPort &port) {
    const long long n = C.nlocal();
    if (n <= 0) return 0;
    if (F.ax && F.ay && F.az && C.nx >= 64 && (((uintptr_t)rf) & 7) == 0) {
        constexpr int TJ = 2;      // measured (tools/ab_bench.py): 2-row strips, 48 registers, 83 % of HBM peak; 4-row strips 84 registers, 70 %
        const int KC = C.zm > 64 ? 8 : 4;      // thin slabs (multi-GPU): more, smaller chunks keep all SMs busy
        dim3 grid((unsigned)((C.nx + 127) / 128), (unsigned)((C.ny + TJ - 1) / TJ), (unsigned)((C.zm + KC - 1) / KC));
        if (port.sync) restrict3d_kernel<TJ, true><<<grid, 128, 0, st>>>(F, C, KC, rf, bc, port);
        else restrict3d_kernel<TJ, false><<<grid, 128, 0, st>>>(F, C, KC, rf, bc, port);
        P4B_LAUNCH_CHECK();
        return 0;
    }
    restrict_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(F, C, rf, bc, port);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_prolong_add(cudaStream_t st, const LevelDesc &F, const LevelDesc &C, const double *xc, double *xf,
                       const HaloPort &port) {
    const long long n = F.nlocal();
    if (n <= 0) return 0;
    if (F.ax && F.ay && F.az && F.nx >= 64 && (long long)F.nx * C.ny < (1LL << 30)) {
        const int Kfirst = F.zs / 2, Klast = (F.zs + F.zm - 1) / 2;
        dim3 grid((unsigned)(((long long)F.nx * C.ny + 255) / 256), (unsigned)(Klast - Kfirst + 1));
        if (port.sync) prolong_add3d_kernel<true><<<grid, 256, 0, st>>>(F, C, Kfirst, xc, xf, port);
        else prolong_add3d_kernel<false><<<grid, 256, 0, st>>>(F, C, Kfirst, xc, xf, port);
        P4B_LAUNCH_CHECK();
        return 0;
    }
    prolong_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(F, C, xc, xf, port);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
