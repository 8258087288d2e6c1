// vecops.cu -- Krylov vector kernels (K5 of SURVEY.md 2.3).
//
// [PETSc] KSPSolve_CG calls VecTDot / VecAXPY / VecNorm / VecAYPX as separate passes; here the
// CG scalars stay on the device (alpha = beta/(p,w) is formed inside the consuming kernel from
// two device doubles), x += a p and r -= a w share one pass, and (z,z),(z,r) share one pass.
// All reductions are fixed-order (common.cuh: grid_sum_finalize), no floating-point atomics.
#include "kernels.h"

namespace p4b {

static const int VT = 256;            // threads per block
static const int VU = 4;              // elements per thread per grid-stride step

static inline unsigned vec_blocks(long long n, int cap) {
    long long nb = (n + (long long)VT * VU - 1) / ((long long)VT * VU);
    if (nb < 1) nb = 1;
    if (nb > cap) nb = cap;
    return (unsigned)nb;
}
static const int STREAM_BLOCKS = 148 * 8;

__global__ void __launch_bounds__(VT) dot2_kernel(long long n, const double *__restrict__ x,
                                                   const double *__restrict__ y, double *partials,
                                                   unsigned int *ticket, double *out2) {
    double v[2] = {0.0, 0.0};   // (x,x), (x,y)
    const long long stride = (long long)gridDim.x * VT * VU;
    for (long long base = (long long)blockIdx.x * VT * VU + threadIdx.x; base < n; base += stride) {
        double a[VU], b[VU];
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            a[q] = (i < n) ? x[i] : 0.0;
            b[q] = (i < n) ? y[i] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < VU; q++) {
            v[0] += a[q] * a[q];
            v[1] += a[q] * b[q];
        }
    }
    grid_sum_finalize<2, VT>(v, partials, ticket, out2);
}

__global__ void __launch_bounds__(VT) dot1_kernel(long long n, const double *__restrict__ x,
                                                   const double *__restrict__ y, double *partials,
                                                   unsigned int *ticket, double *out1) {
    double v[1] = {0.0};
    const long long stride = (long long)gridDim.x * VT * VU;
    for (long long base = (long long)blockIdx.x * VT * VU + threadIdx.x; base < n; base += stride) {
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) v[0] += x[i] * y[i];
        }
    }
    grid_sum_finalize<1, VT>(v, partials, ticket, out1);
}

// sum_i ((x_i - y_i) / (atol + rtol max(|x_i|, |y_i|)))^2 : the weighted error of [PETSc] TSErrorWeightedNorm2 (the
// local-truncation-error estimate of the adaptive time steppers); same fixed-order reduction as the dot products
__global__ void __launch_bounds__(VT) wrms_kernel(long long n, const double *__restrict__ x, const double *__restrict__ y,
                                                   double atol, double rtol, double *partials, unsigned int *ticket,
                                                   double *out1) {
    double v[1] = {0.0};
    const long long stride = (long long)gridDim.x * VT * VU;
    for (long long base = (long long)blockIdx.x * VT * VU + threadIdx.x; base < n; base += stride) {
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) {
                const double a = x[i], b = y[i];
                const double e = (a - b) / (atol + rtol * fmax(fabs(a), fabs(b)));
                v[0] += e * e;
            }
        }
    }
    grid_sum_finalize<1, VT>(v, partials, ticket, out1);
}

// max |x_i| via the same two-stage scheme (max is order independent, so plain block max + final pass)
__global__ void __launch_bounds__(VT) absmax_kernel(long long n, const double *__restrict__ x, double *partials,
                                                     unsigned int *ticket, double *out1) {
    __shared__ double sm[VT / 32];
    __shared__ bool is_last;
    double m = 0.0;
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) m = fmax(m, fabs(x[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < VT / 32; w++) m = fmax(m, sm[w]);
        partials[blockIdx.x] = m;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double r = 0.0;
        for (unsigned int i = 0; i < gridDim.x; i++) r = fmax(r, ((volatile double *)partials)[i]);
        out1[0] = r;
        *ticket = 0u;
    }
}

__global__ void __launch_bounds__(VT) axpy2_kernel(long long n, const double *__restrict__ num,
                                                    const double *__restrict__ den, const double *__restrict__ p,
                                                    const double *__restrict__ w, double *__restrict__ x,
                                                    double *__restrict__ r) {
    const double a = num[0] / den[0];
    const long long stride = (long long)gridDim.x * VT * VU;
    for (long long base = (long long)blockIdx.x * VT * VU + threadIdx.x; base < n; base += stride) {
        double pv[VU], wv[VU], xv[VU], rv[VU];
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) { pv[q] = p[i]; wv[q] = w[i]; xv[q] = x[i]; rv[q] = r[i]; }
        }
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) { x[i] = xv[q] + a * pv[q]; r[i] = rv[q] - a * wv[q]; }
        }
    }
}

__global__ void __launch_bounds__(VT) aypx_dev_kernel(long long n, const double *__restrict__ num,
                                                       const double *__restrict__ den, const double *__restrict__ z,
                                                       double *__restrict__ p, int first) {
    const double b = first ? 0.0 : num[0] / den[0];
    const long long stride = (long long)gridDim.x * VT * VU;
    for (long long base = (long long)blockIdx.x * VT * VU + threadIdx.x; base < n; base += stride) {
        double zv[VU], pv[VU];
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) { zv[q] = z[i]; pv[q] = first ? 0.0 : p[i]; }
        }
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) p[i] = first ? zv[q] : zv[q] + b * pv[q];
        }
    }
}

// One pass for the CG direction update AND the previous iteration's solution update:
//   x += a_prev * p   (a_prev = an/ad, the alpha of the iteration that produced this p; skipped when first)
//   p  = z + b * p    (b = bn/bd)
// p is read once for both, which saves the separate 24 N pass x += a p of KSPSolve_CG.
template <bool MG>
__global__ void __launch_bounds__(VT) xp_update_kernel(long long n, const double *__restrict__ an,
                                                        const double *__restrict__ ad, const double *__restrict__ bn,
                                                        const double *__restrict__ bd, const double *__restrict__ z,
                                                        double *__restrict__ p, double *__restrict__ x, int first,
                                                        const HaloPort port) {
    // boundary planes of p first (comm.h SegRot): the CTAs that write them wait for the neighbours before, and
    // signal right after, their last boundary segment
    const SegRot rot = seg_rot(MG ? port : HaloPort(), n, VT * VU);      // (MG = false: the exchange code compiles out)
    const bool bcta = MG && (long long)blockIdx.x < rot.NB;
    if (MG) port_wait(port, bcta);
    const double a = first ? 0.0 : an[0] / ad[0];
    const double b = first ? 0.0 : bn[0] / bd[0];
    for (long long g = blockIdx.x; g < rot.S; g += gridDim.x) {
        const long long base = rot.phys(g) * (VT * VU) + threadIdx.x;
        double zv[VU], pv[VU], xv[VU];
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) { zv[q] = z[i]; pv[q] = first ? 0.0 : p[i]; xv[q] = first ? 0.0 : x[i]; }
        }
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) {
                if (!first) x[i] = xv[q] + a * pv[q];
                const double pn = first ? zv[q] : zv[q] + b * pv[q];
                p[i] = pn;
                if (MG) port_store(port, i, pn);
            }
        }
        if (MG && bcta && g < rot.NB && g + gridDim.x >= rot.NB) {       // this CTA's last boundary segment
            __syncthreads();
            port_signal_nosync(port, (unsigned int)min((long long)gridDim.x, rot.NB), threadIdx.x == 0);
        }
    }
}

// r -= a w  with a = num/den on the device
template <bool MG>
__global__ void __launch_bounds__(VT) r_update_kernel(long long n, const double *__restrict__ num,
                                                       const double *__restrict__ den, const double *__restrict__ w,
                                                       double *__restrict__ r, const HaloPort port) {
    const SegRot rot = seg_rot(MG ? port : HaloPort(), n, VT * VU);      // boundary planes of r first, see xp_update_kernel
    const bool bcta = MG && (long long)blockIdx.x < rot.NB;
    if (MG) port_wait(port, bcta);
    const double a = num[0] / den[0];
    for (long long g = blockIdx.x; g < rot.S; g += gridDim.x) {
        const long long base = rot.phys(g) * (VT * VU) + threadIdx.x;
        double wv[VU], rv[VU];
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) { wv[q] = w[i]; rv[q] = r[i]; }
        }
#pragma unroll
        for (int q = 0; q < VU; q++) {
            const long long i = base + (long long)q * VT;
            if (i < n) {
                const double rn = rv[q] - a * wv[q];
                r[i] = rn;
                if (MG) port_store(port, i, rn);
            }
        }
        if (MG && bcta && g < rot.NB && g + gridDim.x >= rot.NB) {
            __syncthreads();
            port_signal_nosync(port, (unsigned int)min((long long)gridDim.x, rot.NB), threadIdx.x == 0);
        }
    }
}

// out = a x + b y ; x or y may be null (treated as zero) ; out may alias x or y
__global__ void __launch_bounds__(VT) axpby_kernel(long long n, double a, const double *x, double b, const double *y,
                                                    double *out) {
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) {
        double v = 0.0;
        if (x) v = a * x[i];
        if (y) v += b * y[i];
        if (!x && !y) v = a;
        out[i] = v;
    }
}

// ---- reduced-space variational inequalities ([PETSc] SNESVINEWTONRSLS under c/ch12/obstacle.c) --------
// mask_i = 0 where the lower bound is active (u_i <= lo_i + 1e-8 and F_i > 0, [PETSc] vi.c / obstacle.c:196-205), else 1
__global__ void __launch_bounds__(VT) vi_mask_kernel(long long n, const double *__restrict__ u, const double *__restrict__ lo,
                                                      const double *__restrict__ F, double *__restrict__ mask) {
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride)
        mask[i] = (u[i] <= lo[i] + 1.0e-8 && F[i] > 0.0) ? 0.0 : 1.0;
}
// out = x .* y   (op 0)     out = max(x, y)   (op 1: projection onto the bound)
__global__ void __launch_bounds__(VT) pointwise_kernel(long long n, int op, const double *x, const double *y, double *out) {
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride)
        out[i] = op ? fmax(x[i], y[i]) : x[i] * y[i];
}
int launch_vi_mask(cudaStream_t st, long long n, const double *u, const double *lo, const double *F, double *mask) {
    if (n <= 0) return 0;
    long long nb = (n + VT - 1) / VT;
    if (nb > STREAM_BLOCKS * 4) nb = STREAM_BLOCKS * 4;
    vi_mask_kernel<<<(unsigned)nb, VT, 0, st>>>(n, u, lo, F, mask);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_pointwise(cudaStream_t st, long long n, int op, const double *x, const double *y, double *out) {
    if (n <= 0) return 0;
    long long nb = (n + VT - 1) / VT;
    if (nb > STREAM_BLOCKS * 4) nb = STREAM_BLOCKS * 4;
    pointwise_kernel<<<(unsigned)nb, VT, 0, st>>>(n, op, x, y, out);
    P4B_LAUNCH_CHECK();
    return 0;
}

#define RED_CHECK(nb)                                                                                   \
    if ((int)(nb) > red.max_blocks) return fail(63, "reduction: %u blocks exceed scratch %d", (unsigned)(nb), red.max_blocks)

int launch_dot2(cudaStream_t st, long long n, const double *x, const double *y, double *out2, const Reducer &red) {
    unsigned nb = vec_blocks(n, STREAM_BLOCKS);
    RED_CHECK(nb);
    dot2_kernel<<<nb, VT, 0, st>>>(n, x, y, red.partials, red.ticket, out2);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_dotn(cudaStream_t st, long long n, const double *x, const double *y, double *out1, const Reducer &red) {
    unsigned nb = vec_blocks(n, STREAM_BLOCKS);
    RED_CHECK(nb);
    dot1_kernel<<<nb, VT, 0, st>>>(n, x, y, red.partials, red.ticket, out1);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_wrms(cudaStream_t st, long long n, const double *x, const double *y, double atol, double rtol, double *out1,
                const Reducer &red) {
    unsigned nb = vec_blocks(n, STREAM_BLOCKS);
    RED_CHECK(nb);
    wrms_kernel<<<nb, VT, 0, st>>>(n, x, y, atol, rtol, red.partials, red.ticket, out1);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_absmax(cudaStream_t st, long long n, const double *x, double *out1, const Reducer &red) {
    unsigned nb = vec_blocks(n, STREAM_BLOCKS);
    RED_CHECK(nb);
    absmax_kernel<<<nb, VT, 0, st>>>(n, x, red.partials, red.ticket, out1);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_axpy2(cudaStream_t st, long long n, const double *num, const double *den, const double *p,
                 const double *w, double *x, double *r) {
    if (n <= 0) return 0;
    axpy2_kernel<<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, num, den, p, w, x, r);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_aypx_dev(cudaStream_t st, long long n, const double *num, const double *den, const double *z, double *p,
                    int first) {
    if (n <= 0) return 0;
    aypx_dev_kernel<<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, num, den, z, p, first);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_xp_update(cudaStream_t st, long long n, const double *an, const double *ad, const double *bn, const double *bd,
                     const double *z, double *p, double *x, int first, const HaloPort &port) {
    if (n <= 0) return 0;
    if (port.sync) xp_update_kernel<true><<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, an, ad, bn, bd, z, p, x, first, port);
    else xp_update_kernel<false><<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, an, ad, bn, bd, z, p, x, first, port);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_r_update(cudaStream_t st, long long n, const double *num, const double *den, const double *w, double *r,
                    const HaloPort &port) {
    if (n <= 0) return 0;
    if (port.sync) r_update_kernel<true><<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, num, den, w, r, port);
    else r_update_kernel<false><<<vec_blocks(n, STREAM_BLOCKS), VT, 0, st>>>(n, num, den, w, r, port);
    P4B_LAUNCH_CHECK();
    return 0;
}
// x += (num/den) p
__global__ void __launch_bounds__(VT) x_flush_kernel(long long n, const double *__restrict__ num,
                                                      const double *__restrict__ den, const double *__restrict__ p,
                                                      double *__restrict__ x) {
    const double a = num[0] / den[0];
    const long long stride = (long long)gridDim.x * VT;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) x[i] += a * p[i];
}
int launch_x_flush(cudaStream_t st, long long n, const double *num, const double *den, const double *p, double *x) {
    if (n <= 0) return 0;
    x_flush_kernel<<<vec_blocks(n, STREAM_BLOCKS * 4), VT, 0, st>>>(n, num, den, p, x);
    P4B_LAUNCH_CHECK();
    return 0;
}

static int launch_axpby(cudaStream_t st, long long n, double a, const double *x, double b, const double *y,
                        double *out) {
    if (n <= 0) return 0;
    long long nb = (n + VT - 1) / VT;
    if (nb > STREAM_BLOCKS * 4) nb = STREAM_BLOCKS * 4;
    axpby_kernel<<<(unsigned)nb, VT, 0, st>>>(n, a, x, b, y, out);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_axpy(cudaStream_t st, long long n, double a, const double *x, double *y) {
    return launch_axpby(st, n, a, x, 1.0, y, y);
}
int launch_aypx(cudaStream_t st, long long n, double a, const double *x, double *y) {
    return launch_axpby(st, n, 1.0, x, a, y, y);
}
int launch_set(cudaStream_t st, long long n, double a, double *y) {
    return launch_axpby(st, n, a, nullptr, 0.0, nullptr, y);
}
int launch_scale_copy(cudaStream_t st, long long n, double a, const double *x, double *y) {
    return launch_axpby(st, n, a, x, 0.0, nullptr, y);
}
int launch_axpby_out(cudaStream_t st, long long n, double a, const double *x, double b, const double *y,
                     double *out) {
    return launch_axpby(st, n, a, x, b, y, out);
}

// ---------------------------------------------------------------------------------------------
// coarse solve: x = Ainv b, one warp per row (K9).  [PETSc] PCLU / PCREDUNDANT on level 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_matvec_kernel(int n, const double *__restrict__ Ainv,
                                                            const double *__restrict__ b, double *__restrict__ x) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *a = Ainv + (size_t)row * n;
    double s = 0.0;
    for (int c = lane; c < n; c += 32) s += a[c] * b[c];
    s = warp_sum(s);
    if (lane == 0) x[row] = s;
}

int launch_dense_matvec(cudaStream_t st, int n, const double *Ainv, const double *b, double *x) {
    dense_matvec_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, Ainv, b, x);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
