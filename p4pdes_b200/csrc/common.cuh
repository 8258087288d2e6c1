// common.cuh -- shared descriptors and device helpers for the p4b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "p4b200.h"

namespace p4b {

// One multigrid level as the kernels see it.  Every grid is carried in three slots
// (nx fastest, nz slowest).  A slot is "active" when it is a PDE dimension (it has
// Dirichlet boundaries and stencil coupling); inactive slots have extent 1.
//   3-D fish: (mx,my,mz) all active        2-D: (mx,1,my), y slot inactive        1-D: (mx,1,1)
// so the slowest active dimension is always slot z: that is the slab (multi-GPU) dimension,
// and the memory layout idx = (k*ny + j)*nx + i equals the DMDA natural ordering.
struct LevelDesc {
    int nx, ny, nz;          // global extents
    int ax, ay, az;          // 1 = active slot
    int zs, zm;              // planes [zs, zs+zm) of slot z live on this device
    double cx, cy, cz;       // off-diagonal magnitudes sc_d (poissonfunctions.c:38-40,78-81)
    double diag;             // constant diagonal scdiag (poissonfunctions.h:36-38)
    double vol;              // cell volume: the factor on f_rhs in the residual
    double hx, hy, hz;       // spacings (slot order)
    __host__ __device__ long long plane() const { return (long long)nx * ny; }
    __host__ __device__ long long nlocal() const { return (long long)nx * ny * zm; }
    __host__ __device__ long long nglobal() const { return (long long)nx * ny * nz; }
};

// error plumbing -----------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const char *fmt, ...);

#define P4B_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return p4b::fail(70 + (int)e_ % 20, "%s:%d CUDA error %d (%s) in %s", __FILE__,     \
                             __LINE__, (int)e_, cudaGetErrorString(e_), #call);                 \
    } while (0)

#define P4B_CHECK(call)              \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != 0) return rc_;    \
    } while (0)

// every kernel launch goes through this: counts launches (p4b_launch_count) and checks the launch
extern long long g_launch_count;
#define P4B_LAUNCH_CHECK()              \
    do {                                \
        p4b::g_launch_count++;          \
        P4B_CUDA(cudaGetLastError());   \
    } while (0)

// device helpers -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid-wide sum of NV per-thread values.
//   1. warp shuffle tree, 2. per-block tree in shared memory, 3. one partial per block is
//   written to `partials`, 4. the last block to finish (ticket counter) adds the partials in a
//   fixed order and stores the NV results.  For a fixed launch geometry the result is
//   bit-reproducible (no floating-point atomics).  partials holds NV*gridDim.x doubles.
template <int NV, int NT>
__device__ __forceinline__ void grid_sum_finalize(double (&v)[NV], double *partials, unsigned int *ticket,
                                                  double *out) {
    __shared__ double sm[NV][NT / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        double s = warp_sum(v[q]);
        if (lane == 0) sm[q][wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double s = (lane < NT / 32) ? sm[q][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) partials[(size_t)q * gridDim.x + blockIdx.x] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double s = 0.0;
            for (unsigned int i = threadIdx.x; i < gridDim.x; i += NT)
                s += ((volatile double *)partials)[(size_t)q * gridDim.x + i];
            s = warp_sum(s);
            __syncthreads();
            if (lane == 0) sm[q][wid] = s;
            __syncthreads();
            if (wid == 0) {
                double t2 = (lane < NT / 32) ? sm[q][lane] : 0.0;
                t2 = warp_sum(t2);
                if (lane == 0) out[q] = t2;
            }
        }
        if (threadIdx.x == 0) *ticket = 0u;   // re-arm for the next launch on this stream
    }
}

}  // namespace p4b
