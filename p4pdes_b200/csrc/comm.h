// comm.h -- peer-memory communication structures (see comm.cu).
#pragma once
#include "common.cuh"

namespace p4b {

constexpr int MAX_RANKS = 16;

struct RedSlot {
    double v[3];
    unsigned long long epoch;
};

// One per rank, in that rank's device memory, mapped into every peer through CUDA IPC.
struct Mailbox {
    unsigned long long halo_flag[2];            // [0] written by rank-1, [1] written by rank+1
    unsigned long long bar_flag[MAX_RANKS];     // barrier epochs, slot [r] written by rank r
    RedSlot red[2][MAX_RANKS];                  // allreduce slots, double buffered by epoch parity
};

// private to a rank
struct LocalSync {
    unsigned long long halo_epoch, bar_epoch, red_epoch;
    unsigned int done;
    unsigned int pdone;      // CTAs of the running kernel that have finished their fused push (HaloPort)
};

struct PeerTable {
    int rank, nranks;
    Mailbox *mbox[MAX_RANKS];                   // mbox[rank] is the local one
};

struct GatherTable {
    double *base[MAX_RANKS];                    // arena base of every rank (peer mapped)
};

// ---------------------------------------------------------------------------------------------------------
// Fused ghost exchange.  A HaloPort rides along with a kernel that writes a ghosted vector of a distributed level:
//   * port_wait   : the CTAs that touch the slab boundary (they read a ghost plane or store into a neighbour's) wait
//                 until the neighbours' flags have reached this rank's exchange count: every earlier push has landed,
//                 so ghost planes may be read, and -- because a rank signals only after ALL its boundary CTAs are
//                 done -- nothing on the neighbours still reads the ghost planes this kernel is about to overwrite
//                 (mg.cu keeps the one exception, the same vector exchanged twice in a row, apart with a flag barrier);
//   * port_store  : where the kernel stores an element of its first / last owned plane, the same value goes straight
//                 into the neighbour's ghost plane over NVLink;
//   * port_signal : the last boundary CTA to finish bumps the exchange count and releases it to both neighbours.
// An exchange therefore costs no kernel of its own: the copy rides on the producer and the wait on the consumer, and
// interior CTAs never see it.
// All zero = inactive (single rank, replicated level, NCCL transport).
// ---------------------------------------------------------------------------------------------------------
struct HaloPort {
    LocalSync *sync;                            // nullptr = inactive
    const unsigned long long *my_flags;         // this rank's halo_flag[2]: [0] set by rank-1, [1] by rank+1
    unsigned long long *flag_lo, *flag_hi;      // the neighbours' flag that this rank sets (nullptr = no neighbour)
    double *lo_dst, *hi_dst;                    // the neighbours' ghost planes of the vector written (nullptr = none)
    long long plane;                            // doubles per plane of that vector
    long long hi_start;                         // local index of the first element of its last owned plane
    int push;                                   // 1 = this kernel pushes and advances the exchange count
};

#ifdef __CUDACC__
__device__ __forceinline__ void port_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long port_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Only "boundary" CTAs -- those that read a ghost plane or store into a neighbour's -- take part in the protocol;
// `boundary` must be uniform over the CTA, and the same CTAs must call port_signal with it.
// All their threads, before the first ghost read / peer store:
__device__ __forceinline__ void port_wait(const HaloPort &hp, bool boundary) {
    if (!hp.sync || !boundary) return;
    if (threadIdx.x == 0) {
        const unsigned long long e = *(volatile unsigned long long *)&hp.sync->halo_epoch;
        if (hp.flag_lo) while (port_ld_acquire_sys(&hp.my_flags[0]) < e) { }
        if (hp.flag_hi) while (port_ld_acquire_sys(&hp.my_flags[1]) < e) { }
        asm volatile("fence.proxy.async;" ::: "memory");     // ghost planes may be read by bulk (TMA) copies
    }
    __syncthreads();
}
// idx = local index (first owned element = 0) of the element just computed
__device__ __forceinline__ void port_store(const HaloPort &hp, long long idx, double v) {
    if (hp.lo_dst && idx < hp.plane) hp.lo_dst[idx] = v;
    if (hp.hi_dst && idx >= hp.hi_start) hp.hi_dst[idx - hp.hi_start] = v;
}
// All threads of the boundary CTAs, after their last peer store; nboundary = number of boundary CTAs of the grid.
// The last of them to arrive bumps the exchange count and releases it to both neighbours.
// (stored = false: this CTA is only counted, it has no peer stores of its own to publish)
__device__ __forceinline__ void port_signal(const HaloPort &hp, bool boundary, unsigned int nboundary,
                                            bool stored = true) {
    if (!hp.sync || !hp.push || !boundary) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (stored) __threadfence_system();
        if (atomicAdd(&hp.sync->pdone, 1u) == nboundary - 1u) {
            __threadfence_system();
            const unsigned long long e = hp.sync->halo_epoch + 1ull;
            if (hp.flag_lo) port_st_release_sys(hp.flag_lo, e);
            if (hp.flag_hi) port_st_release_sys(hp.flag_hi, e);
            hp.sync->halo_epoch = e;
            hp.sync->pdone = 0u;
            __threadfence();
        }
    }
}
// Boundary CTAs of a kernel whose CTA b owns the `per` consecutive elements starting at b * per (one pass, no grid
// stride) of a slab of n elements: those that touch the first plane (when there is a lower neighbour) or the last.
struct PortSpan { bool boundary; unsigned int nboundary; };
__device__ __forceinline__ PortSpan port_span_linear(const HaloPort &hp, long long plane, long long n, long long per) {
    PortSpan s = {false, 0u};
    if (!hp.sync) return s;
    const long long nb = (n + per - 1) / per;
    const long long a = hp.flag_lo ? min((plane + per - 1) / per, nb) : 0;        // blocks [0, a) touch the first plane
    const long long h = hp.flag_hi ? (n - plane) / per : nb;                       // blocks [h, nb) touch the last plane
    const long long b = (long long)blockIdx.x;
    s.boundary = b < a || b >= h;
    s.nboundary = (unsigned int)(a + (nb - h) - max(0LL, a - h));
    return s;
}
#endif

int launch_halo_push(cudaStream_t st, const double *lo_src, double *lo_dst, const double *hi_src, double *hi_dst,
                     long long plane, unsigned long long *flag_prev, unsigned long long *flag_next,
                     const unsigned long long *my_flags, LocalSync *sync);
int launch_allreduce(cudaStream_t st, double *vals, int nv, int op_max, const PeerTable &peers, LocalSync *sync);
int launch_barrier(cudaStream_t st, int all, const PeerTable &peers, LocalSync *sync);
int launch_gather_push(cudaStream_t st, const double *src, long long n, long long off_doubles, const GatherTable &dst,
                       int rank, int nranks);

}  // namespace p4b
