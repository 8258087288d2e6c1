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
    // exchange statistics (p4b_comm_stats): what the boundary CTAs spent spinning on the neighbours' flags and in the
    // system-scope fence that publishes their peer stores; sums over CTAs of %globaltimer nanoseconds
    unsigned long long wait_ns, wait_n, fence_ns, fence_n, wait_max_ns;
};

// Scalars the host must see (Krylov convergence tests, step-size control): the producing kernel stores them straight
// into pinned, device-mapped host memory and then a sequence number (system-scope release); the host spins on the
// sequence number.  No copy engine, no stream synchronisation: the round trip is ~2 us instead of ~30.
struct HostPoll {
    double v[64];
    unsigned long long seq;
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
//                 done -- nothing on the neighbours still reads the ghost planes this kernel is about to overwrite.
//                 The one pattern this cannot serve is a kernel that reads the ghosts of the very vector it pushes;
//                 no kernel does (stencil operand and result are distinct buffers, in-place updates are element-wise
//                 and read no ghosts).  tests/test_exchange_protocol_model.py checks every interleaving of the
//                 V-cycle + CG kernel sequence on an abstract 3-rank model;
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
    int opts;                                   // experiments (p4b_tune "port_opts"): 1 = keep the natural chunk order
                                                // and direction, 2 = signal at the end of the boundary CTAs' work
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long port_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// release-pattern fence at system scope (lighter than __threadfence_system(), which is fence.sc.sys)
__device__ __forceinline__ void port_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void port_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long port_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// gpu-scope release fence of a boundary CTA before it counts itself in (cheap: the stores only have to reach L2 order)
__device__ __forceinline__ void port_fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void port_st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int PORT_OPT_STATS = 4;     // HaloPort::opts bit: collect wait / fence times (globaltimer + atomics; off by default)
// Only "boundary" CTAs -- those that read a ghost plane or store into a neighbour's -- take part in the protocol;
// `boundary` must be uniform over the CTA, and the same CTAs must call port_signal with it.
// All their threads, before the first ghost read / peer store:
// (async_proxy: the ghost planes are read by bulk / TMA copies -- the plane-marching kernels -- so the generic-proxy
// acquire must be carried over to the async proxy; kernels that read them with ordinary loads skip that fence)
__device__ __forceinline__ void port_wait(const HaloPort &hp, bool boundary, bool async_proxy = false) {
    if (!hp.sync || !boundary) return;
    if (threadIdx.x == 0) {
        const bool stats = (hp.opts & PORT_OPT_STATS) != 0;
        const unsigned long long t0 = stats ? port_now_ns() : 0ull;
        const unsigned long long e = *(volatile unsigned long long *)&hp.sync->halo_epoch;
        if (hp.opts & (8 | 16)) {
            // spin with relaxed loads, then ONE acquire fence (an acquire load per spin invalidates L1 every time)
            if (hp.flag_lo) while (*(volatile const unsigned long long *)&hp.my_flags[0] < e) { }
            if (hp.flag_hi) while (*(volatile const unsigned long long *)&hp.my_flags[1] < e) { }
            if (!(hp.opts & 16)) port_fence_sys();         // (16: attribution experiment only -- no acquire at all)
        } else {
            if (hp.flag_lo) while (port_ld_acquire_sys(&hp.my_flags[0]) < e) { }
            if (hp.flag_hi) while (port_ld_acquire_sys(&hp.my_flags[1]) < e) { }
        }
        if (async_proxy) asm volatile("fence.proxy.async;" ::: "memory");
        if (stats) {
            const unsigned long long dt = port_now_ns() - t0;
            atomicAdd(&hp.sync->wait_ns, dt);
            atomicAdd(&hp.sync->wait_n, 1ull);
            atomicMax(&hp.sync->wait_max_ns, dt);
        }
    }
    __syncthreads();
}
// idx = local index (first owned element = 0) of the element just computed
__device__ __forceinline__ void port_store(const HaloPort &hp, long long idx, double v) {
    if (hp.lo_dst && idx < hp.plane) hp.lo_dst[idx] = v;
    if (hp.hi_dst && idx >= hp.hi_start) hp.hi_dst[idx - hp.hi_start] = v;
}
// Publication of a kernel's pushes: ONE system-scope fence per kernel.  Every boundary CTA orders its peer stores with
// a gpu-scope release fence and counts itself in (pdone); the CTA that completes the count has thereby observed all of
// them (gpu-scope acquire through the counter), so its single fence.acq_rel.sys -- cumulative in the PTX memory model
// -- orders every boundary CTA's stores before the flags it then writes.  (Round 1 had every boundary CTA issue its
// own system-scope fence, 4-5 us each and hundreds of them per kernel: measured with local stand-in targets, that --
// not NVLink -- was a quarter of the 8-GPU step, profiles/r02_exchange.md.)
__device__ __forceinline__ void port_publish(const HaloPort &hp, unsigned int nboundary, bool stored) {
    if (stored && !(hp.opts & 32)) port_fence_gpu();       // (32: attribution experiment only)
    if (atomicAdd(&hp.sync->pdone, 1u) == nboundary - 1u) {
        const bool stats = (hp.opts & PORT_OPT_STATS) != 0;
        const unsigned long long t0 = stats ? port_now_ns() : 0ull;
        port_fence_sys();
        const unsigned long long e = hp.sync->halo_epoch + 1ull;
        if (hp.flag_lo) port_st_relaxed_sys(hp.flag_lo, e);
        if (hp.flag_hi) port_st_relaxed_sys(hp.flag_hi, e);
        hp.sync->halo_epoch = e;
        hp.sync->pdone = 0u;
        __threadfence();
        if (stats) {
            atomicAdd(&hp.sync->fence_ns, port_now_ns() - t0);
            atomicAdd(&hp.sync->fence_n, 1ull);
        }
    }
}
// Variant with a dedicated signalling CTA (the plane-marching kernels, where the publishing CTA would otherwise stall
// its own march for the 4-5 us of the system-scope fence): boundary CTAs only count themselves in (port_arrive, one
// elected thread); one extra CTA of the grid waits for the count and publishes (port_signaller, one thread).
__device__ __forceinline__ void port_arrive(const HaloPort &hp) {
    if (!hp.sync || !hp.push) return;
    port_fence_gpu();
    atomicAdd(&hp.sync->pdone, 1u);
}
__device__ __forceinline__ void port_signaller(const HaloPort &hp, unsigned int nboundary) {
    if (!hp.sync || !hp.push) return;
    unsigned int seen;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&hp.sync->pdone) : "memory");
    } while (seen < nboundary);
    port_fence_sys();
    const unsigned long long e = hp.sync->halo_epoch + 1ull;
    if (hp.flag_lo) port_st_relaxed_sys(hp.flag_lo, e);
    if (hp.flag_hi) port_st_relaxed_sys(hp.flag_hi, e);
    hp.sync->halo_epoch = e;
    hp.sync->pdone = 0u;
    __threadfence();
}
// All threads of the boundary CTAs, after their last peer store; nboundary = number of boundary CTAs of the grid.
// (stored = false: this CTA is only counted, it has no peer stores of its own to publish)
__device__ __forceinline__ void port_signal(const HaloPort &hp, bool boundary, unsigned int nboundary,
                                            bool stored = true) {
    if (!hp.sync || !hp.push || !boundary) return;
    __syncthreads();
    if (threadIdx.x == 0) port_publish(hp, nboundary, stored);
}
// The same for a caller that has just passed a __syncthreads() after its last peer store: `elected` is true in
// exactly one thread of the CTA.
__device__ __forceinline__ void port_signal_nosync(const HaloPort &hp, unsigned int nboundary, bool elected) {
    if (!hp.sync || !hp.push || !elected) return;
    port_publish(hp, nboundary, true);
}
// Launch-order remap of a block index that runs over planes / chunks of the slab: with an upper neighbour the block
// that holds the LAST plane is scheduled second, so both boundary planes are produced (and signalled) first.
__device__ __forceinline__ unsigned int port_remap(const HaloPort &hp, unsigned int b, unsigned int nb) {
    if (!hp.sync || !hp.flag_hi || nb < 2u || (hp.opts & 1)) return b;
    return b == 0u ? 0u : (b == 1u ? nb - 1u : b - 1u);
}
// Grid-stride elementwise kernels: the slab is cut into segments of W elements and CTA b takes the logical segments
// b, b + G, b + 2G, ...  Logical segment g is the physical segment seg_phys(g): the segments of the last plane come
// first, then those of the first plane, then the interior -- so the boundary planes are written in the first pass
// and can be signalled while the rest of the vector is still being streamed.
struct SegRot {
    long long S, H, NB;       // segments, segments of the last plane rotated to the front, boundary segments
    __device__ __forceinline__ long long phys(long long g) const { return g < H ? S - H + g : g - H; }
};
__device__ __forceinline__ SegRot seg_rot(const HaloPort &hp, long long n, long long W) {
    SegRot r;
    r.S = (n + W - 1) / W;
    const bool on = hp.sync != nullptr && hp.push;
    r.H = (on && hp.flag_hi) ? r.S - hp.hi_start / W : 0;
    const long long A = (on && hp.flag_lo) ? min((hp.plane + W - 1) / W, r.S) : 0;
    r.NB = min(r.S, r.H + A);
    return r;
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

// one-thread kernel that only waits for the neighbours' flags (for consumers that do not carry a HaloPort themselves)
int launch_port_wait(cudaStream_t st, const HaloPort &port);
int launch_halo_push(cudaStream_t st, const double *lo_src, double *lo_dst, const double *hi_src, double *hi_dst,
                     long long plane, unsigned long long *flag_prev, unsigned long long *flag_next,
                     const unsigned long long *my_flags, LocalSync *sync);
// hp != nullptr: the result is also published to the host (HostPoll) with sequence number seq
// (extra != nullptr: one more device scalar, already reduced, rides along into hp->v[nv] under the same sequence number)
int launch_allreduce(cudaStream_t st, double *vals, int nv, int op_max, const PeerTable &peers, LocalSync *sync,
                     HostPoll *hp = nullptr, unsigned long long seq = 0, const double *extra = nullptr);
// vals[0..nv) -> hp->v [, extra[0] -> hp->v[nv]], then hp->seq = seq (one tiny kernel; nv < 64)
int launch_publish(cudaStream_t st, const double *vals, int nv, HostPoll *hp, unsigned long long seq,
                   const double *extra = nullptr);
int launch_barrier(cudaStream_t st, int all, const PeerTable &peers, LocalSync *sync);
int launch_gather_push(cudaStream_t st, const double *src, long long n, long long off_doubles, const GatherTable &dst,
                       int rank, int nranks);

}  // namespace p4b
