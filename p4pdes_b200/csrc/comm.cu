// comm.cu -- peer-memory communication over NVLink/NVSwitch for the slab-decomposed solver.
//
// One process per GPU.  Every rank maps its neighbours' level arenas and every rank's small "mailbox"
// through CUDA IPC, after which the [PETSc] DMGlobalToLocal ghost exchange, the scalar MPI_Allreduce of the
// Krylov dot products and the PCREDUNDANT-style gather are ordinary kernels that store straight into peer
// memory and synchronise with release/acquire flags (system scope):
//
//   halo_push_kernel   my first/last owned plane  -> the neighbours' ghost planes, then flag + wait
//   allreduce_kernel   my partial sums -> slot [rank] of every peer's mailbox; every rank then adds the slots
//                      in rank order (bit-identical result on all ranks, independent of arrival order)
//   gather_push_kernel my owned planes of a replicated level's vector -> the same planes on every peer
//
// No NCCL kernel is launched on this path: an exchange costs one small kernel (~launch + NVLink latency)
// instead of a grouped ncclSend/ncclRecv, and everything is CUDA-graph capturable.
//
// Hazards.  Exchanges happen in the same program order on all ranks and carry one monotone epoch.  A push
// of exchange n can only race with a neighbour's kernel that still reads the same ghost plane from exchange
// n-1 (the neighbour is known to have passed n-1, so anything older is finished); mg.cu therefore inserts a
// flag-only neighbour barrier when the same buffer is exchanged twice in a row.  The gather targets the
// same buffer every cycle, so it is bracketed by an all-rank barrier.
#include "comm.h"

namespace p4b {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// copy n doubles src -> dst with the widest aligned access both pointers allow
__device__ __forceinline__ void copy_span(double *__restrict__ dst, const double *__restrict__ src, long long n,
                                          int tid, int nthreads) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const long long n2 = n >> 1;
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (long long i = tid; i < n2; i += nthreads) d2[i] = s2[i];
        if ((n & 1) && tid == 0) dst[n - 1] = src[n - 1];
    } else {
        for (long long i = tid; i < n; i += nthreads) dst[i] = src[i];
    }
}

// all CTAs copy, the last one to finish signals the neighbours and waits for their signal
__global__ void __launch_bounds__(512) halo_push_kernel(const double *lo_src, double *lo_dst, const double *hi_src,
                                                         double *hi_dst, long long plane, unsigned long long *flag_prev,
                                                         unsigned long long *flag_next, const unsigned long long *my_flags,
                                                         LocalSync *sync) {
    const int nb = gridDim.x, half = nb / 2;
    // every earlier exchange has landed here, i.e. the neighbours have finished the kernels that read the ghost
    // planes about to be overwritten (same rule as the fused HaloPort exchanges, comm.h)
    if (threadIdx.x == 0) {
        const unsigned long long e0 = *(volatile unsigned long long *)&sync->halo_epoch;
        if (flag_prev) while (ld_acquire_sys(&my_flags[0]) < e0) { }
        if (flag_next) while (ld_acquire_sys(&my_flags[1]) < e0) { }
    }
    __syncthreads();
    // first half of the CTAs serves the lower neighbour, second half the upper one
    if (lo_dst && (int)blockIdx.x < half)
        copy_span(lo_dst, lo_src, plane, blockIdx.x * blockDim.x + threadIdx.x, half * blockDim.x);
    if (hi_dst && (int)blockIdx.x >= half)
        copy_span(hi_dst, hi_src, plane, (blockIdx.x - half) * blockDim.x + threadIdx.x, (nb - half) * blockDim.x);
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(&sync->done, 1u) == (unsigned)nb - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        const unsigned long long e = sync->halo_epoch + 1;
        if (flag_prev) st_release_sys(flag_prev, e);
        if (flag_next) st_release_sys(flag_next, e);
        if (flag_prev) while (ld_acquire_sys(&my_flags[0]) < e) { }
        if (flag_next) while (ld_acquire_sys(&my_flags[1]) < e) { }
        sync->halo_epoch = e;
        sync->done = 0u;
        __threadfence();
    }
}

// sum nv (<= 3) doubles over all ranks; result overwrites vals on every rank
__global__ void __launch_bounds__(64) allreduce_kernel(double *vals, int nv, int op_max, PeerTable peers, LocalSync *sync,
                                                       HostPoll *hp, unsigned long long seq, const double *extra) {
    const int r = threadIdx.x;
    const unsigned long long e = sync->red_epoch + 1;
    const int par = (int)(e & 1ull);
    if (r < peers.nranks) {
        RedSlot *slot = &peers.mbox[r]->red[par][peers.rank];      // my slot in rank r's mailbox
        for (int q = 0; q < nv; q++) slot->v[q] = vals[q];
        __threadfence_system();
        st_release_sys(&slot->epoch, e);
    }
    __syncthreads();
    if (r < peers.nranks) {
        const RedSlot *mine = &peers.mbox[peers.rank]->red[par][r];
        while (ld_acquire_sys(&mine->epoch) < e) { }
    }
    __syncthreads();
    if (r < nv) {
        const Mailbox *mb = peers.mbox[peers.rank];
        double s = op_max ? -1.0e308 : 0.0;
        for (int q = 0; q < peers.nranks; q++) {
            const double v = ((volatile const double *)mb->red[par][q].v)[r];
            s = op_max ? fmax(s, v) : s + v;
        }
        vals[r] = s;
        if (hp) ((volatile double *)hp->v)[r] = s;
    }
    __syncthreads();
    if (r == 0) {
        sync->red_epoch = e;
        if (hp) {
            if (extra) ((volatile double *)hp->v)[nv] = extra[0];
            __threadfence_system();
            st_release_sys(&hp->seq, seq);
        }
    }
}

__global__ void __launch_bounds__(64) publish_kernel(const double *vals, int nv, HostPoll *hp, unsigned long long seq,
                                                     const double *extra) {
    if ((int)threadIdx.x < nv) ((volatile double *)hp->v)[threadIdx.x] = vals[threadIdx.x];
    if (extra && (int)threadIdx.x == nv) ((volatile double *)hp->v)[nv] = extra[0];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&hp->seq, seq);
    }
}

// flag-only barrier: among the two slab neighbours (all = 0) or all ranks (all = 1)
__global__ void __launch_bounds__(64) barrier_kernel(int all, PeerTable peers, LocalSync *sync) {
    const int r = threadIdx.x;
    const unsigned long long e = sync->bar_epoch + 1;
    const bool peer = r < peers.nranks && r != peers.rank && (all || r == peers.rank - 1 || r == peers.rank + 1);
    if (peer) {
        __threadfence_system();
        st_release_sys(&peers.mbox[r]->bar_flag[peers.rank], e);
    }
    if (peer) while (ld_acquire_sys(&peers.mbox[peers.rank]->bar_flag[r]) < e) { }
    __syncthreads();
    if (r == 0) sync->bar_epoch = e;
}

// replicated level: every rank stores its owned planes into the same place of every peer's copy
__global__ void __launch_bounds__(512) gather_push_kernel(const double *src, long long n, long long off_doubles,
                                                           GatherTable dst, int rank, int nranks) {
    // blockIdx.y = destination rank
    const int d = blockIdx.y;
    if (d == rank || d >= nranks || n <= 0) return;
    copy_span(dst.base[d] + off_doubles, src, n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    __threadfence_system();
}

__global__ void __launch_bounds__(32) port_wait_kernel(const HaloPort port) { port_wait(port, true, true); }

int launch_port_wait(cudaStream_t st, const HaloPort &port) {
    if (!port.sync) return 0;
    port_wait_kernel<<<1, 32, 0, st>>>(port);
    P4B_LAUNCH_CHECK();
    return 0;
}

int launch_halo_push(cudaStream_t st, const double *lo_src, double *lo_dst, const double *hi_src, double *hi_dst,
                     long long plane, unsigned long long *flag_prev, unsigned long long *flag_next,
                     const unsigned long long *my_flags, LocalSync *sync) {
    int nb = (int)((plane + 4095) / 4096) * 2;
    if (nb < 2) nb = 2;
    if (nb > 64) nb = 64;
    halo_push_kernel<<<nb, 512, 0, st>>>(lo_src, lo_dst, hi_src, hi_dst, plane, flag_prev, flag_next, my_flags, sync);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_allreduce(cudaStream_t st, double *vals, int nv, int op_max, const PeerTable &peers, LocalSync *sync,
                     HostPoll *hp, unsigned long long seq, const double *extra) {
    if (nv < 1 || nv > 3) return fail(62, "peer allreduce handles 1..3 values");
    allreduce_kernel<<<1, 64, 0, st>>>(vals, nv, op_max, peers, sync, hp, seq, extra);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_publish(cudaStream_t st, const double *vals, int nv, HostPoll *hp, unsigned long long seq, const double *extra) {
    if (nv < 1 || nv > 64 || (extra && nv > 63)) return fail(62, "publish handles 1..64 values");
    publish_kernel<<<1, 64, 0, st>>>(vals, nv, hp, seq, extra);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_barrier(cudaStream_t st, int all, const PeerTable &peers, LocalSync *sync) {
    barrier_kernel<<<1, 64, 0, st>>>(all, peers, sync);
    P4B_LAUNCH_CHECK();
    return 0;
}
int launch_gather_push(cudaStream_t st, const double *src, long long n, long long off_doubles, const GatherTable &dst,
                       int rank, int nranks) {
    if (n <= 0) return 0;
    int nbx = (int)((n + 8191) / 8192);
    if (nbx < 1) nbx = 1;
    if (nbx > 16) nbx = 16;
    dim3 grid(nbx, nranks);
    gather_push_kernel<<<grid, 512, 0, st>>>(src, n, off_doubles, dst, rank, nranks);
    P4B_LAUNCH_CHECK();
    return 0;
}

}  // namespace p4b
