// stencil_fast.cu -- TMA-staged plane-marching version of the stencil family for big grids (sm_100a).
//
// Work decomposition: the (i,j) plane is cut into linear "bands" of Q = NT*P consecutive nodes
// (natural ordering, so a band is a contiguous piece of memory in every z-plane, whatever nx is:
// no tail lanes for odd nx = 513), and the z range into chunks of KC planes.  One CTA marches one
// band through one chunk.
//
// Data movement per z-plane and CTA:
//   * the stencil operand u: ONE bulk async copy (cp.async.bulk -> SASS UBLKCP, the 1-D TMA path)
//     of the band plus a halo of H = nx nodes on both sides (the rows above/below) into a ring of
//     NS shared-memory stages, completion signalled on an mbarrier (expect_tx/complete_tx).
//     A thread's x/y neighbours are read from the stage; its z neighbours are the centre values of
//     the previous/next plane, kept in registers (each node's value is read from shared memory once).
//   * the streamed operands (b, pm1) are prefetched one plane ahead into registers with plain
//     coalesced loads; the result is stored straight from registers.
// DRAM traffic is therefore the algorithmic minimum (each vector is read/written once) plus
// 2H/Q halo re-reads that hit in L2 because neighbouring bands run concurrently.
//
// Alignment: cp.async.bulk needs 16-byte aligned addresses and sizes, but nx*ny is odd for the
// 2^k+1 grids, so a plane's band may start on an odd element.  The copy then starts one element
// early (the stage keeps a per-plane parity `adj`) and an odd trailing element is moved by the
// producer thread with an ordinary load/store before it arrives on the barrier.
#include "kernels.h"

namespace p4b {

struct MarchCfg {
    int KC;        // planes per chunk
    int NS;        // stages in the ring
    int H;         // halo nodes on each side of a band in the stage
    int stage_doubles;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

enum { MK_W = 1, MK_E = 2, MK_S = 4, MK_N = 8, MK_BD = 16, MK_VALID = 32 };

template <int MODE, int P, int NT>
__global__ void __launch_bounds__(NT) stencil_march_kernel(const LevelDesc L, const StencilOp op, const MarchCfg cfg,
                                                           double *partials, unsigned int *ticket) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);              // NS barriers (128 bytes reserved)
    double *stages = reinterpret_cast<double *>(smem_raw + 128);
    constexpr int Q = NT * P;
    const int tid = threadIdx.x;
    const int plane = L.nx * L.ny;
    const int q0 = blockIdx.x * Q;
    const int k0 = blockIdx.y * cfg.KC;
    const int k1 = min(k0 + cfg.KC, L.zm);                                 // local planes [k0, k1)
    const int NS = cfg.NS, H = cfg.H, SD = cfg.stage_doubles;
    const int lo = max(q0 - H, 0), hi = min(q0 + Q + H, plane);            // staged linear range of the plane
    const int cnt = hi - lo;
    const double *__restrict__ u = op.u;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // plane t of this chunk is local plane kl = k0 - 1 + t, t = 0 .. T-1 (one extra plane on each side)
    const int T = (k1 - k0) + 2;
    auto plane_valid = [&](int t) {
        const int kg = L.zs + k0 - 1 + t;
        return kg >= 0 && kg <= L.nz - 1;
    };
    auto plane_adj = [&](int t) {   // parity of the first staged element's offset from u
        const long long e0 = (long long)(k0 - 1 + t) * plane + lo;
        return (int)(e0 & 1LL);
    };
    auto issue = [&](int t) {       // producer thread only
        if (t >= T) return;
        const int s = t % NS;
        if (!plane_valid(t)) {          // nothing to load: complete the phase so slot parities stay in step
            mbar_arrive(&bars[s]);
            return;
        }
        double *st = stages + (size_t)s * SD;
        const long long e0 = (long long)(k0 - 1 + t) * plane + lo;
        const int adj = (int)(e0 & 1LL);
        const int total = cnt + adj;              // elements from the aligned start
        const int even = total & ~1;
        const double *src = u + (e0 - adj);
        if (total & 1) st[2 + total - 1] = src[total - 1];   // odd tail by the generic proxy
        fence_proxy_async();
        mbar_arrive_expect_tx(&bars[s], (uint32_t)even * 8u);
        bulk_g2s(st + 2, src, (uint32_t)even * 8u, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS; t++) issue(t);

    // per-thread nodes: q = q0 + tid + p*NT
    int mask[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        const int q = q0 + tid + p * NT;
        int mk = 0;
        if (q < plane) {
            const int j = q / L.nx, i = q - j * L.nx;
            mk = MK_VALID;
            if (i == 0 || i == L.nx - 1 || (L.ay && (j == 0 || j == L.ny - 1))) mk |= MK_BD;
            if (i - 1 > 0) mk |= MK_W;
            if (i + 1 < L.nx - 1) mk |= MK_E;
            if (L.ay && j - 1 > 0) mk |= MK_S;
            if (L.ay && j + 1 < L.ny - 1) mk |= MK_N;
        }
        mask[p] = mk;
    }
    const int sbase = 2 + (q0 - lo) + tid;      // stage index of node p=0 before the parity shift

    double prev[P], cur[P], nxt[P], bq[P], pq[P];
    double dv[1] = {0.0};
#pragma unroll
    for (int p = 0; p < P; p++) { prev[p] = 0.0; cur[p] = 0.0; nxt[p] = 0.0; bq[p] = 0.0; pq[p] = 0.0; }

    // centre values of plane t = 0
    if (plane_valid(0)) {
        mbar_wait(&bars[0], 0);
        const double *st = stages + plane_adj(0);
#pragma unroll
        for (int p = 0; p < P; p++)
            if (mask[p] & MK_VALID) cur[p] = st[sbase + p * NT];
    }

    for (int t = 0; t + 1 < T; t++) {
        const int kl = k0 - 1 + t;            // plane computed in this iteration (when t >= 1)
        const int kg = L.zs + kl;
        // prefetch the streamed operands of the next plane to compute (t+1) into registers
        double bn[P], pn[P];
        if (MODE == ST_LIN || MODE == ST_LIN_PM1) {
            const bool more = (t + 1 < T - 1);
            const long long nb = (long long)(kl + 1) * plane + q0 + tid;
#pragma unroll
            for (int p = 0; p < P; p++) {
                bn[p] = (more && (mask[p] & MK_VALID)) ? op.b[nb + p * NT] : 0.0;
                if (MODE == ST_LIN_PM1) pn[p] = (more && (mask[p] & MK_VALID)) ? op.pm1[nb + p * NT] : 0.0;
            }
        }
        // centre values of plane t+1
        {
            const int tn = t + 1, sn = tn % NS;
            if (plane_valid(tn)) {
                mbar_wait(&bars[sn], (uint32_t)((tn / NS) & 1));
                const double *st = stages + (size_t)sn * SD + plane_adj(tn);
#pragma unroll
                for (int p = 0; p < P; p++)
                    if (mask[p] & MK_VALID) nxt[p] = st[sbase + p * NT];
            } else {
#pragma unroll
                for (int p = 0; p < P; p++) nxt[p] = 0.0;
            }
        }
        if (t >= 1) {
            const double *st = stages + (size_t)(t % NS) * SD + plane_adj(t);
            const bool kbd = (kg == 0 || kg == L.nz - 1);
            const bool dn_ok = (kg - 1 > 0), up_ok = (kg + 1 < L.nz - 1);
            const long long nb = (long long)kl * plane + q0 + tid;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const int mk = mask[p];
                if (!(mk & MK_VALID)) continue;
                const int si = sbase + p * NT;
                const double uc = cur[p];
                double Au = L.diag * uc;
                if (!kbd && !(mk & MK_BD)) {
                    const double uw = (mk & MK_W) ? st[si - 1] : 0.0;
                    const double ue = (mk & MK_E) ? st[si + 1] : 0.0;
                    Au -= L.cx * (uw + ue);
                    if (L.ay) {
                        const double us = (mk & MK_S) ? st[si - L.nx] : 0.0;
                        const double un = (mk & MK_N) ? st[si + L.nx] : 0.0;
                        Au -= L.cy * (us + un);
                    }
                    const double ud = dn_ok ? prev[p] : 0.0;
                    const double uu = up_ok ? nxt[p] : 0.0;
                    Au -= L.cz * (uu + ud);
                }
                double o;
                if (MODE == ST_APPLY || MODE == ST_APPLY_DOT) {
                    o = Au;
                    if (MODE == ST_APPLY_DOT) dv[0] += uc * Au;
                } else if (MODE == ST_LIN_BU) {
                    o = op.cb * uc + op.cg * (uc - Au);
                } else {
                    o = op.cb * uc + op.cg * (bq[p] - Au);
                    if (MODE == ST_LIN_PM1) o += op.ca * pq[p];
                }
                op.out[nb + p * NT] = o;
            }
        }
        __syncthreads();                       // everyone is done with stage t%NS (and with plane 0 at t = 0)
        if (tid == 0) issue(t + NS);
#pragma unroll
        for (int p = 0; p < P; p++) {
            prev[p] = cur[p];
            cur[p] = nxt[p];
            if (MODE == ST_LIN || MODE == ST_LIN_PM1) {
                bq[p] = bn[p];
                if (MODE == ST_LIN_PM1) pq[p] = pn[p];
            }
        }
    }
    if (MODE == ST_APPLY_DOT) {
        // one partial per CTA, fixed order: block id = blockIdx.y * gridDim.x + blockIdx.x
        __shared__ double red[NT / 32];
        __shared__ bool is_last;
        const int lane = tid & 31, wid = tid >> 5;
        double s = warp_sum(dv[0]);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        const unsigned int nblk = gridDim.x * gridDim.y;
        if (wid == 0) {
            double t2 = (lane < NT / 32) ? red[lane] : 0.0;
            t2 = warp_sum(t2);
            if (lane == 0) {
                partials[blockIdx.y * gridDim.x + blockIdx.x] = t2;
                __threadfence();
                is_last = (atomicAdd(ticket, 1u) == nblk - 1);
            }
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double a = 0.0;
            for (unsigned int i = tid; i < nblk; i += NT) a += ((volatile double *)partials)[i];
            a = warp_sum(a);
            __syncthreads();
            if (lane == 0) red[wid] = a;
            __syncthreads();
            if (wid == 0) {
                double t3 = (lane < NT / 32) ? red[lane] : 0.0;
                t3 = warp_sum(t3);
                if (lane == 0) { op.dot_out[0] = t3; *ticket = 0u; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: eligibility, configuration, launch
// ---------------------------------------------------------------------------------------------
struct MarchTune { int P, NT, NS, enabled, min_plane; };

static MarchTune &tune() {
    static MarchTune t = [] {
        MarchTune x = {4, 512, 4, 1, 16384};
        if (const char *e = getenv("P4B_MARCH")) {       // "P,NT,NS" or "0" to disable (tuning / A-B runs)
            int a = 0, b = 0, c = 0;
            const int n = sscanf(e, "%d,%d,%d", &a, &b, &c);
            if (n == 1 && a == 0) x.enabled = 0;
            if (n == 3) { x.P = a; x.NT = b; x.NS = c; }
        }
        return x;
    }();
    return t;
}

int tune_march(const char *key, long v) {
    MarchTune &t = tune();
    const std::string k(key);
    if (k == "march_enabled") t.enabled = (int)v;
    else if (k == "march_min_plane") t.min_plane = (int)v;
    else if (k == "march_P") t.P = (int)v;
    else if (k == "march_NT") t.NT = (int)v;
    else if (k == "march_NS") t.NS = (int)v;
    else return 1;
    return 0;
}

bool stencil_fast_eligible(const LevelDesc &L) {
    const MarchTune &t = tune();
    return t.enabled && L.ax && L.az && (long long)L.nx * L.ny >= t.min_plane && L.zm >= 8 &&
           (long long)L.nx * L.ny * (L.zm + 2) < (1LL << 31);
}

template <int MODE, int P, int NT>
static int launch_march(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red, int NS) {
    static int sm_count = 0, max_smem = 0;
    if (!sm_count) {
        int dev = 0;
        P4B_CUDA(cudaGetDevice(&dev));
        P4B_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        P4B_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    constexpr int Q = NT * P;
    const int plane = L.nx * L.ny;
    MarchCfg cfg;
    cfg.H = L.ay ? L.nx : 2;
    cfg.stage_doubles = (Q + 2 * cfg.H + 4 + 1) & ~1;
    const size_t stage_bytes = (size_t)cfg.stage_doubles * 8;
    while (NS > 3 && 128 + NS * stage_bytes + 1024 > (size_t)max_smem) NS--;
    cfg.NS = NS;
    const size_t smem = 128 + NS * stage_bytes;
    if (smem + 1024 > (size_t)max_smem) return fail(62, "plane-marching stage does not fit shared memory");
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        P4B_CUDA(cudaFuncSetAttribute(stencil_march_kernel<MODE, P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_smem = smem;
    }
    int occ = 1;
    P4B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stencil_march_kernel<MODE, P, NT>, NT, smem));
    if (occ < 1) occ = 1;
    const int bands = (plane + Q - 1) / Q;
    const long long slots = (long long)sm_count * occ;
    // number of z chunks: minimise waves * (KC + 2)
    int best_nc = 1;
    double best_cost = 1e300;
    for (int nc = 1; nc <= L.zm / 4 && nc <= 256; nc++) {
        const int KC = (L.zm + nc - 1) / nc;
        const int nce = (L.zm + KC - 1) / KC;
        const long long ctas = (long long)bands * nce;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (KC + 2 + 1.5);   // +1.5: pipeline fill per CTA
        if (cost < best_cost - 1e-9) { best_cost = cost; best_nc = nce; }
    }
    cfg.KC = (L.zm + best_nc - 1) / best_nc;
    const int nchunks = (L.zm + cfg.KC - 1) / cfg.KC;
    if (MODE == ST_APPLY_DOT && bands * nchunks > red.max_blocks) return fail(63, "reducer scratch too small");
    dim3 grid(bands, nchunks);
    stencil_march_kernel<MODE, P, NT><<<grid, NT, smem, st>>>(L, op, cfg, red.partials, red.ticket);
    P4B_LAUNCH_CHECK();
    return 0;
}

template <int P, int NT>
static int launch_march_mode(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red, int NS) {
    switch (op.mode) {
        case ST_APPLY: return launch_march<ST_APPLY, P, NT>(st, L, op, red, NS);
        case ST_APPLY_DOT: return launch_march<ST_APPLY_DOT, P, NT>(st, L, op, red, NS);
        case ST_LIN: return launch_march<ST_LIN, P, NT>(st, L, op, red, NS);
        case ST_LIN_PM1: return launch_march<ST_LIN_PM1, P, NT>(st, L, op, red, NS);
        case ST_LIN_BU: return launch_march<ST_LIN_BU, P, NT>(st, L, op, red, NS);
    }
    return fail(62, "unknown stencil mode %d", op.mode);
}

int launch_stencil_generic(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red);

int launch_stencil_fast(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red) {
    if (((uintptr_t)op.u & 15) != 0) return launch_stencil_generic(st, L, op, red);   // bulk copies need 16-byte alignment
    const MarchTune &t = tune();
    if (t.P == 4 && t.NT == 512) return launch_march_mode<4, 512>(st, L, op, red, t.NS);
    if (t.P == 8 && t.NT == 256) return launch_march_mode<8, 256>(st, L, op, red, t.NS);
    if (t.P == 8 && t.NT == 512) return launch_march_mode<8, 512>(st, L, op, red, t.NS);
    if (t.P == 4 && t.NT == 256) return launch_march_mode<4, 256>(st, L, op, red, t.NS);
    if (t.P == 2 && t.NT == 512) return launch_march_mode<2, 512>(st, L, op, red, t.NS);
    return fail(62, "P4B_MARCH=%d,%d is not instantiated", t.P, t.NT);
}

}  // namespace p4b
