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
// Multi-GPU slabs (template parameter MG, comm.h HaloPort): the CTAs of the first chunk read the lower ghost plane,
// those of the last chunk the upper one, after waiting for the neighbours' flags; a kernel that writes a ghosted
// vector stores its first / last owned plane into the neighbours' ghost planes as well.  So that BOTH boundary
// planes are the first planes computed, the last chunk is scheduled second and marched downwards (the z stencil
// sum is formed from two rounded products, so the direction does not change the bits), and the push is published
// (system-scope fence + flags) right after that first plane, while the rest of the chunk is still being marched.
// MG = false compiles all of this out.
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

enum { MK_BD = 1, MK_VALID = 2 };

// Inner-loop design notes (the first version of this kernel was instruction-issue bound at ~200
// instructions per node, see profiles/r01_march_v1.md):
//   * Dirichlet masking is applied to the DATA, not the operator: once a plane has landed, the thread
//     that owns a boundary node reads its true value into a register and then zeroes it in the stage
//     (halo boundary nodes are zeroed by a few designated threads).  Interior rows then use the plain
//     7-point formula with no per-neighbour predicates; boundary rows are a final select.
//   * slot index, mbarrier parity bits, stage parity shift and all global pointers advance
//     incrementally (no division or 64-bit multiply per plane), invalid tail nodes are clamped and
//     only their store is predicated off (no divergent control flow).
//   * MG (multi-GPU) = false compiles the slab-exchange code (HaloPort waits / pushes / signals, downward march of
//     the last chunk) out: the single-GPU kernel carries none of its registers or instructions.
template <int MODE, int P, int NT, bool MG>
__global__ void __launch_bounds__(NT, (NT <= 256 && P <= 8) ? 2 : 1) stencil_march_kernel(const LevelDesc L, const StencilOp op, const MarchCfg cfg,
                                                           double *partials, unsigned int *ticket) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);              // NS barriers (128 bytes reserved)
    double *stages = reinterpret_cast<double *>(smem_raw + 128);
    constexpr int Q = NT * P;
    constexpr int HZ = 3;                                                  // halo nodes a thread may have to zero
    const int tid = threadIdx.x;
    // MG grids carry one extra column of CTAs: (nbands, 0) is the signalling CTA (comm.h port_signaller), the rest of
    // the column exits at once
    const unsigned int nbands = MG ? gridDim.x - 1u : gridDim.x;
    const unsigned int n_port_ctas =
        nbands * (gridDim.y == 1 ? 1u : (op.port.flag_lo != nullptr) + (op.port.flag_hi != nullptr));
    if (MG && blockIdx.x == nbands) {
        if (blockIdx.y == 0 && tid == 0 && n_port_ctas > 0u) port_signaller(op.port, n_port_ctas);
        return;
    }
    const int nx = L.nx;
    const int plane = L.nx * L.ny;
    const int q0 = blockIdx.x * Q;
    // With an upper slab neighbour the LAST chunk is scheduled second and marched downwards, so that both boundary
    // planes of the slab are the first planes computed: they are pushed to the neighbours and signalled while the
    // rest of the kernel still runs (comm.h).  Without one, chunks are taken in order, all upwards.
    const bool hi_nb = MG && op.port.push && !(op.port.opts & 1) && op.port.flag_hi != nullptr && gridDim.y > 1;
    const int chunk = !hi_nb ? (int)blockIdx.y
                             : (blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? (int)gridDim.y - 1 : (int)blockIdx.y - 1));
    const bool rev = hi_nb && chunk == (int)gridDim.y - 1;
    const int k0 = chunk * cfg.KC;
    const int k1 = min(k0 + cfg.KC, L.zm);                                 // local planes [k0, k1)
    const int NS = cfg.NS, H = cfg.H, SD = cfg.stage_doubles;
    const int lo = max(q0 - H, 0), hi = min(q0 + Q + H, plane);            // staged linear range of the plane
    const int cnt = hi - lo;
    const int podd = plane & 1;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // CTAs of the first / last chunk read u's ghost plane and push out's boundary plane
    const bool port_cta = MG && op.port.sync != nullptr && ((chunk == 0 && op.port.flag_lo != nullptr) ||
                                                      (chunk == (int)gridDim.y - 1 && op.port.flag_hi != nullptr));
    if (MG) port_wait(op.port, port_cta, true);

    // plane t of this chunk, t = 0 .. T-1 (one extra plane on each side), is local plane k0 - 1 + t going up
    // and k1 - t going down
    const int T = (k1 - k0) + 2;
    const int dir = (MG && rev) ? -1 : 1;
    const int kl0 = rev ? k1 : k0 - 1;                                     // local index of plane t = 0
    const int kg0 = L.zs + kl0;                                            // its global index
    const long long dstep = (long long)dir * plane;                        // element step from plane t to t + 1
    const long long e00 = (long long)kl0 * plane + lo;                     // element offset of plane 0's staged range
    const int adj0 = (int)(e00 & 1LL);

    // producer state (thread 0): next plane to issue
    int it = 0, islot = 0;
    const double *isrc = op.u + e00;                                       // start of plane `it`'s staged range
    auto issue_next = [&]() {       // producer thread only; issues plane `it` into slot `islot`
        if (it < T) {
            const int kg = kg0 + dir * it;
            if (kg < 0 || kg > L.nz - 1) {
                mbar_arrive(&bars[islot]);      // nothing to load: complete the phase, slot parities stay in step
            } else {
                double *st = stages + (size_t)islot * SD;
                const int adj = (adj0 + it * podd) & 1;
                const int total = cnt + adj;              // elements from the aligned start
                const int even = total & ~1;
                const double *src = isrc - adj;
                if (total & 1) st[2 + total - 1] = src[total - 1];   // odd tail by the generic proxy
                fence_proxy_async();
                mbar_arrive_expect_tx(&bars[islot], (uint32_t)even * 8u);
                bulk_g2s(st + 2, src, (uint32_t)even * 8u, &bars[islot]);
            }
        }
        it++;
        isrc += dstep;
        if (++islot == NS) islot = 0;
    };
    if (tid == 0)
        for (int t = 0; t < NS; t++) issue_next();

    // per-thread nodes: q = q0 + tid + p*NT ; tail nodes beyond the plane are clamped onto its last node
    // Per-node indices.  Node p of a thread is `off(p)` elements after its first node, its stage index is sidx(p).
    // Normally both live in registers (si/gi).  The two-operand mode at 8 nodes/thread does not fit 128 registers
    // (two CTAs per SM) that way and recomputes them instead (LEAN): off(p) = min(p*NT, lim), lim = the thread's
    // largest in-plane offset -- which also clamps a tail node beyond the plane onto the plane's last node.  The
    // one-operand modes are instruction-issue bound and keep the arrays.  Boundary / validity flags: 2 bits per node.
    constexpr bool LEAN = (MODE == ST_LIN && P >= 8);
    const int lim = (plane - 1) - (q0 + tid);
    const int sb = 2 + (q0 - lo);
    const int sbase = sb + tid;
    int si[P], gi[P];
    unsigned int mbits = 0u;
    int mask[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        int q = q0 + tid + p * NT;
        unsigned int mk = MK_VALID;
        if (q >= plane) { q = plane - 1; mk = 0u; }
        const int j = q / nx, i = q - j * nx;
        if (i == 0 || i == nx - 1 || (L.ay && (j == 0 || j == L.ny - 1))) mk |= MK_BD;
        mbits |= mk << (2 * p);
        mask[p] = (int)mk;
        si[p] = sb + (q - q0);
        gi[p] = q - q0 - tid;
    }
    auto off = [&](int p) { return LEAN ? min(p * NT, lim) : gi[p]; };
    auto sidx = [&](int p) { return LEAN ? sbase + min(p * NT, lim) : si[p]; };
    auto is_bd = [&](int p) { return LEAN ? (mbits & ((unsigned)MK_BD << (2 * p))) != 0u : (mask[p] & MK_BD) != 0; };
    auto is_valid = [&](int p) { return LEAN ? (mbits & ((unsigned)MK_VALID << (2 * p))) != 0u : (mask[p] & MK_VALID) != 0; };
    // halo nodes (not owned by this CTA) that are boundary nodes and must be zeroed in every stage
    int hz[HZ];
    {
        const int nlo = q0 - lo, nhi = hi - min(q0 + Q, plane);
#pragma unroll
        for (int r = 0; r < HZ; r++) {
            const int hh = tid + r * NT;
            int q = -1;
            if (hh < nlo) q = lo + hh;
            else if (hh < nlo + nhi) q = min(q0 + Q, plane) + (hh - nlo);
            hz[r] = -1;
            if (q >= 0) {
                const int j = q / nx, i = q - j * nx;
                if (i == 0 || i == nx - 1 || (L.ay && (j == 0 || j == L.ny - 1))) hz[r] = 2 + (q - lo);
            }
        }
    }

    double prev[P], cur[P], nxt[P], bq[P], pq[P];
    double dv[2] = {0.0, 0.0};
#pragma unroll
    for (int p = 0; p < P; p++) { prev[p] = 0.0; cur[p] = 0.0; nxt[p] = 0.0; bq[p] = 0.0; pq[p] = 0.0; }

    // take the centre values of a landed plane, then apply the Dirichlet mask to the staged copy
    auto take_plane = [&](double *st, double (&c)[P]) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int s0 = sidx(p);
            c[p] = st[s0];
            if (is_bd(p) && is_valid(p)) st[s0] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < HZ; r++)
            if (hz[r] >= 0) st[hz[r]] = 0.0;
        fence_proxy_async();          // these generic writes precede the next bulk copy into this slot
    };

    uint32_t phase = 0;               // bit s = parity to wait for on slot s
    int slot_c = 0;                   // slot of plane t
    int adj_c = adj0;                 // parity shift of plane t
    if (kg0 >= 0 && kg0 <= L.nz - 1) {
        mbar_wait(&bars[0], 0);
        take_plane(stages + adj_c, cur);
    }
    phase ^= 1u;

    // global pointers of plane t (the plane computed in iteration t), advanced by `plane` per step
    const long long g0 = (long long)kl0 * plane + q0 + tid;
    double *outp = op.out + g0;
    constexpr bool HAS_B = (MODE == ST_LIN || MODE == ST_LIN_PM1 || MODE == ST_LIN_PM1_DOT2);
    constexpr bool HAS_PM1 = (MODE == ST_LIN_PM1 || MODE == ST_LIN_PM1_DOT2);
    const double *bp = HAS_B ? op.b + g0 : nullptr;
    const double *pp = HAS_PM1 ? op.pm1 + g0 : nullptr;
    const double diag = L.diag, cx = L.cx, cy = L.ay ? L.cy : 0.0, cz = L.cz;
    const int oy = L.ay ? nx : 0;
    const double ca = op.ca, cb = op.cb, cg = op.cg;

    // the iteration after which both boundary planes of this CTA have been pushed
    const int t_sig = ((chunk == (int)gridDim.y - 1 && op.port.flag_hi != nullptr && !rev) || (op.port.opts & 2)) ? T - 2 : 1;
    for (int t = 0; t + 1 < T; t++) {
        const int kg = kg0 + dir * t;         // global plane computed in this iteration (when t >= 1)
        int slot_n = slot_c + 1;
        if (slot_n == NS) slot_n = 0;
        const int adj_n = (adj_c + podd) & 1;
        // prefetch the streamed operands of the next plane to compute (t+1) into registers
        double bn[P], pn[P];
        if (HAS_B) {
            if (t + 1 < T - 1) {
                if (MG) {       // (runtime direction: one pointer bump instead of a 64-bit add per node)
                    const double *bnx = bp + dstep, *pnx = HAS_PM1 ? pp + dstep : nullptr;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        bn[p] = bnx[off(p)];
                        if (HAS_PM1) pn[p] = pnx[off(p)];
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        bn[p] = bp[plane + off(p)];
                        if (HAS_PM1) pn[p] = pp[plane + off(p)];
                    }
                }
            }
        }
        // centre values of plane t+1
        if (kg + dir >= 0 && kg + dir <= L.nz - 1) {
            mbar_wait(&bars[slot_n], (phase >> slot_n) & 1u);
            take_plane(stages + (size_t)slot_n * SD + adj_n, nxt);
        } else {
#pragma unroll
            for (int p = 0; p < P; p++) nxt[p] = 0.0;
        }
        phase ^= (1u << slot_n);
        if (t >= 1) {
            const double *st = stages + (size_t)slot_c * SD + adj_c;
            const bool kbd = (kg == 0 || kg == L.nz - 1);
            // first / last owned plane of the slab: the value also goes into the neighbour's ghost plane
            const bool push_lo = MG && op.port.lo_dst != nullptr && kg == L.zs;
            const bool push_hi = MG && op.port.hi_dst != nullptr && kg == L.zs + L.zm - 1;
            const double czd = (kg - 1 > 0) ? cz : 0.0, czu = (kg + 1 < L.nz - 1) ? cz : 0.0;
            // coefficients by march order: `nxt` is the plane above when going up, the plane below when going down
            const double czn = (MG && rev) ? czd : czu, czp = (MG && rev) ? czu : czd;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const bool bd = kbd || is_bd(p);
                const int s0 = sidx(p);
                const int dx = bd ? 0 : 1, dy = bd ? 0 : oy;     // boundary rows never leave their own node
                const double uc = cur[p];
                double Ai = diag * uc - cx * (st[s0 - dx] + st[s0 + dx]);
                Ai -= cy * (st[s0 - dy] + st[s0 + dy]);
                // MG: two rounded products, then their (commutative) sum -- the same bits in both march directions
                if (MG) Ai -= __dadd_rn(__dmul_rn(czn, nxt[p]), __dmul_rn(czp, prev[p]));
                else Ai -= czu * nxt[p] + czd * prev[p];
                const double Au = bd ? diag * uc : Ai;
                double o;
                if (MODE == ST_APPLY || MODE == ST_APPLY_DOT) {
                    o = Au;
                    if (MODE == ST_APPLY_DOT) dv[0] += is_valid(p) ? uc * Au : 0.0;
                } else if (MODE == ST_LIN_BU) {
                    o = cb * uc + cg * (uc - Au);
                } else {
                    o = cb * uc + cg * (bq[p] - Au);
                    if (HAS_PM1) o += ca * pq[p];
                    if (MODE == ST_LIN_PM1_DOT2 && is_valid(p)) {
                        dv[0] += o * o;
                        dv[1] += o * bq[p];
                    }
                }
                if (is_valid(p)) outp[off(p)] = o;
            }
            // first / last owned plane of the slab: the values also go into the neighbour's ghost plane.  Kept out
            // of the loop above (this happens once per kernel and CTA): the thread reads back what it just stored.
            if (MG && (push_lo || push_hi)) {
                double *dst = (push_lo ? op.port.lo_dst : op.port.hi_dst) + q0 + tid;
                double *dst2 = (push_lo && push_hi) ? op.port.hi_dst + q0 + tid : nullptr;     // one-plane slab
#pragma unroll
                for (int p = 0; p < P; p++) {
                    if (is_valid(p)) {
                        const double o = outp[off(p)];
                        dst[off(p)] = o;
                        if (dst2) dst2[off(p)] = o;
                    }
                }
            }
        }
        __syncthreads();                       // everyone is done with slot_c (and with plane 0 at t = 0)
        if (tid == 0) issue_next();
        // boundary planes are out: publish them (one thread of another warp than the copy producer's)
        if (MG && port_cta && t == t_sig && tid == NT - 32) port_arrive(op.port);
        outp += dstep;
        if (HAS_B) bp += dstep;
        if (HAS_PM1) pp += dstep;
        slot_c = slot_n;
        adj_c = adj_n;
#pragma unroll
        for (int p = 0; p < P; p++) {
            prev[p] = cur[p];
            cur[p] = nxt[p];
            if (HAS_B) {
                bq[p] = bn[p];
                if (HAS_PM1) pq[p] = pn[p];
            }
        }
    }
    if (MG && port_cta && T - 1 <= t_sig) {          // (a chunk too short to have reached t_sig inside the loop)
        __syncthreads();
        if (tid == NT - 32) port_arrive(op.port);
    }
    if (MODE == ST_APPLY_DOT || MODE == ST_LIN_PM1_DOT2) {
        // one partial (per value) per CTA, summed in fixed order by the last CTA to finish
        constexpr int NV = (MODE == ST_LIN_PM1_DOT2) ? 2 : 1;
        __shared__ double red[2][NT / 32];
        __shared__ bool is_last;
        const int lane = tid & 31, wid = tid >> 5;
        const unsigned int nblk = nbands * gridDim.y;
        const unsigned int bid = blockIdx.y * nbands + blockIdx.x;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const double s = warp_sum(dv[v]);
            if (lane == 0) red[v][wid] = s;
        }
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double t2 = (lane < NT / 32) ? red[v][lane] : 0.0;
                t2 = warp_sum(t2);
                if (lane == 0) partials[(size_t)v * nblk + bid] = t2;
            }
            if (lane == 0) {
                __threadfence();
                is_last = (atomicAdd(ticket, 1u) == nblk - 1);
            }
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double a = 0.0;
                for (unsigned int i = tid; i < nblk; i += NT) a += ((volatile double *)partials)[(size_t)v * nblk + i];
                a = warp_sum(a);
                __syncthreads();
                if (lane == 0) red[v][wid] = a;
                __syncthreads();
                if (wid == 0) {
                    double t3 = (lane < NT / 32) ? red[v][lane] : 0.0;
                    t3 = warp_sum(t3);
                    if (lane == 0) op.dot_out[v] = t3;
                }
            }
            if (tid == 0) *ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: eligibility, configuration, launch
// ---------------------------------------------------------------------------------------------
struct MarchTune { int P, NT, NS, enabled, min_plane, alt; };

static MarchTune &tune() {
    static MarchTune t = [] {
        MarchTune x = {0, 0, 4, 1, 16384, 1};   // P = 0: per-mode default; alt = 1: 7 nodes/thread where the wave model prefers it
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
    else if (k == "march_alt") t.alt = (int)v;
    else return 1;
    return 0;
}

bool stencil_fast_eligible(const LevelDesc &L) {
    const MarchTune &t = tune();
    // 3-D grids: planes of at least min_plane nodes; 2-D grids (slots x,z): rows of at least min_plane/16 nodes
    const long long plane = (long long)L.nx * L.ny;
    const bool big = L.ay ? plane >= t.min_plane : plane >= t.min_plane / 16;
    return t.enabled && L.ax && L.az && big && L.zm >= 8 && plane * (L.zm + 2) < (1LL << 31);
}

// Number of z chunks for bands of Q nodes on `slots` resident CTAs: minimise waves * (planes per chunk + pipeline fill).
// Returns that cost (in plane-steps of one wave); *nchunks = the chunk count to launch.
static double march_plan(int plane, int zm, int Q, long long slots, int *nchunks) {
    const int bands = (plane + Q - 1) / Q;
    int best_nc = 1;
    double best_cost = 1e300;
    for (int nc = 1; nc <= zm / 4 && nc <= 256; nc++) {
        const int KC = (zm + nc - 1) / nc;
        const int nce = (zm + KC - 1) / KC;
        const long long ctas = (long long)bands * nce;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (KC + 2 + 1.5);   // +1.5: pipeline fill per CTA
        if (cost < best_cost - 1e-9) { best_cost = cost; best_nc = nce; }
    }
    *nchunks = best_nc;
    return best_cost;
}
// Wave model of a whole launch: every resident CTA streams Q nodes per plane-step, so time ~ cost * Q * CTAs per SM.
// Used to choose between 8 and 7 nodes per thread: on thin slabs (multi-GPU) 129 bands x 2 chunks of Q = 2048 fill only
// 258 of the 296 CTA slots, 147 bands of Q = 1792 fill 294 (measured on a 513 x 513 x 65 slab: 16 N modes 7 % faster).
static double march_model(int plane, int zm, int P, int NT, int sm_count) {
    const int occ = NT <= 256 ? 2 : 1;
    int nc;
    return march_plan(plane, zm, P * NT, (long long)sm_count * occ, &nc) * (double)(P * NT) * occ;
}
static int sm_count_cached() {
    static thread_local int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int MODE, int P, int NT, bool MG>
static int launch_march_mg(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red, int NS) {
    // per-device facts and attributes of this instantiation, cached per host thread and device (function attributes are
    // per device: a process that drives several GPUs, one host thread each, must set them on every one)
    struct Cache { int dev = -1, sm_count = 0, max_smem = 0, occ = 0; size_t attr_smem = 0, occ_smem = 0; };
    static thread_local Cache K;
    {
        int dev = 0;
        P4B_CUDA(cudaGetDevice(&dev));
        if (K.dev != dev) {
            K = Cache();
            K.dev = dev;
            P4B_CUDA(cudaDeviceGetAttribute(&K.sm_count, cudaDevAttrMultiProcessorCount, dev));
            P4B_CUDA(cudaDeviceGetAttribute(&K.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        }
    }
    const int sm_count = K.sm_count, max_smem = K.max_smem;
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
    if (smem > K.attr_smem) {
        P4B_CUDA(cudaFuncSetAttribute(stencil_march_kernel<MODE, P, NT, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        K.attr_smem = smem;
    }
    // occupancy does not change between launches of one instantiation with one stage size: ask once
    if (!K.occ || K.occ_smem != smem) {
        int occ = 1;
        P4B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stencil_march_kernel<MODE, P, NT, MG>, NT, smem));
        K.occ = occ < 1 ? 1 : occ;
        K.occ_smem = smem;
    }
    const int occ_cached = K.occ;
    const int bands = (plane + Q - 1) / Q;
    int best_nc = 1;
    march_plan(plane, L.zm, Q, (long long)sm_count * occ_cached, &best_nc);
    cfg.KC = (L.zm + best_nc - 1) / best_nc;
    const int nchunks = (L.zm + cfg.KC - 1) / cfg.KC;
    if ((MODE == ST_APPLY_DOT || MODE == ST_LIN_PM1_DOT2) && bands * nchunks > red.max_blocks)
        return fail(63, "reducer scratch too small");
    dim3 grid(bands + (MG ? 1 : 0), nchunks);      // MG: + the column that holds the signalling CTA
    stencil_march_kernel<MODE, P, NT, MG><<<grid, NT, smem, st>>>(L, op, cfg, red.partials, red.ticket);
    P4B_LAUNCH_CHECK();
    return 0;
}

template <int MODE, int P, int NT>
static int launch_march(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red, int NS) {
    if (op.port.sync && op.port.push) return launch_march_mg<MODE, P, NT, true>(st, L, op, red, NS);
    // a kernel that only READS ghost planes keeps the lean single-GPU code: the wait for the neighbours' flags is a
    // one-thread kernel in front of it (measured: the exchange-capable variant costs the 16 N modes 8-25 %)
    if (op.port.sync) P4B_CHECK(launch_port_wait(st, op.port));
    return launch_march_mg<MODE, P, NT, false>(st, L, op, red, NS);
}

template <int P, int NT>
static int launch_march_mode(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red, int NS) {
    switch (op.mode) {
        case ST_APPLY: return launch_march<ST_APPLY, P, NT>(st, L, op, red, NS);
        case ST_APPLY_DOT: return launch_march<ST_APPLY_DOT, P, NT>(st, L, op, red, NS);
        case ST_LIN: return launch_march<ST_LIN, P, NT>(st, L, op, red, NS);
        case ST_LIN_PM1: return launch_march<ST_LIN_PM1, P, NT>(st, L, op, red, NS);
        case ST_LIN_BU: return launch_march<ST_LIN_BU, P, NT>(st, L, op, red, NS);
        case ST_LIN_PM1_DOT2: return launch_march<ST_LIN_PM1_DOT2, P, NT>(st, L, op, red, NS);
    }
    return fail(62, "unknown stencil mode %d", op.mode);
}

int launch_stencil_generic(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red);

int launch_stencil_fast(cudaStream_t st, const LevelDesc &L, const StencilOp &op, const Reducer &red) {
    if (((uintptr_t)op.u & 15) != 0) return launch_stencil_generic(st, L, op, red);   // bulk copies need 16-byte alignment
    const MarchTune &t = tune();
    if (t.P == 0) {
        // measured on B200 at 513^3 (profiles/r01_march_tuning.md): the register-heavy three-operand mode
        // wants 4 nodes/thread x 512 threads, the others 8 nodes/thread x 256 threads (2 CTAs/SM)
        if (op.mode == ST_LIN_PM1) return launch_march<ST_LIN_PM1, 4, 512>(st, L, op, red, t.NS);
        if (op.mode == ST_LIN_PM1_DOT2) return launch_march<ST_LIN_PM1_DOT2, 4, 512>(st, L, op, red, t.NS);
        // with the slab-exchange code the two-operand mode no longer fits 128 registers at 8 nodes/thread; at 7 it does
        // (measured on thin slabs, profiles/r02_exchange.md: 0.084 ms against 0.098 ms with 4 x 512)
        if (op.mode == ST_LIN && op.port.sync && op.port.push) return launch_march<ST_LIN, 7, 256>(st, L, op, red, t.NS);
        // the 16 N modes are the ones the wave model describes (measured: the two-operand mode is as fast either way on
        // thin slabs and 8 % slower with 7 nodes/thread on 513 planes)
        const int plane = L.nx * L.ny, sms = sm_count_cached();
        const bool one_operand = op.mode == ST_APPLY || op.mode == ST_APPLY_DOT || op.mode == ST_LIN_BU;
        if (t.alt && one_operand && march_model(plane, L.zm, 7, 256, sms) < 0.97 * march_model(plane, L.zm, 8, 256, sms))
            return launch_march_mode<7, 256>(st, L, op, red, t.NS);
        return launch_march_mode<8, 256>(st, L, op, red, t.NS);
    }
    if (t.P == 4 && t.NT == 512) return launch_march_mode<4, 512>(st, L, op, red, t.NS);
    if (t.P == 8 && t.NT == 256) return launch_march_mode<8, 256>(st, L, op, red, t.NS);
    if (t.P == 7 && t.NT == 256) return launch_march_mode<7, 256>(st, L, op, red, t.NS);
    if (t.P == 7 && t.NT == 512) return launch_march_mode<7, 512>(st, L, op, red, t.NS);
    return fail(62, "P4B_MARCH=%d,%d is not instantiated", t.P, t.NT);
}

}  // namespace p4b
