// mg.cu -- context, level hierarchy, multigrid cycle, preconditioned CG and the C ABI (p4b200.h).
//
// Restates, B200-first, what PETSc does for `./fish -pc_type mg`:
//   PCSetUp_MG   -> p4b_mg_create : levels by DMDA coarsening, rediscretised constant-coefficient
//                   operators (fish.c:7), Chebyshev targets, dense inverse of the coarsest operator
//   PCApply_MG   -> Mg::cycle     : multiplicative V/W cycle, R = P^T, Chebyshev/Jacobi smoothing
//   KSPSolve_CG  -> Mg::cg        : preconditioned-norm CG; all scalars stay on the device, one
//                   8-byte D2H per iteration for the convergence test
// Multi-GPU: slabs of the slowest dimension, ghost planes by grouped ncclSend/ncclRecv, dot products
// by ncclAllReduce, small levels replicated (the PCREDUNDANT idea applied to whole levels).
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <string.h>
#include <unistd.h>

#include <mutex>
#include <vector>

#include <array>

#include "comm.h"
#include <cstddef>
#include "kernels.h"

namespace p4b {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
long long g_launch_count = 0;
void set_error(const std::string &m) { g_err = m; }
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code ? code : 1;
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time NCCL dependency, so a single-GPU host never
// needs it and a torch process reuses the libnccl.so.2 torch already loaded.
// ------------------------------------------------------------------------------------------------
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;

static std::mutex g_nccl_mutex;
static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(80, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LD(field, sym)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                           \
    if (!g_nccl.field) return fail(80, "libnccl lacks symbol %s", sym)
    LD(GetUniqueId, "ncclGetUniqueId");
    LD(CommInitRank, "ncclCommInitRank");
    LD(CommDestroy, "ncclCommDestroy");
    LD(AllReduce, "ncclAllReduce");
    LD(Broadcast, "ncclBroadcast");
    LD(AllGather, "ncclAllGather");
    LD(Send, "ncclSend");
    LD(Recv, "ncclRecv");
    LD(GroupStart, "ncclGroupStart");
    LD(GroupEnd, "ncclGroupEnd");
    LD(GetErrorString, "ncclGetErrorString");
#undef LD
    g_nccl.h = h;
    return 0;
}

#define P4B_NCCL(call)                                                                                 \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess)                                                                         \
            return p4b::fail(81, "%s:%d NCCL error %d (%s) in %s", __FILE__, __LINE__, (int)r_,        \
                             g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?", #call);          \
    } while (0)

}  // namespace p4b

using namespace p4b;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct p4b_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Reducer red;
    double *d_scal = nullptr;      // 16 device doubles: CG scalars
    double *h_scal = nullptr;      // 16 pinned host doubles
    double *d_mdot = nullptr;      // 64 doubles, on first use (p4b_vec_mdot)
    HostPoll *h_poll = nullptr, *d_poll = nullptr;    // pinned host memory the kernels publish scalars into / its device alias
    unsigned long long poll_seq = 0;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // peer-memory path (comm.cu): mailboxes of all ranks mapped through CUDA IPC
    bool peer = false;
    Mailbox *mbox = nullptr;
    bool mbox_ipc[MAX_RANKS] = {};      // which peer mailboxes were opened through CUDA IPC (other processes)
    LocalSync *sync = nullptr;
    PeerTable peers;
    // p4b_tune("force_mg"): single-GPU stand-ins for the neighbours (kernel A/B measurements only)
    double *dummy_planes = nullptr;
    size_t dummy_plane_cap = 0;
    unsigned long long *dummy_flags = nullptr;
};

static long long g_comm_peer = 1;      // p4b_tune("comm_peer", 0) keeps everything on NCCL

// Map one device allocation of every rank into this rank.  Ranks in other PROCESSES (torchrun, one process per GPU) are
// mapped through CUDA IPC; ranks in THIS process (the PETSc-shaped shim driving N GPUs with one host thread each,
// -p4b_gpus N) are reached through the raw pointer after cudaDeviceEnablePeerAccess -- an IPC handle cannot be opened
// by the process that exported it.  The records travel over NCCL as raw bytes.  own[r] says which mappings have to be
// closed with cudaIpcCloseMemHandle.
struct PeerRecord {
    cudaIpcMemHandle_t handle;
    long long pid;
    void *ptr;
    int device;
};
static int map_peers(p4b_ctx *c, void *mine, void **mapped, bool *ipc_opened) {
    PeerRecord rec;
    memset(&rec, 0, sizeof rec);
    P4B_CUDA(cudaIpcGetMemHandle(&rec.handle, mine));
    rec.pid = (long long)getpid();
    rec.ptr = mine;
    rec.device = c->device;
    const size_t hb = sizeof(PeerRecord);
    char *d = nullptr;
    P4B_CUDA(cudaMalloc(&d, hb * c->nranks));
    P4B_CUDA(cudaMemcpy(d + hb * c->rank, &rec, hb, cudaMemcpyHostToDevice));
    P4B_NCCL(g_nccl.AllGather(d + hb * c->rank, d, hb, ncclChar, c->comm, c->stream));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<PeerRecord> all(c->nranks);
    P4B_CUDA(cudaMemcpy(all.data(), d, hb * c->nranks, cudaMemcpyDeviceToHost));
    P4B_CUDA(cudaFree(d));
    for (int r = 0; r < c->nranks; r++) {
        ipc_opened[r] = false;
        if (r == c->rank) { mapped[r] = mine; continue; }
        if (all[r].pid == rec.pid) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(72, "cudaDeviceEnablePeerAccess(%d) failed: %s", all[r].device, cudaGetErrorString(e));
            cudaGetLastError();
            mapped[r] = all[r].ptr;
        } else {
            void *p = nullptr;
            P4B_CUDA(cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
            mapped[r] = p;
            ipc_opened[r] = true;
        }
    }
    return 0;
}

static int nccl_barrier(p4b_ctx *c) {
    if (c->nranks == 1) return 0;
    P4B_NCCL(g_nccl.AllReduce(c->d_scal + 15, c->d_scal + 15, 1, ncclFloat64, ncclSum, c->comm, c->stream));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int peer_setup(p4b_ctx *c) {
    if (c->nranks == 1 || c->nranks > MAX_RANKS || !g_comm_peer) return 0;
    P4B_CUDA(cudaMalloc(&c->mbox, sizeof(Mailbox)));
    P4B_CUDA(cudaMemset(c->mbox, 0, sizeof(Mailbox)));
    P4B_CUDA(cudaMalloc(&c->sync, sizeof(LocalSync)));
    P4B_CUDA(cudaMemset(c->sync, 0, sizeof(LocalSync)));
    c->peers.rank = c->rank;
    c->peers.nranks = c->nranks;
    void *mapped[MAX_RANKS];
    P4B_CHECK(map_peers(c, c->mbox, mapped, c->mbox_ipc));
    for (int r = 0; r < c->nranks; r++) c->peers.mbox[r] = (Mailbox *)mapped[r];
    P4B_CHECK(nccl_barrier(c));
    c->peer = true;
    return 0;
}

static int ctx_allreduce_op(p4b_ctx *c, double *d, int count, int op_max) {
    if (c->nranks == 1) return 0;
    if (c->peer && count <= 3) return launch_allreduce(c->stream, d, count, op_max, c->peers, c->sync);
    if (c->peer) {      // longer vectors (p4b_vec_mdot): three values at a time
        for (int q = 0; q < count; q += 3)
            P4B_CHECK(launch_allreduce(c->stream, d + q, count - q < 3 ? count - q : 3, op_max, c->peers, c->sync));
        return 0;
    }
    P4B_NCCL(g_nccl.AllReduce(d, d, (size_t)count, ncclFloat64, op_max ? ncclMax : ncclSum, c->comm, c->stream));
    return 0;
}
static int ctx_allreduce(p4b_ctx *c, double *d, int count) { return ctx_allreduce_op(c, d, count, 0); }

// Wait until the kernels have published sequence number `seq` into the context's HostPoll, then copy `count` values.
// The host spins on pinned memory (no cudaStreamSynchronize); every ~65k spins it asks the stream whether it died.
static int poll_wait(p4b_ctx *c, unsigned long long seq, int count, double *h) {
    volatile HostPoll *hp = c->h_poll;
    unsigned long spins = 0;
    while (__atomic_load_n(&hp->seq, __ATOMIC_ACQUIRE) < seq) {
        if ((++spins & 0xFFFFul) == 0) {
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return fail(70, "CUDA error %d (%s) while waiting for a device scalar", (int)e, cudaGetErrorString(e));
            if (e == cudaSuccess && __atomic_load_n(&hp->seq, __ATOMIC_ACQUIRE) < seq)
                return fail(70, "stream is idle but the device scalar was never published (sequence %llu)", seq);
        }
        __builtin_ia32_pause();
    }
    for (int i = 0; i < count; i++) h[i] = hp->v[i];
    return 0;
}

// all-reduce `count` device scalars (sum or max) and hand them to the host: on the peer path one kernel does both
static int allreduce_fetch(p4b_ctx *c, double *d, int count, int op_max, double *h) {
    const unsigned long long seq = ++c->poll_seq;
    if (c->nranks > 1 && c->peer && count <= 3) {
        P4B_CHECK(launch_allreduce(c->stream, d, count, op_max, c->peers, c->sync, c->d_poll, seq));
    } else {
        P4B_CHECK(ctx_allreduce_op(c, d, count, op_max));
        P4B_CHECK(launch_publish(c->stream, d, count, c->d_poll, seq));
    }
    return poll_wait(c, seq, count, h);
}

// ------------------------------------------------------------------------------------------------
// grid -> level descriptor
// ------------------------------------------------------------------------------------------------
static int make_desc(const p4b_grid *g, LevelDesc *L) {
    if (!g || g->dim < 1 || g->dim > 3) return fail(1, "invalid dim for DMDA creation");
    if (g->cx <= 0 || g->cy <= 0 || g->cz <= 0) return fail(2, "positivity required for coefficients cx,cy,cz");
    const int m[3] = {g->mx, g->dim >= 2 ? g->my : 1, g->dim >= 3 ? g->mz : 1};
    for (int d = 0; d < g->dim; d++)
        if (m[d] < 3) return fail(60, "grid needs at least 3 nodes per dimension (got %d)", m[d]);
    memset(L, 0, sizeof *L);
    const double hx = g->Lx / (m[0] - 1);
    const double hy = g->dim >= 2 ? g->Ly / (m[1] - 1) : 1.0;
    const double hz = g->dim >= 3 ? g->Lz / (m[2] - 1) : 1.0;
    if (g->dim == 1) {   // poissonfunctions.c:13-21,130-137
        L->nx = m[0]; L->ny = 1; L->nz = 1;
        L->ax = 1; L->ay = 0; L->az = 0;
        L->cx = g->cx / hx; L->cy = 0; L->cz = 0;
        L->diag = g->cx * 2.0 / hx;
        L->vol = hx;
        L->hx = hx; L->hy = 1; L->hz = 1;
    } else if (g->dim == 2) {   // :37-40 ; the y direction lives in slot z
        L->nx = m[0]; L->ny = 1; L->nz = m[1];
        L->ax = 1; L->ay = 0; L->az = 1;
        const double scx = g->cx * hy / hx, scy = g->cy * hx / hy;
        L->cx = scx; L->cy = 0; L->cz = scy;
        L->diag = 2.0 * (scx + scy);
        L->vol = hx * hy;
        L->hx = hx; L->hy = 1; L->hz = hy;
    } else {   // :77-81
        L->nx = m[0]; L->ny = m[1]; L->nz = m[2];
        L->ax = L->ay = L->az = 1;
        const double dvol = hx * hy * hz;
        L->cx = g->cx * dvol / (hx * hx);
        L->cy = g->cy * dvol / (hy * hy);
        L->cz = g->cz * dvol / (hz * hz);
        L->diag = 2.0 * (L->cx + L->cy + L->cz);
        L->vol = dvol;
        L->hx = hx; L->hy = hy; L->hz = hz;
    }
    L->zs = 0;
    L->zm = L->nz;
    return 0;
}

static bool can_coarsen(const p4b_grid &g) {
    const int m[3] = {g.mx, g.my, g.mz};
    for (int d = 0; d < g.dim; d++)
        if (m[d] <= 3 || (m[d] - 1) % 2) return false;
    return true;
}
static p4b_grid coarsen(const p4b_grid &g) {
    p4b_grid c = g;
    c.mx = (g.mx - 1) / 2 + 1;
    if (g.dim >= 2) c.my = (g.my - 1) / 2 + 1;
    if (g.dim >= 3) c.mz = (g.mz - 1) / 2 + 1;
    return c;
}

static double lambda_max(const LevelDesc &L) {
    const double PI = 3.14159265358979323846;
    double num = 0, den = 0;
    if (L.ax) { num += L.cx * cos(PI / (L.nx - 1)); den += L.cx; }
    if (L.ay) { num += L.cy * cos(PI / (L.ny - 1)); den += L.cy; }
    if (L.az) { num += L.cz * cos(PI / (L.nz - 1)); den += L.cz; }
    return 1.0 + num / den;
}

// ------------------------------------------------------------------------------------------------
// hierarchy
// ------------------------------------------------------------------------------------------------
struct Level {
    p4b_grid g;
    LevelDesc d;                 // what the kernels see: this rank's slab (distributed) or the whole grid (replicated)
    LevelDesc own;               // replicated level fed from a distributed one: the part this rank restricts into
    bool replicated = false;
    double emin = 0, emax = 0, lam = 0;
    std::vector<double> omega;   // Chebyshev omega_i, i >= 1
    double scale = 0;            // 2/(emax+emin)
    double *x = nullptr, *b = nullptr, *t = nullptr;   // first owned element of each ghosted vector
    std::vector<int> zs_all, zm_all;                   // ownership of every rank (2K rule) on this level
    // arena offsets (in doubles) of the first owned element of this level's buffers: here and on the slab neighbours
    std::array<size_t, 7> off_me{}, off_prev{}, off_next{};
    int nslots = 3;
    int zm_prev = 0, zm_next = 0;
};

struct Prof {
    int on = 0;                  // 1 = finest level, 2 = trace (every level, no graph)
    struct Rec { int cls; int level; cudaEvent_t a, b; double bytes; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    cudaEvent_t last = nullptr;  // the newest recorded event, reusable as the next bracket's start while
    long long last_launch = -1;  // g_launch_count still has this value (nothing launched since)
    p4b_kernel_stat stat[P4B_K_NCLASSES];
    p4b_kernel_stat lstat[P4B_MAX_LEVELS][P4B_K_NCLASSES];
};

struct p4b_mg {
    p4b_ctx *ctx = nullptr;
    p4b_mg_opts o;
    std::vector<Level> lev;      // lev[0] coarsest
    int top = 0;
    double *arena = nullptr;
    size_t arena_doubles = 0;
    double *Ainv = nullptr;      // dense inverse of the coarsest operator
    int n0 = 0;
    double *p = nullptr, *w = nullptr, *fbuf = nullptr, *gbuf = nullptr;   // CG vectors / fish scratch
    Prof prof;
    // peer-memory path: arenas of all ranks mapped through CUDA IPC
    bool peer = false;
    GatherTable peer_arena;
    bool arena_ipc[MAX_RANKS] = {};
    const double *last_halo = nullptr;     // the vector exchanged most recently (hazard tracking, comm.cu "Hazards")
    const double *pushed = nullptr;        // the vector whose ghost planes are current (exchanged, not written since)
    bool dot2_fused = false;     // set by smooth() when the last smoother kernel also produced (z,z), (z,r)
    double *dot2_target = nullptr;
    cudaGraphExec_t coarse_graph = nullptr;   // the whole sub-cycle below the finest level, captured once
    long long graph_kernels = 0;              // kernel launches one replay stands for
    const double *graph_last_halo = nullptr;  // host-side exchange tracking as the captured sweep leaves it
    const double *graph_pushed = nullptr;
    bool graph_failed = false;
};

static double alg_bytes(int cls, double N, double Nc) {
    switch (cls) {
        case P4B_K_APPLY_DOT: return 16 * N;
        case P4B_K_RESIDUAL: return 24 * N;
        case P4B_K_CHEB_ZERO: return 16 * N;
        case P4B_K_CHEB_FIRST: return 24 * N;
        case P4B_K_CHEB_NEXT: return 32 * N;
        case P4B_K_RESTRICT: return 8 * N + 8 * Nc;
        case P4B_K_PROLONG: return 16 * N + 8 * Nc;
        case P4B_K_AXPY2: return 48 * N;
        case P4B_K_DOT2: return 16 * N;
        case P4B_K_AYPX: return 24 * N;
        case P4B_K_RESID_RESTRICT: return 16 * N + 8 * Nc;
        case P4B_K_XP_UPDATE: return 40 * N;
        case P4B_K_R_UPDATE: return 24 * N;
    }
    return 0;
}

static const char *k_names[P4B_K_NCLASSES] = {"apply_dot", "residual", "cheb_zero", "cheb_first", "cheb_next", "restrict",
                                              "prolong_add", "axpy2", "dot2", "aypx", "resid_restrict", "xp_update", "r_update",
                                              "halo", "gather", "allreduce", "coarse_solve", "subcycle"};

// Profiling bracket: only finest-level launches are timed (prof.on == 1), or every launch (2 = trace).  Consecutive
// brackets share one event (the end of a kernel is the start of the next when nothing was launched in between), which
// halves the number of event records the GPU front end has to process between kernels -- at 8 GPUs, with ~60 us
// kernels, two records per boundary were a measurable part of the step.  A host synchronisation point breaks the chain
// (prof_break), so host latency is never booked on the next kernel.
struct ProfScope {
    p4b_mg *m;
    int idx = -1;
    static cudaEvent_t next_event(Prof &P) {
        if (P.pool.size() < P.used + 1) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            P.pool.push_back(e);
        }
        return P.pool[P.used++];
    }
    ProfScope(p4b_mg *mg, int level, int cls) : m(mg) {
        if (!m->prof.on || (m->prof.on == 1 && level != m->top)) return;
        Prof &P = m->prof;
        Prof::Rec r;
        r.cls = cls;
        r.level = level;
        if (P.last && P.last_launch == g_launch_count) {
            r.a = P.last;
        } else {
            r.a = next_event(P);
            cudaEventRecord(r.a, m->ctx->stream);
            P.last = r.a;
            P.last_launch = g_launch_count;
        }
        r.b = nullptr;
        const Level &L = m->lev[level];
        const double N = (double)L.d.nlocal();
        const double Nc = level > 0 ? (double)m->lev[level - 1].d.nglobal() / (m->ctx->nranks) : 0.0;
        r.bytes = alg_bytes(cls, N, Nc);
        idx = (int)P.recs.size();
        P.recs.push_back(r);
    }
    ~ProfScope() {
        if (idx < 0) return;
        Prof &P = m->prof;
        if (P.last && P.last_launch == g_launch_count) {
            P.recs[idx].b = P.last;           // an inner bracket ended on the same launch
        } else {
            P.recs[idx].b = next_event(P);
            cudaEventRecord(P.recs[idx].b, m->ctx->stream);
            P.last = P.recs[idx].b;
            P.last_launch = g_launch_count;
        }
    }
};
static inline void prof_break(p4b_mg *m) { m->prof.last = nullptr; }

static void prof_collect(p4b_mg *m) {
    Prof &P = m->prof;
    if (!P.on) return;
    cudaStreamSynchronize(m->ctx->stream);
    for (auto &r : P.recs) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        if (r.level == m->top) {
            P.stat[r.cls].launches++;
            P.stat[r.cls].ms += ms;
            P.stat[r.cls].bytes += r.bytes;
        }
        p4b_kernel_stat &ls = P.lstat[r.level][r.cls];
        ls.launches++;
        ls.ms += ms;
        ls.bytes += r.bytes;
    }
    P.recs.clear();
    P.used = 0;
    P.last = nullptr;
}

// ghost-plane exchange of a ghosted vector on a distributed level ([PETSc] DMGlobalToLocal)
static int halo(p4b_mg *m, int l, double *v) {
    p4b_ctx *c = m->ctx;
    Level &L = m->lev[l];
    if (c->nranks == 1 || L.replicated) return 0;
    const size_t plane = (size_t)L.d.plane();
    const bool lo = L.d.zs > 0, hi = L.d.zs + L.d.zm < L.d.nz;
    if (m->peer && v == m->pushed) return 0;      // the producing kernel pushed them; the consumer's port waits
    ProfScope ps(m, l, P4B_K_HALO);
    if (m->peer) {
        // push my boundary planes into the neighbours' ghost planes (comm.cu)
        int slot = -1;
        const size_t voff = (size_t)(v - m->arena);
        for (int q = 0; q < L.nslots; q++)
            if (L.off_me[q] == voff) slot = q;
        if (slot < 0) return fail(64, "halo: vector is not one of this level's arena buffers");
        if (v == m->last_halo) P4B_CHECK(launch_barrier(c->stream, 0, c->peers, c->sync));   // see comm.cu "Hazards"
        m->last_halo = v;
        m->pushed = v;
        double *lo_dst = lo ? m->peer_arena.base[c->rank - 1] + L.off_prev[slot] + (size_t)L.zm_prev * plane : nullptr;
        double *hi_dst = hi ? m->peer_arena.base[c->rank + 1] + L.off_next[slot] - plane : nullptr;
        unsigned long long *fp = lo ? &c->peers.mbox[c->rank - 1]->halo_flag[1] : nullptr;
        unsigned long long *fn = hi ? &c->peers.mbox[c->rank + 1]->halo_flag[0] : nullptr;
        return launch_halo_push(c->stream, v, lo_dst, v + (size_t)(L.d.zm - 1) * plane, hi_dst, (long long)plane, fp, fn,
                                c->mbox->halo_flag, c->sync);
    }
    P4B_NCCL(g_nccl.GroupStart());
    if (lo) {
        P4B_NCCL(g_nccl.Send(v, plane, ncclFloat64, c->rank - 1, c->comm, c->stream));
        P4B_NCCL(g_nccl.Recv(v - plane, plane, ncclFloat64, c->rank - 1, c->comm, c->stream));
    }
    if (hi) {
        P4B_NCCL(g_nccl.Send(v + (size_t)(L.d.zm - 1) * plane, plane, ncclFloat64, c->rank + 1, c->comm, c->stream));
        P4B_NCCL(g_nccl.Recv(v + (size_t)L.d.zm * plane, plane, ncclFloat64, c->rank + 1, c->comm, c->stream));
    }
    P4B_NCCL(g_nccl.GroupEnd());
    return 0;
}

static long long g_force_mg = 0;       // p4b_tune("force_mg", 1|2): run the multi-GPU kernel variants on ONE GPU for
                                       // measurements: 1 = no neighbours, 2 = both "neighbours" are scratch memory
                                       // on this device (peer stores, fences and flag traffic stay local)
static long long g_port_opts = 0;      // p4b_tune("port_opts", bits): HaloPort::opts experiments
static long long g_fused_halo = 1;     // p4b_tune("fused_halo", 0): every exchange is a kernel of its own again (A/B)

// a kernel wrote `v` without pushing its boundary planes: its ghost copies on the neighbours are stale
static void wrote(p4b_mg *m, const double *v) {
    if (m->pushed == v) m->pushed = nullptr;
}

// The HaloPort (comm.h) of a kernel launched on level l that writes the ghosted vector `out` (nullptr: it only
// reads ghost planes): wait at its start, push of out's boundary planes, signal at its end.  Inactive (all zero)
// without the peer-memory transport, on replicated levels and when out is not one of the level's arena vectors --
// then the caller's halo() before the consumer does the exchange as a kernel of its own.
static int make_port(p4b_mg *m, int l, double *out, HaloPort *hp) {
    memset(hp, 0, sizeof *hp);
    p4b_ctx *c = m->ctx;
    Level &L = m->lev[l];
    if (out) wrote(m, out);
    if (g_force_mg && c->nranks == 1 && !L.replicated && L.d.az) {
        const size_t plane = (size_t)L.d.plane();
        if (!c->sync) {
            P4B_CUDA(cudaMalloc(&c->sync, sizeof(LocalSync)));
            P4B_CUDA(cudaMemset(c->sync, 0, sizeof(LocalSync)));
            P4B_CUDA(cudaMalloc(&c->dummy_flags, 4 * sizeof(unsigned long long)));
            const unsigned long long init[4] = {1ull << 62, 1ull << 62, 0ull, 0ull};
            P4B_CUDA(cudaMemcpy(c->dummy_flags, init, sizeof init, cudaMemcpyHostToDevice));
        }
        if (plane > c->dummy_plane_cap) {
            P4B_CUDA(cudaStreamSynchronize(c->stream));
            if (c->dummy_planes) cudaFree(c->dummy_planes);
            P4B_CUDA(cudaMalloc(&c->dummy_planes, 2 * plane * sizeof(double)));
            c->dummy_plane_cap = plane;
        }
        hp->sync = c->sync;
        hp->opts = (int)g_port_opts;
        hp->my_flags = c->dummy_flags;
        hp->plane = (long long)plane;
        hp->hi_start = (long long)(L.d.zm - 1) * (long long)plane;
        hp->push = out != nullptr;
        if (g_force_mg >= 2) {
            hp->flag_lo = c->dummy_flags + 2;
            hp->flag_hi = c->dummy_flags + 3;
            if (out) { hp->lo_dst = c->dummy_planes; hp->hi_dst = c->dummy_planes + plane; }
        }
        return 0;
    }
    if (!m->peer || c->nranks == 1 || L.replicated || !g_fused_halo) return 0;
    const bool lo = L.d.zs > 0, hi = L.d.zs + L.d.zm < L.d.nz;
    hp->sync = c->sync;
    hp->opts = (int)g_port_opts;
    hp->my_flags = c->mbox->halo_flag;
    hp->flag_lo = lo ? &c->peers.mbox[c->rank - 1]->halo_flag[1] : nullptr;
    hp->flag_hi = hi ? &c->peers.mbox[c->rank + 1]->halo_flag[0] : nullptr;
    const size_t plane = (size_t)L.d.plane();
    hp->plane = (long long)plane;
    hp->hi_start = (long long)(L.d.zm - 1) * (long long)plane;
    if (!out) return 0;
    int slot = -1;
    const size_t voff = (size_t)(out - m->arena);
    for (int q = 0; q < L.nslots; q++)
        if (L.off_me[q] == voff) slot = q;
    if (slot < 0) return 0;
    if (out == m->last_halo) P4B_CHECK(launch_barrier(c->stream, 0, c->peers, c->sync));   // comm.cu "Hazards"
    m->last_halo = out;
    m->pushed = out;
    hp->push = 1;
    hp->lo_dst = lo ? m->peer_arena.base[c->rank - 1] + L.off_prev[slot] + (size_t)L.zm_prev * plane : nullptr;
    hp->hi_dst = hi ? m->peer_arena.base[c->rank + 1] + L.off_next[slot] - plane : nullptr;
    return 0;
}

// make a replicated level's vector complete on every rank: each rank contributes the planes it owns
static int gather_replicated(p4b_mg *m, int l, double *v) {
    p4b_ctx *c = m->ctx;
    Level &L = m->lev[l];
    if (c->nranks == 1) return 0;
    const size_t plane = (size_t)L.d.plane();
    ProfScope ps(m, l, P4B_K_GATHER);
    if (m->peer) {
        // replicated levels are carved first and have the same size everywhere: same arena offset on every rank
        const long long off = (long long)(v - m->arena) + (long long)L.own.zs * (long long)plane;
        P4B_CHECK(launch_barrier(c->stream, 1, c->peers, c->sync));     // everyone is done reading the old copy
        P4B_CHECK(launch_gather_push(c->stream, v + (size_t)L.own.zs * plane, (long long)L.own.zm * (long long)plane, off,
                                     m->peer_arena, c->rank, c->nranks));
        return launch_barrier(c->stream, 1, c->peers, c->sync);          // every contribution has landed
    }
    P4B_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < c->nranks; r++) {
        if (L.zm_all[r] <= 0) continue;
        double *part = v + (size_t)L.zs_all[r] * plane;
        P4B_NCCL(g_nccl.Broadcast(part, part, (size_t)L.zm_all[r] * plane, ncclFloat64, r, c->comm, c->stream));
    }
    P4B_NCCL(g_nccl.GroupEnd());
    return 0;
}

// ---- smoothers ---------------------------------------------------------------------------------
static int smooth(p4b_mg *m, int l, bool zero_guess, double *dot2_out = nullptr, bool result_exchanged = true) {
    Level &L = m->lev[l];
    cudaStream_t st = m->ctx->stream;
    const Reducer &red = m->ctx->red;
    const int its = m->o.smooth_its;
    if (its <= 0) {
        if (zero_guess) {
            wrote(m, L.x);
            P4B_CHECK(launch_set(st, L.d.nlocal(), 0.0, L.x));
        }
        return 0;
    }
    StencilOp op;
    memset(&op, 0, sizeof op);
    const double s1 = (m->o.smoother == P4B_SMOOTH_RICHARDSON ? 1.0 : L.scale) / L.d.diag;
    if (m->o.smoother == P4B_SMOOTH_RICHARDSON) {
        // [PETSc] KSPSolve_Richardson: x <- x + B (b - A x), B = D^-1
        for (int i = 0; i < its; i++) {
            if (i == 0 && zero_guess) {
                ProfScope ps(m, l, P4B_K_CHEB_ZERO);
                wrote(m, L.t);
                P4B_CHECK(launch_scale_copy(st, L.d.nlocal(), s1, L.b, L.t));
            } else {
                P4B_CHECK(halo(m, l, L.x));
                ProfScope ps(m, l, P4B_K_CHEB_FIRST);
                op.mode = ST_LIN; op.u = L.x; op.b = L.b; op.out = L.t; op.cb = 1.0; op.cg = s1;
                P4B_CHECK(make_port(m, l, L.t, &op.port));
                P4B_CHECK(launch_stencil(st, L.d, op, red));
            }
            std::swap(L.x, L.t);
        }
        return 0;
    }
    // [PETSc] KSPSolve_Chebyshev, first kind; `its` = number of preconditioner applications
    if (zero_guess && m->o.fuse && its == 2) {
        // p1 = s1 b ; p2 = w p1 + w s1 (b - A p1)  ==  cb*b + cg*(b - A b)  with cg = w s1^2, cb = w s1 + w s1 - cg
        const double w1 = L.omega[0];
        const double cg = w1 * s1 * s1;
        const double cb = 2.0 * w1 * s1 - cg;
        P4B_CHECK(halo(m, l, L.b));
        ProfScope ps(m, l, P4B_K_CHEB_ZERO);
        op.mode = ST_LIN_BU; op.u = L.b; op.out = L.x; op.cb = cb; op.cg = cg;
        P4B_CHECK(make_port(m, l, L.x, &op.port));
        return launch_stencil(st, L.d, op, red);
    }
    double *pm1 = L.x, *pk = L.t;
    bool pm1_zero = zero_guess;
    if (zero_guess) {
        ProfScope ps(m, l, P4B_K_CHEB_ZERO);
        wrote(m, pk);
        P4B_CHECK(launch_scale_copy(st, L.d.nlocal(), s1, L.b, pk));
    } else {
        P4B_CHECK(halo(m, l, pm1));
        ProfScope ps(m, l, P4B_K_CHEB_FIRST);
        op.mode = ST_LIN; op.u = pm1; op.b = L.b; op.out = pk; op.cb = 1.0; op.cg = s1;
        P4B_CHECK(make_port(m, l, pk, &op.port));
        P4B_CHECK(launch_stencil(st, L.d, op, red));
    }
    for (int i = 1; i < its; i++) {
        const double w = L.omega[i - 1];
        P4B_CHECK(halo(m, l, pk));
        ProfScope ps(m, l, P4B_K_CHEB_NEXT);
        op.u = pk; op.b = L.b; op.out = pm1; op.cb = w; op.cg = w * s1;
        if (pm1_zero) {
            op.mode = ST_LIN;
        } else {
            op.mode = ST_LIN_PM1; op.pm1 = pm1; op.ca = 1.0 - w;
            if (dot2_out && i == its - 1) {     // last step of the cycle: also (z,z), (z,r) for KSPSolve_CG
                op.mode = ST_LIN_PM1_DOT2; op.dot_out = dot2_out;
                m->dot2_fused = true;
            }
        }
        // the last iterate's ghost planes are not needed when nothing applies the operator to it next
        if (i == its - 1 && !result_exchanged) {
            P4B_CHECK(make_port(m, l, nullptr, &op.port));
            wrote(m, pm1);
        } else {
            P4B_CHECK(make_port(m, l, pm1, &op.port));
        }
        P4B_CHECK(launch_stencil(st, L.d, op, red));
        std::swap(pm1, pk);
        pm1_zero = false;
    }
    // the newest iterate is in pk
    if (pk != L.x) std::swap(L.x, L.t);
    return 0;
}

static int coarse_solve(p4b_mg *m) {
    Level &L = m->lev[0];
    if (!m->Ainv)
        return fail(61, "coarsest grid has %lld nodes (> 2400): increase -pc_mg_levels", (long long)L.d.nglobal());
    ProfScope ps(m, 0, P4B_K_COARSE);
    return launch_dense_matvec(m->ctx->stream, m->n0, m->Ainv, L.b, L.x);
}

static int cycle(p4b_mg *m, int l, bool zero_guess);

// The levels below the finest are launch-latency bound (L2-resident grids, ~6 kernels each), so their whole
// down-and-up sweep is captured once into a CUDA graph and replayed; the finest-level kernels stay ordinary
// launches (they are timed individually by the profiler).  Needs a capturable (non-default) stream, no net
// buffer swaps per smoother call (even -mg_levels_ksp_max_it) and kernel-only communication (peer path).
static bool graph_usable(const p4b_mg *m) {
    return m->o.use_graph && m->prof.on != 2 && !m->graph_failed && m->top >= 2 && m->ctx->stream != nullptr &&
           m->ctx->stream != cudaStreamLegacy && (m->o.smooth_its % 2 == 0) && (m->ctx->nranks == 1 || m->peer);
}

static int coarse_cycle_graph(p4b_mg *m) {
    cudaStream_t st = m->ctx->stream;
    if (!m->coarse_graph) {
        const long long before = g_launch_count;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            m->graph_failed = true;
            return cycle(m, m->top - 1, true);
        }
        const int rc = cycle(m, m->top - 1, true);
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc || e != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            m->graph_failed = true;
            if (rc) return rc;
            return cycle(m, m->top - 1, true);
        }
        m->graph_kernels = g_launch_count - before;
        g_launch_count = before;
        m->graph_last_halo = m->last_halo;
        m->graph_pushed = m->pushed;
        const cudaError_t e2 = cudaGraphInstantiate(&m->coarse_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            m->coarse_graph = nullptr;
            m->graph_failed = true;
            return cycle(m, m->top - 1, true);
        }
    }
    P4B_CUDA(cudaGraphLaunch(m->coarse_graph, st));
    g_launch_count += m->graph_kernels;
    m->last_halo = m->graph_last_halo;
    m->pushed = m->graph_pushed;
    return 0;
}

static int cycle(p4b_mg *m, int l, bool zero_guess) {
    if (l == 0) return coarse_solve(m);
    Level &L = m->lev[l];
    Level &C = m->lev[l - 1];
    cudaStream_t st = m->ctx->stream;
    const Reducer &red = m->ctx->red;
    P4B_CHECK(smooth(m, l, zero_guess));
    {   // r = b - A x  -> t ;  b_{l-1} = P^T r
        P4B_CHECK(halo(m, l, L.x));
        {
            ProfScope ps(m, l, P4B_K_RESIDUAL);
            StencilOp op;
            memset(&op, 0, sizeof op);
            op.mode = ST_LIN; op.u = L.x; op.b = L.b; op.out = L.t; op.cb = 0.0; op.cg = 1.0;
            P4B_CHECK(make_port(m, l, L.t, &op.port));
            P4B_CHECK(launch_stencil(st, L.d, op, red));
        }
        P4B_CHECK(halo(m, l, L.t));
        const bool boundary = C.replicated && !L.replicated;
        const LevelDesc &Cd = boundary ? C.own : C.d;
        double *bc = boundary ? C.b + (size_t)C.own.zs * C.d.plane() : C.b;
        {
            ProfScope ps(m, l, P4B_K_RESTRICT);
            HaloPort port;
            if (!L.replicated && !C.replicated) {
                P4B_CHECK(make_port(m, l - 1, C.b, &port));      // the coarse right-hand side is exchanged next
            } else {
                P4B_CHECK(make_port(m, l, nullptr, &port));      // only waits for the ghost planes of t
                wrote(m, C.b);
            }
            P4B_CHECK(launch_restrict(st, L.d, Cd, L.t, bc, port));
        }
        if (boundary) P4B_CHECK(gather_replicated(m, l - 1, C.b));
    }
    const int cycles = (l == 1 || m->o.cycle == P4B_CYCLE_V) ? 1 : 2;
    for (int c = 0; c < cycles; c++) {
        if (l == m->top && c == 0 && graph_usable(m)) {
            ProfScope ps(m, l, P4B_K_SUBCYCLE);
            P4B_CHECK(coarse_cycle_graph(m));
        } else if (l == m->top) {
            ProfScope ps(m, l, P4B_K_SUBCYCLE);
            P4B_CHECK(cycle(m, l - 1, c == 0));
        } else {
            P4B_CHECK(cycle(m, l - 1, c == 0));
        }
    }
    P4B_CHECK(halo(m, l - 1, C.x));
    {
        ProfScope ps(m, l, P4B_K_PROLONG);
        HaloPort port;
        P4B_CHECK(make_port(m, l, L.x, &port));
        P4B_CHECK(launch_prolong_add(st, L.d, C.d, C.x, L.x, port));
    }
    // z = M^-1 r on the finest level feeds vector updates only: its ghost planes are not exchanged
    return smooth(m, l, false, (l == m->top && m->o.fuse) ? m->dot2_target : nullptr, l != m->top);
}

// z = M^-1 r with r already in lev[top].b ; result in lev[top].x.  With dot2 != NULL (and fused kernels on)
// the last smoother kernel also leaves (z,z), (z,r) there; m->dot2_fused says whether it did.
static int mg_apply_internal(p4b_mg *m, double *dot2 = nullptr) {
    m->dot2_fused = false;
    m->dot2_target = dot2;
    int rc = cycle(m, m->top, true);
    m->dot2_target = nullptr;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// setup
// ------------------------------------------------------------------------------------------------
static int build_coarse_inverse(p4b_mg *m) {
    const LevelDesc &L = m->lev[0].d;
    const long long n = L.nglobal();
    // too large for a dense inverse: only -pc_type mg needs it, and coarse_solve() says so when it is asked for
    // (-pc_type none / jacobi run on a one-level "hierarchy" of any size)
    if (n > 2400) { m->n0 = 0; return 0; }
    m->n0 = (int)n;
    std::vector<double> A((size_t)n * n, 0.0);
    auto bd = [&](int i, int j, int k) {
        return (L.ax && (i == 0 || i == L.nx - 1)) || (L.ay && (j == 0 || j == L.ny - 1)) ||
               (L.az && (k == 0 || k == L.nz - 1));
    };
    auto id = [&](int i, int j, int k) { return ((size_t)k * L.ny + j) * L.nx + i; };
    for (int k = 0; k < L.nz; k++)
        for (int j = 0; j < L.ny; j++)
            for (int i = 0; i < L.nx; i++) {
                const size_t p = id(i, j, k);
                A[p * n + p] = L.diag;
                if (bd(i, j, k)) continue;
                if (L.ax) {
                    if (!bd(i - 1, j, k)) A[p * n + id(i - 1, j, k)] = -L.cx;
                    if (!bd(i + 1, j, k)) A[p * n + id(i + 1, j, k)] = -L.cx;
                }
                if (L.ay) {
                    if (!bd(i, j - 1, k)) A[p * n + id(i, j - 1, k)] = -L.cy;
                    if (!bd(i, j + 1, k)) A[p * n + id(i, j + 1, k)] = -L.cy;
                }
                if (L.az) {
                    if (!bd(i, j, k - 1)) A[p * n + id(i, j, k - 1)] = -L.cz;
                    if (!bd(i, j, k + 1)) A[p * n + id(i, j, k + 1)] = -L.cz;
                }
            }
    // Cholesky A = G G^T (lower), then Ainv = G^-T G^-1   ([PETSc] PCLU on level 0: an exact solve)
    std::vector<double> G(A);
    for (long long c = 0; c < n; c++) {
        double d = G[c * n + c];
        for (long long q = 0; q < c; q++) d -= G[c * n + q] * G[c * n + q];
        if (!(d > 0)) return fail(61, "coarsest operator is not positive definite");
        d = sqrt(d);
        G[c * n + c] = d;
        for (long long r = c + 1; r < n; r++) {
            double s = G[r * n + c];
            for (long long q = 0; q < c; q++) s -= G[r * n + q] * G[c * n + q];
            G[r * n + c] = s / d;
        }
    }
    std::vector<double> Gi((size_t)n * n, 0.0);   // G^-1, lower triangular
    for (long long c = 0; c < n; c++) {
        Gi[c * n + c] = 1.0 / G[c * n + c];
        for (long long r = c + 1; r < n; r++) {
            double s = 0;
            for (long long q = c; q < r; q++) s -= G[r * n + q] * Gi[q * n + c];
            Gi[r * n + c] = s / G[r * n + r];
        }
    }
    std::vector<double> Ai((size_t)n * n, 0.0);
    for (long long r = 0; r < n; r++)
        for (long long c = 0; c <= r; c++) {
            double s = 0;
            for (long long q = r; q < n; q++) s += Gi[q * n + r] * Gi[q * n + c];
            Ai[r * n + c] = s;
            Ai[c * n + r] = s;
        }
    P4B_CUDA(cudaMalloc(&m->Ainv, sizeof(double) * n * n));
    P4B_CUDA(cudaMemcpy(m->Ainv, Ai.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    return 0;
}

static void slab_range(int m, int P, int r, int *s, int *c) {
    const int base = m / P, rem = m % P;
    *c = base + (r < rem ? 1 : 0);
    *s = r * base + (r < rem ? r : rem);
}

// ------------------------------------------------------------------------------------------------
// level + slab planning (host only; shared by p4b_mg_create and p4b_plan_levels)
// ------------------------------------------------------------------------------------------------
static long long g_rep_points = 70LL * 70 * 70;   // p4b_tune("rep_points", n)

struct LevelPlan {
    std::vector<p4b_grid> grids;                 // index 0 = coarsest
    std::vector<std::vector<int>> zs, zm;        // [level][rank]: planes of the slowest dimension each rank owns
    int lrep = -1;                               // levels <= lrep are replicated on every rank
};

static int plan_levels(const p4b_grid *g, const p4b_mg_opts &o, int P, LevelPlan *pl) {
    LevelDesc probe;
    P4B_CHECK(make_desc(g, &probe));
    std::vector<p4b_grid> gs;
    gs.push_back(*g);
    while ((o.levels <= 0 || (int)gs.size() < o.levels) && (int)gs.size() < P4B_MAX_LEVELS && can_coarsen(gs.back()))
        gs.push_back(coarsen(gs.back()));
    if (o.levels > 0 && (int)gs.size() < o.levels)
        return fail(60, "cannot build %d multigrid levels from this grid (got %d)", o.levels, (int)gs.size());
    const int nl = (int)gs.size();
    pl->grids.assign(gs.rbegin(), gs.rend());
    pl->zs.assign(nl, std::vector<int>(P, 0));
    pl->zm.assign(nl, std::vector<int>(P, 0));
    if (P > 1 && g->dim == 1) return fail(60, "1-D grids are not distributed");
    // finest by p4b_slab_range; coarser: coarse plane K belongs to the owner of fine plane 2K
    for (int l = nl - 1; l >= 0; l--) {
        const p4b_grid &gl = pl->grids[l];
        const int nz = gl.dim == 3 ? gl.mz : (gl.dim == 2 ? gl.my : 1);
        for (int r = 0; r < P; r++) {
            if (l == nl - 1) {
                slab_range(nz, P, r, &pl->zs[l][r], &pl->zm[l][r]);
            } else {
                const int fs = pl->zs[l + 1][r], fe = fs + pl->zm[l + 1][r];     // [fs, fe)
                const int ks = (fs + 1) / 2, ke = (fe - 1) / 2;                   // K with fs <= 2K <= fe-1
                pl->zs[l][r] = ks;
                pl->zm[l][r] = (pl->zm[l + 1][r] > 0 && ke >= ks) ? ke - ks + 1 : 0;
            }
        }
    }
    // replicate small levels: every rank must own >= 2 planes on a distributed level, and levels at or
    // below rep_points nodes are cheaper to compute redundantly than to exchange ghosts for
    const long long rep_points = g_rep_points;
    pl->lrep = -1;
    if (P > 1) {
        pl->lrep = 0;
        for (int l = 0; l < nl - 1; l++) {
            int minzm = 1 << 30;
            for (int r = 0; r < P; r++) minzm = std::min(minzm, pl->zm[l][r]);
            const p4b_grid &gl = pl->grids[l];
            if (minzm < 2 || (long long)gl.mx * gl.my * gl.mz <= rep_points) pl->lrep = l;
        }
        int minzm = 1 << 30;
        for (int r = 0; r < P; r++) minzm = std::min(minzm, pl->zm[nl - 1][r]);
        if (minzm < 2 || nl < 2) return fail(60, "grid too small to distribute over %d ranks", P);
    }
    return 0;
}

namespace p4b {
cudaStream_t ctx_stream(p4b_ctx *c) { return c->stream; }     // for the other translation units (nk_device.cu)
int ctx_rank(p4b_ctx *c) { return c->rank; }
int ctx_nranks(p4b_ctx *c) { return c->nranks; }
// y-slabs of a PERIODIC 2-D grid (pattern.c, DM_BOUNDARY_PERIODIC: c/ch5/pattern.c:79-84): ring exchange of one ghost row
// on each side ([PETSc] DMGlobalToLocal).  `owned` = first owned row, nrows rows of rowlen doubles; the ghost rows are
// owned - rowlen and owned + nrows * rowlen.  Grouped ncclSend / ncclRecv with the two ring neighbours (with two ranks
// both neighbours are the same peer: sends and receives pair up in issue order).
int ctx_ring_halo(p4b_ctx *c, double *owned, size_t rowlen, int nrows) {
    if (c->nranks == 1) {      // the rank is its own neighbour on both sides
        P4B_CUDA(cudaMemcpyAsync(owned - rowlen, owned + (size_t)(nrows - 1) * rowlen, sizeof(double) * rowlen,
                                 cudaMemcpyDeviceToDevice, c->stream));
        P4B_CUDA(cudaMemcpyAsync(owned + (size_t)nrows * rowlen, owned, sizeof(double) * rowlen, cudaMemcpyDeviceToDevice,
                                 c->stream));
        return 0;
    }
    const int prev = (c->rank + c->nranks - 1) % c->nranks, next = (c->rank + 1) % c->nranks;
    g_launch_count++;          // (one NCCL kernel per grouped exchange)
    P4B_NCCL(g_nccl.GroupStart());
    P4B_NCCL(g_nccl.Send(owned, rowlen, ncclFloat64, prev, c->comm, c->stream));                                   // my first row
    P4B_NCCL(g_nccl.Send(owned + (size_t)(nrows - 1) * rowlen, rowlen, ncclFloat64, next, c->comm, c->stream));   // my last row
    P4B_NCCL(g_nccl.Recv(owned + (size_t)nrows * rowlen, rowlen, ncclFloat64, next, c->comm, c->stream));         // next's first
    P4B_NCCL(g_nccl.Recv(owned - rowlen, rowlen, ncclFloat64, prev, c->comm, c->stream));                         // prev's last
    P4B_NCCL(g_nccl.GroupEnd());
    return 0;
}
// in-place all-gather: rank r's `count` doubles live at full + r * count
int ctx_allgather(p4b_ctx *c, double *full, size_t count) {
    if (c->nranks == 1) return 0;
    g_launch_count++;
    P4B_NCCL(g_nccl.AllGather(full + (size_t)c->rank * count, full, count, ncclFloat64, c->comm, c->stream));
    return 0;
}
extern long long g_recognise_residual, g_gmres_cgs;           // defined in nk_device.cu
}  // namespace p4b

extern "C" {

int p4b_version(void) { return P4B_VERSION; }
const char *p4b_last_error(void) { return g_err.c_str(); }
const char *p4b_kernel_name(int c) { return (c >= 0 && c < P4B_K_NCLASSES) ? k_names[c] : "?"; }

int p4b_tune(const char *key, long value) {
    if (!key) return fail(62, "null tuning key");
    if (tune_march(key, value) == 0) return 0;
    if (std::string(key) == "rep_points") { g_rep_points = value; return 0; }
    if (std::string(key) == "comm_peer") { g_comm_peer = value; return 0; }
    if (std::string(key) == "fused_halo") { g_fused_halo = value; return 0; }
    if (std::string(key) == "port_opts") { g_port_opts = value; return 0; }
    if (std::string(key) == "force_mg") { g_force_mg = value; return 0; }
    if (std::string(key) == "recognise_residual") { g_recognise_residual = value; return 0; }
    if (std::string(key) == "gmres_cgs") { g_gmres_cgs = value; return 0; }
    return fail(62, "unknown tuning key %s", key);
}

int p4b_device_count(int *n) {
    P4B_CUDA(cudaGetDeviceCount(n));
    return 0;
}

int p4b_ctx_create(int device, void *stream, p4b_ctx **out) {
    if (!out) return fail(62, "null ctx pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(70, "no CUDA device available (%s): p4b200 has no CPU fallback", cudaGetErrorString(e));
    P4B_CUDA(cudaSetDevice(device));
    {   // stream-ordered allocations (cudaMallocAsync in the Newton / time-stepping / FAS hosts): keep freed memory in the
        // device's pool instead of returning it to the driver at every synchronisation point (the default threshold is 0)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    p4b_ctx *c = new p4b_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;   // exactly the caller's stream; NULL is the (legacy) default stream
    c->red.max_blocks = 1 << 20;
    P4B_CUDA(cudaMalloc(&c->red.partials, sizeof(double) * 4 * (size_t)c->red.max_blocks));
    P4B_CUDA(cudaMalloc(&c->red.ticket, sizeof(unsigned int)));
    P4B_CUDA(cudaMemset(c->red.ticket, 0, sizeof(unsigned int)));
    P4B_CUDA(cudaMalloc(&c->d_scal, sizeof(double) * 16));
    P4B_CUDA(cudaMemset(c->d_scal, 0, sizeof(double) * 16));
    P4B_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * 16));
    P4B_CUDA(cudaHostAlloc((void **)&c->h_poll, sizeof(HostPoll), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c->h_poll, 0, sizeof(HostPoll));
    P4B_CUDA(cudaHostGetDevicePointer((void **)&c->d_poll, c->h_poll, 0));
    P4B_CUDA(cudaEventCreate(&c->ev0));
    P4B_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return 0;
}

// the same with a stream of its own (non-blocking, so the coarse levels can be captured into a CUDA graph): for C hosts
// that have no CUDA headers -- the PETSc-shaped shim creates one context per GPU this way (-p4b_gpus N)
int p4b_ctx_create_own_stream(int device, p4b_ctx **out) {
    if (!out) return fail(62, "null ctx pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(70, "no CUDA device available (%s): p4b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(62, "device %d of %d", device, ndev);
    P4B_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    P4B_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int rc = p4b_ctx_create(device, (void *)st, out);
    if (rc) { cudaStreamDestroy(st); return rc; }
    (*out)->own_stream = true;
    return 0;
}

int p4b_ctx_destroy(p4b_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->peer) {
        nccl_barrier(c);
        for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && c->peers.mbox[r] && c->mbox_ipc[r]) cudaIpcCloseMemHandle(c->peers.mbox[r]);
        nccl_barrier(c);
        cudaFree(c->mbox);
        cudaFree(c->sync);
        cudaFree(c->dummy_planes);
        cudaFree(c->dummy_flags);
    }
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaFree(c->red.partials);
    cudaFree(c->red.ticket);
    cudaFree(c->d_scal);
    cudaFreeHost(c->h_scal);
    if (c->d_mdot) cudaFree(c->d_mdot);
    if (c->h_poll) cudaFreeHost(c->h_poll);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int p4b_ctx_sync(p4b_ctx *c) {
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int p4b_comm_unique_id(void *id128) {
    P4B_CHECK(nccl_load());
    ncclUniqueId id;
    P4B_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return 0;
}

int p4b_comm_init(p4b_ctx *c, const void *id128, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(62, "bad rank %d of %d", rank, nranks);
    c->rank = rank;
    c->nranks = nranks;
    if (nranks == 1) return 0;
    P4B_CHECK(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    P4B_CUDA(cudaSetDevice(c->device));
    P4B_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    return peer_setup(c);
}

int p4b_comm_stats(p4b_ctx *c, unsigned long long out[5], int reset) {
    for (int i = 0; i < 5; i++) out[i] = 0;
    if (!c->sync) return 0;
    LocalSync h;
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    P4B_CUDA(cudaMemcpy(&h, c->sync, sizeof h, cudaMemcpyDeviceToHost));
    out[0] = h.wait_n; out[1] = h.wait_ns; out[2] = h.wait_max_ns; out[3] = h.fence_n; out[4] = h.fence_ns;
    if (reset) {
        const size_t off = offsetof(LocalSync, wait_ns);
        P4B_CUDA(cudaMemset((char *)c->sync + off, 0, sizeof(LocalSync) - off));
    }
    return 0;
}

int p4b_slab_range(int m, int nranks, int rank, int *start, int *count) {
    if (nranks < 1 || rank < 0 || rank >= nranks || m < 0) return fail(62, "bad slab arguments");
    slab_range(m, nranks, rank, start, count);
    return 0;
}

int p4b_malloc(p4b_ctx *c, size_t bytes, void **d) {
    P4B_CUDA(cudaSetDevice(c->device));
    P4B_CUDA(cudaMalloc(d, bytes ? bytes : 8));
    return 0;
}
int p4b_free(p4b_ctx *c, void *d) {
    P4B_CUDA(cudaSetDevice(c->device));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    P4B_CUDA(cudaFree(d));
    return 0;
}
int p4b_memcpy_h2d(p4b_ctx *c, void *dst, const void *src, size_t bytes) {
    P4B_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int p4b_memcpy_d2h(p4b_ctx *c, void *dst, const void *src, size_t bytes) {
    P4B_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- single-slab building blocks ------------------------------------------------------------------
int p4b_stencil_apply(p4b_ctx *c, const p4b_grid *g, const double *u, double *y) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    StencilOp op;
    memset(&op, 0, sizeof op);
    op.mode = ST_APPLY; op.u = u; op.out = y;
    return launch_stencil(c->stream, L, op, c->red);
}

int p4b_stencil_residual(p4b_ctx *c, const p4b_grid *g, const double *b, const double *u, double *r) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    StencilOp op;
    memset(&op, 0, sizeof op);
    op.mode = ST_LIN; op.u = u; op.b = b; op.out = r; op.cb = 0.0; op.cg = 1.0;
    return launch_stencil(c->stream, L, op, c->red);
}

int p4b_lambda_max_jacobi(const p4b_grid *g, double *lam) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    *lam = lambda_max(L);
    return 0;
}

static void cheb_omegas(double emin, double emax, int its, double *scale, std::vector<double> *omega) {
    *scale = 2.0 / (emax + emin);
    const double alpha = 1.0 - (*scale) * emin;
    const double mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
    double cm1 = 1.0, ck = mu;
    omega->clear();
    for (int i = 1; i < its; i++) {
        const double cp1 = 2.0 * mu * ck - cm1;
        omega->push_back(omegaprod * ck / cp1);
        cm1 = ck;
        ck = cp1;
    }
}

int p4b_cheb_jacobi(p4b_ctx *c, const p4b_grid *g, double emin, double emax, int its, int zero_guess,
                    const double *b, double *x, double *work) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    if (its <= 0) return 0;
    double scale;
    std::vector<double> om;
    cheb_omegas(emin, emax, its, &scale, &om);
    const double s1 = scale / L.diag;
    const long long n = L.nlocal();
    StencilOp op;
    memset(&op, 0, sizeof op);
    double *pm1 = x, *pk = work;
    bool pm1_zero = zero_guess != 0;
    if (zero_guess) {
        P4B_CHECK(launch_scale_copy(c->stream, n, s1, b, pk));
    } else {
        op.mode = ST_LIN; op.u = pm1; op.b = b; op.out = pk; op.cb = 1.0; op.cg = s1;
        P4B_CHECK(launch_stencil(c->stream, L, op, c->red));
    }
    for (int i = 1; i < its; i++) {
        const double w = om[i - 1];
        op.u = pk; op.b = b; op.out = pm1; op.cb = w; op.cg = w * s1;
        if (pm1_zero) {
            op.mode = ST_LIN;
        } else {
            op.mode = ST_LIN_PM1; op.pm1 = pm1; op.ca = 1.0 - w;
        }
        P4B_CHECK(launch_stencil(c->stream, L, op, c->red));
        std::swap(pm1, pk);
        pm1_zero = false;
    }
    if (pk != x) P4B_CUDA(cudaMemcpyAsync(x, pk, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int p4b_restrict(p4b_ctx *c, const p4b_grid *gf, const double *rf, double *bc) {
    if (!can_coarsen(*gf)) return fail(60, "grid cannot be coarsened");
    p4b_grid gc = coarsen(*gf);
    LevelDesc F, C;
    P4B_CHECK(make_desc(gf, &F));
    P4B_CHECK(make_desc(&gc, &C));
    return launch_restrict(c->stream, F, C, rf, bc);
}

int p4b_prolong_add(p4b_ctx *c, const p4b_grid *gf, const double *xc, double *xf) {
    if (!can_coarsen(*gf)) return fail(60, "grid cannot be coarsened");
    p4b_grid gc = coarsen(*gf);
    LevelDesc F, C;
    P4B_CHECK(make_desc(gf, &F));
    P4B_CHECK(make_desc(&gc, &C));
    return launch_prolong_add(c->stream, F, C, xc, xf);
}

int p4b_residual_restrict(p4b_ctx *c, const p4b_grid *gf, const double *b, const double *x, double *bc) {
    // a composition over a scratch vector from the async pool (no fused kernel exists: DESIGN.md section 4)
    LevelDesc F;
    P4B_CHECK(make_desc(gf, &F));
    double *t = nullptr;
    P4B_CUDA(cudaMallocAsync((void **)&t, sizeof(double) * F.nlocal(), c->stream));
    int rc = p4b_stencil_residual(c, gf, b, x, t);
    if (!rc) rc = p4b_restrict(c, gf, t, bc);
    cudaFreeAsync(t, c->stream);
    return rc;
}

int p4b_vec_dot(p4b_ctx *c, size_t n, const double *x, const double *y, double *res) {
    P4B_CHECK(launch_dotn(c->stream, (long long)n, x, y, c->d_scal + 8, c->red));
    return allreduce_fetch(c, c->d_scal + 8, 1, 0, res);
}
// k dot products (X[i], y) with ONE read-back: the launches of the single dot, one after the other on the stream, each
// into its own slot ([PETSc] VecMDot; the orthogonalisation of one GMRES step with classical Gram-Schmidt)
int p4b_vec_mdot(p4b_ctx *c, size_t n, int k, const double *const *X, const double *y, double *res) {
    if (k <= 0) return 0;
    if (k > 64) return fail(62, "p4b_vec_mdot: at most 64 vectors");
    if (!c->d_mdot) P4B_CUDA(cudaMalloc(&c->d_mdot, sizeof(double) * 64));
    for (int i = 0; i < k; i++) P4B_CHECK(launch_dotn(c->stream, (long long)n, X[i], y, c->d_mdot + i, c->red));
    return allreduce_fetch(c, c->d_mdot, k, 0, res);
}
int p4b_vec_wrms2(p4b_ctx *c, size_t n, const double *x, const double *y, double atol, double rtol, double *res) {
    P4B_CHECK(launch_wrms(c->stream, (long long)n, x, y, atol, rtol, c->d_scal + 8, c->red));
    return allreduce_fetch(c, c->d_scal + 8, 1, 0, res);
}
int p4b_vec_norm2(p4b_ctx *c, size_t n, const double *x, double *res) {
    P4B_CHECK(p4b_vec_dot(c, n, x, x, res));
    *res = sqrt(*res);
    return 0;
}
int p4b_vec_norminf(p4b_ctx *c, size_t n, const double *x, double *res) {
    P4B_CHECK(launch_absmax(c->stream, (long long)n, x, c->d_scal + 8, c->red));
    return allreduce_fetch(c, c->d_scal + 8, 1, 1, res);
}
int p4b_vec_axpy(p4b_ctx *c, size_t n, double a, const double *x, double *y) {
    return launch_axpy(c->stream, (long long)n, a, x, y);
}
int p4b_vec_aypx(p4b_ctx *c, size_t n, double a, const double *x, double *y) {
    return launch_aypx(c->stream, (long long)n, a, x, y);
}
int p4b_vec_set(p4b_ctx *c, size_t n, double a, double *y) { return launch_set(c->stream, (long long)n, a, y); }

// ---- fish problem -----------------------------------------------------------------------------------
int p4b_fish_sample(p4b_ctx *c, const p4b_grid *g, int problem, double *f, double *gb) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    if (problem == P4B_PROBLEM_MANUEXP && (g->cx != 1.0 || g->cy != 1.0 || g->cz != 1.0))
        return fail(3, "cx=cy=cz=1 required for problem MANUEXP");   // fish.c:191-193
    return launch_fish_sample(c->stream, L, g->dim, problem, g->cx, g->cy, g->cz, f, gb);
}
int p4b_initial_state(p4b_ctx *c, const p4b_grid *g, const double *gb, int gonboundary, double *u) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    if (gonboundary && !gb) return fail(62, "gb required when gonboundary is set");
    return launch_initial_state(c->stream, L, gb, gonboundary, u);
}
int p4b_poisson_function(p4b_ctx *c, const p4b_grid *g, const double *u, const double *f, const double *gb,
                         double *F) {
    LevelDesc L;
    P4B_CHECK(make_desc(g, &L));
    return launch_poisson_function(c->stream, L, g->dim, g->cx, u, f, gb, F);
}

// ---- PCMG -------------------------------------------------------------------------------------------
int p4b_mg_default_opts(p4b_mg_opts *o) {
    memset(o, 0, sizeof *o);
    o->levels = 0;
    o->cycle = P4B_CYCLE_V;
    o->smoother = P4B_SMOOTH_CHEBYSHEV;
    o->smooth_its = 2;
    o->emin = 0; o->emax = 0;
    o->est_lo = 0.1; o->est_hi = 1.1;
    o->fuse = 1;
    o->use_graph = 1;
    return 0;
}

static int mg_create_impl(p4b_ctx *c, const p4b_grid *g, const p4b_mg_opts *oin, const double *coef, int ncoef,
                          p4b_mg **out);

int p4b_mg_create(p4b_ctx *c, const p4b_grid *g, const p4b_mg_opts *oin, p4b_mg **out) {
    return mg_create_impl(c, g, oin, nullptr, 0, out);
}

// Same hierarchy, but the operator of every level comes from the caller: coef[4*l + {0,1,2,3}] =
// (diag, off_x, off_y, off_z) magnitudes of level l, FINEST FIRST (l = 0), as a Mat plugin reads them out of
// the values the user's FormJacobianLocal inserted on that level's DMDA (rediscretisation, fish.c:7).
int p4b_mg_create_stencil(p4b_ctx *c, const p4b_grid *g, const p4b_mg_opts *oin, const double *coef, int nlevels_coef,
                          p4b_mg **out) {
    if (!coef || nlevels_coef < 1) return fail(62, "stencil coefficients required");
    return mg_create_impl(c, g, oin, coef, nlevels_coef, out);
}

static int mg_create_impl(p4b_ctx *c, const p4b_grid *g, const p4b_mg_opts *oin, const double *coef, int ncoef,
                          p4b_mg **out) {
    p4b_mg_opts o;
    if (oin) o = *oin; else p4b_mg_default_opts(&o);
    if (o.cycle != P4B_CYCLE_V && o.cycle != P4B_CYCLE_W) return fail(62, "unknown -pc_mg_cycle_type");
    if (o.smoother != P4B_SMOOTH_CHEBYSHEV && o.smoother != P4B_SMOOTH_RICHARDSON)
        return fail(62, "smoother must be chebyshev or richardson (Jacobi PC); SOR is sequential and not provided");
    P4B_CUDA(cudaSetDevice(c->device));
    p4b_mg *m = new p4b_mg();
    m->ctx = c;
    m->o = o;
    LevelPlan plan;
    const int P = c->nranks, R = c->rank;
    {
        int rc = plan_levels(g, o, P, &plan);
        if (rc) { delete m; return rc; }
    }
    const int nl = (int)plan.grids.size();
    m->lev.resize(nl);
    m->top = nl - 1;
    for (int l = 0; l < nl; l++) {
        Level &L = m->lev[l];
        L.g = plan.grids[l];
        int rc = make_desc(&L.g, &L.d);
        if (rc) { delete m; return rc; }
        if (coef) {
            const int lc = nl - 1 - l;      // caller's index: finest first
            if (lc >= ncoef) { delete m; return fail(62, "coefficients for %d levels given, %d needed", ncoef, nl); }
            const double *q = coef + 4 * lc;
            if (!(q[0] > 0)) { delete m; return fail(62, "non-positive diagonal on level %d", lc); }
            L.d.diag = q[0];
            if (L.g.dim == 1) { L.d.cx = q[1]; }
            else if (L.g.dim == 2) { L.d.cx = q[1]; L.d.cz = q[2]; }     // 2-D: y lives in slot z
            else { L.d.cx = q[1]; L.d.cy = q[2]; L.d.cz = q[3]; }
        }
        L.lam = lambda_max(L.d);
        if (o.emax > 0) { L.emin = o.emin; L.emax = o.emax; }
        else { L.emin = o.est_lo * L.lam; L.emax = o.est_hi * L.lam; }
        cheb_omegas(L.emin, L.emax, o.smooth_its, &L.scale, &L.omega);
        L.zs_all = plan.zs[l];
        L.zm_all = plan.zm[l];
        L.own = L.d;
        L.own.zs = L.zs_all[R];
        L.own.zm = L.zm_all[R];
        L.replicated = (P > 1 && l <= plan.lrep);
        if (P > 1 && !L.replicated) L.d = L.own;
    }
    // arena: ghosted vectors x, b, t per level (+ p, w and two scratch vectors on the finest).  The layout is a
    // pure function of the plan, so a rank can compute where its neighbours keep the same buffers.
    auto layout = [&](int rank, std::vector<std::array<size_t, 7>> *offs) {
        size_t off = 0;
        offs->assign(nl, std::array<size_t, 7>{});
        for (int l = 0; l < nl; l++) {
            const Level &L = m->lev[l];
            const size_t plane = (size_t)L.d.plane();
            const int zm = (P > 1 && !L.replicated) ? L.zm_all[rank] : L.d.nz;
            size_t n = plane * (size_t)(zm + 2) + 8;
            n = (n + 31) & ~(size_t)31;
            const size_t lead = 4 + (plane & 1);      // owned start 16-byte aligned (arena base is 256-byte aligned)
            const int ns = (l == nl - 1) ? 7 : 3;
            for (int q = 0; q < ns; q++) {
                (*offs)[l][q] = off + lead + plane;
                off += n;
            }
        }
        return off;
    };
    std::vector<std::array<size_t, 7>> offs_me, offs_prev, offs_next;
    const size_t total = layout(R, &offs_me);
    if (P > 1 && R > 0) layout(R - 1, &offs_prev);
    if (P > 1 && R < P - 1) layout(R + 1, &offs_next);
    m->arena_doubles = total;
    cudaError_t e = cudaMalloc(&m->arena, sizeof(double) * total);
    if (e != cudaSuccess) {
        delete m;
        return fail(71, "cudaMalloc of %.2f GB for the level hierarchy failed: %s", total * 8e-9, cudaGetErrorString(e));
    }
    P4B_CUDA(cudaMemsetAsync(m->arena, 0, sizeof(double) * total, c->stream));
    for (int l = 0; l < nl; l++) {
        Level &L = m->lev[l];
        L.nslots = (l == nl - 1) ? 7 : 3;
        L.off_me = offs_me[l];
        if (P > 1 && R > 0) { L.off_prev = offs_prev[l]; L.zm_prev = L.zm_all[R - 1]; }
        if (P > 1 && R < P - 1) { L.off_next = offs_next[l]; L.zm_next = L.zm_all[R + 1]; }
        L.x = m->arena + L.off_me[0];
        L.b = m->arena + L.off_me[1];
        L.t = m->arena + L.off_me[2];
        if (l == nl - 1) {
            m->p = m->arena + L.off_me[3];
            m->w = m->arena + L.off_me[4];
            m->fbuf = m->arena + L.off_me[5];
            m->gbuf = m->arena + L.off_me[6];
        }
    }
    if (c->peer) {
        // map every rank's arena (neighbours for ghost planes, everyone for the replicated-level gather)
        P4B_CUDA(cudaStreamSynchronize(c->stream));
        void *mapped[MAX_RANKS];
        P4B_CHECK(map_peers(c, m->arena, mapped, m->arena_ipc));
        for (int r = 0; r < P; r++) m->peer_arena.base[r] = (double *)mapped[r];
        P4B_CHECK(nccl_barrier(c));
        m->peer = true;
    }
    int rc = build_coarse_inverse(m);
    if (rc) return rc;
    memset(m->prof.stat, 0, sizeof m->prof.stat);
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    *out = m;
    return 0;
}

int p4b_plan_levels(const p4b_grid *g, const p4b_mg_opts *oin, int nranks, int *nlevels, int *m3, int *zs, int *zm,
                    int *replicated) {
    p4b_mg_opts o;
    if (oin) o = *oin; else p4b_mg_default_opts(&o);
    if (nranks < 1) return fail(62, "bad nranks");
    LevelPlan pl;
    P4B_CHECK(plan_levels(g, o, nranks, &pl));
    const int nl = (int)pl.grids.size();
    *nlevels = nl;
    for (int l = 0; l < nl; l++) {
        const p4b_grid &gl = pl.grids[l];
        if (m3) { m3[3 * l] = gl.mx; m3[3 * l + 1] = gl.dim >= 2 ? gl.my : 1; m3[3 * l + 2] = gl.dim >= 3 ? gl.mz : 1; }
        for (int r = 0; r < nranks; r++) {
            if (zs) zs[l * nranks + r] = pl.zs[l][r];
            if (zm) zm[l * nranks + r] = pl.zm[l][r];
        }
        if (replicated) replicated[l] = (nranks > 1 && l <= pl.lrep) ? 1 : 0;
    }
    return 0;
}

int p4b_mg_destroy(p4b_mg *m) {
    if (!m) return 0;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (auto e : m->prof.pool) cudaEventDestroy(e);
    if (m->coarse_graph) cudaGraphExecDestroy(m->coarse_graph);
    if (m->peer) {      // collective: nobody may still be storing into an arena that is about to be freed
        nccl_barrier(m->ctx);
        for (int r = 0; r < m->ctx->nranks; r++)
            if (r != m->ctx->rank && m->peer_arena.base[r] && m->arena_ipc[r]) cudaIpcCloseMemHandle(m->peer_arena.base[r]);
        nccl_barrier(m->ctx);
    }
    cudaFree(m->arena);
    cudaFree(m->Ainv);
    delete m;
    return 0;
}

int p4b_mg_nlevels(p4b_mg *m, int *n) { *n = (int)m->lev.size(); return 0; }

int p4b_mg_level_info(p4b_mg *m, int l, int *mm, double *eig) {
    if (l < 0 || l >= (int)m->lev.size()) return fail(62, "level out of range");
    const Level &L = m->lev[l];
    if (mm) { mm[0] = L.g.mx; mm[1] = L.g.dim >= 2 ? L.g.my : 1; mm[2] = L.g.dim >= 3 ? L.g.mz : 1; }
    if (eig) { eig[0] = L.emin; eig[1] = L.emax; }
    return 0;
}

int p4b_mg_local_range(p4b_mg *m, int *start, int *count, size_t *nlocal) {
    const LevelDesc &d = m->lev[m->top].d;
    if (start) *start = d.zs;
    if (count) *count = d.zm;
    if (nlocal) *nlocal = (size_t)d.nlocal();
    return 0;
}

// y = A x on the finest level (MatMult of the level operator, explicit coefficients honoured)
int p4b_mg_matmult(p4b_mg *m, const double *x, double *y) {
    Level &T = m->lev[m->top];
    cudaStream_t st = m->ctx->stream;
    const size_t bytes = sizeof(double) * (size_t)T.d.nlocal();
    m->pushed = nullptr;
    P4B_CUDA(cudaMemcpyAsync(m->p, x, bytes, cudaMemcpyDeviceToDevice, st));
    P4B_CHECK(halo(m, m->top, m->p));
    StencilOp op;
    memset(&op, 0, sizeof op);
    op.mode = ST_APPLY; op.u = m->p; op.out = m->w;
    P4B_CHECK(launch_stencil(st, T.d, op, m->ctx->red));
    P4B_CUDA(cudaMemcpyAsync(y, m->w, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int p4b_mg_apply(p4b_mg *m, const double *r, double *z) {
    Level &T = m->lev[m->top];
    cudaStream_t st = m->ctx->stream;
    const size_t bytes = sizeof(double) * (size_t)T.d.nlocal();
    m->pushed = nullptr;
    P4B_CUDA(cudaMemcpyAsync(T.b, r, bytes, cudaMemcpyDeviceToDevice, st));
    P4B_CHECK(mg_apply_internal(m));
    P4B_CUDA(cudaMemcpyAsync(z, T.x, bytes, cudaMemcpyDeviceToDevice, st));
    prof_collect(m);
    return 0;
}

// [PETSc] KSPSolve_CG.  r lives in lev[top].b, z in lev[top].x (so PCApply is in place).
int p4b_cg_solve(p4b_mg *m, int pc_type, const double *b, double *x, double rtol, double abstol, int max_it,
                 p4b_ksp_result *res) {
    p4b_ctx *c = m->ctx;
    cudaStream_t st = c->stream;
    Level &T = m->lev[m->top];
    const long long n = T.d.nlocal();
    const Reducer &red = c->red;
    double *S = c->d_scal;   // [0,1]=(zz,zr) parity 0 ; [2,3] parity 1 ; [4] = (p,w)
    p4b_ksp_result R;
    memset(&R, 0, sizeof R);
    if (pc_type != P4B_PC_NONE && pc_type != P4B_PC_JACOBI && pc_type != P4B_PC_MG)
        return fail(62, "unknown pc_type %d (the device path provides none, jacobi, mg)", pc_type);
    P4B_CUDA(cudaSetDevice(c->device));
    P4B_CUDA(cudaEventRecord(c->ev0, st));
    m->pushed = nullptr;
    P4B_CUDA(cudaMemcpyAsync(T.b, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    P4B_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, st));
    // z = M^-1 r and the CG scalars (z,z), (z,r) -> dots
    auto precond = [&](double *dots) -> int {
        m->dot2_fused = false;
        if (pc_type == P4B_PC_MG) P4B_CHECK(mg_apply_internal(m, dots));
        else P4B_CHECK(launch_scale_copy(st, n, pc_type == P4B_PC_JACOBI ? 1.0 / T.d.diag : 1.0, T.b, T.x));
        if (!m->dot2_fused) {
            ProfScope ps(m, m->top, P4B_K_DOT2);
            P4B_CHECK(launch_dot2(st, n, T.x, T.b, dots, red));
        }
        return 0;
    };
    int q = 0;
    double h[2];
    P4B_CHECK(precond(S + 2 * q));
    // (z,z), (z,r): all-reduced and published to pinned host memory by one kernel; the host polls (no stream sync)
    // (the iteration's (p, A p), already all-reduced in S[4], rides along as a third value: KSPSolve_CG's indefinite-matrix test)
    double h3[3] = {0.0, 0.0, 1.0};
    auto reduce_fetch2 = [&](double *dots) -> int {
        const unsigned long long seq = ++c->poll_seq;
        {
            ProfScope ps(m, m->top, P4B_K_ALLREDUCE);
            if (c->nranks > 1 && c->peer) {
                P4B_CHECK(launch_allreduce(st, dots, 2, 0, c->peers, c->sync, c->d_poll, seq, S + 4));
            } else {
                P4B_CHECK(ctx_allreduce(c, dots, 2));
                P4B_CHECK(launch_publish(st, dots, 2, c->d_poll, seq, S + 4));
            }
        }
        const int rc = poll_wait(c, seq, 3, h3);
        h[0] = h3[0]; h[1] = h3[1];
        prof_break(m);
        return rc;
    };
    P4B_CHECK(reduce_fetch2(S + 2 * q));
    double dp = sqrt(h[0]);
    R.rnorm0 = dp;
    R.hist[R.nhist++] = dp;
    const double ttol = fmax(rtol * dp, abstol);
    int its = 0;
    R.reason = 0;
    if (!(dp == dp)) R.reason = P4B_DIVERGED_NAN;
    else if (dp <= ttol) R.reason = (dp <= abstol) ? P4B_CONVERGED_ATOL : P4B_CONVERGED_RTOL;
    else if (h[1] <= 0.0) R.reason = P4B_DIVERGED_INDEFINITE_PC;       // [PETSc] KSPSolve_CG: beta = (z, r) must be positive
    while (!R.reason) {
        if (its >= max_it) { R.reason = P4B_DIVERGED_ITS; break; }
        if (m->o.fuse) {
            // x += alpha_{k-1} p (deferred from the previous iteration) and p = z + beta p in one pass over p
            ProfScope ps(m, m->top, P4B_K_XP_UPDATE);
            HaloPort port;
            P4B_CHECK(make_port(m, m->top, m->p, &port));
            P4B_CHECK(launch_xp_update(st, n, S + 2 * (1 - q) + 1, S + 4, S + 2 * q + 1, S + 2 * (1 - q) + 1, T.x, m->p, x,
                                       its == 0, port));
        } else {
            ProfScope ps(m, m->top, P4B_K_AYPX);
            wrote(m, m->p);
            P4B_CHECK(launch_aypx_dev(st, n, S + 2 * q + 1, S + 2 * (1 - q) + 1, T.x, m->p, its == 0));
        }
        P4B_CHECK(halo(m, m->top, m->p));
        {
            ProfScope ps(m, m->top, P4B_K_APPLY_DOT);
            StencilOp op;
            memset(&op, 0, sizeof op);
            op.mode = ST_APPLY_DOT; op.u = m->p; op.out = m->w; op.dot_out = S + 4;
            P4B_CHECK(make_port(m, m->top, nullptr, &op.port));
            wrote(m, m->w);
            P4B_CHECK(launch_stencil(st, T.d, op, red));
        }
        {
            ProfScope ps(m, m->top, P4B_K_ALLREDUCE);
            P4B_CHECK(ctx_allreduce(c, S + 4, 1));
        }
        if (m->o.fuse) {
            ProfScope ps(m, m->top, P4B_K_R_UPDATE);
            HaloPort port;
            P4B_CHECK(make_port(m, m->top, T.b, &port));
            P4B_CHECK(launch_r_update(st, n, S + 2 * q + 1, S + 4, m->w, T.b, port));
        } else {
            ProfScope ps(m, m->top, P4B_K_AXPY2);
            wrote(m, T.b);
            P4B_CHECK(launch_axpy2(st, n, S + 2 * q + 1, S + 4, m->p, m->w, x, T.b));
        }
        q ^= 1;
        P4B_CHECK(precond(S + 2 * q));
        P4B_CHECK(reduce_fetch2(S + 2 * q));
        dp = sqrt(h[0]);
        its++;
        if (R.nhist < P4B_MAX_HIST) R.hist[R.nhist++] = dp;
        const double pw = h3[2];                        // (p, A p) of this iteration
        // [PETSc] KSPSolve_CG's order of tests: indefinite matrix, NaN, convergence / divergence tolerance, indefinite PC
        if (pw <= 0.0) R.reason = P4B_DIVERGED_INDEFINITE_MAT;
        else if (!(dp == dp)) R.reason = P4B_DIVERGED_NAN;
        else if (dp <= ttol) R.reason = (dp <= abstol) ? P4B_CONVERGED_ATOL : P4B_CONVERGED_RTOL;
        else if (dp >= 1.0e5 * R.rnorm0) R.reason = P4B_DIVERGED_DTOL;
        else if (h[1] <= 0.0) R.reason = P4B_DIVERGED_INDEFINITE_PC;
    }
    if (m->o.fuse && its >= 1)      // the last iteration's x += alpha p is still pending
        P4B_CHECK(launch_x_flush(st, n, S + 2 * (1 - q) + 1, S + 4, m->p, x));
    P4B_CUDA(cudaEventRecord(c->ev1, st));
    P4B_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0;
    P4B_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    R.its = its;
    R.rnorm = dp;
    R.solve_ms = ms;
    prof_collect(m);
    if (res) *res = R;
    return 0;
}

int p4b_cg_solve_host(p4b_mg *m, int pc_type, const double *bh, double *xh, double rtol, double abstol, int max_it,
                      p4b_ksp_result *res) {
    p4b_ctx *c = m->ctx;
    Level &T = m->lev[m->top];
    const size_t bytes = sizeof(double) * (size_t)T.d.nlocal();
    P4B_CUDA(cudaMemcpyAsync(m->fbuf, bh, bytes, cudaMemcpyHostToDevice, c->stream));
    P4B_CHECK(p4b_cg_solve(m, pc_type, m->fbuf, m->gbuf, rtol, abstol, max_it, res));
    P4B_CUDA(cudaMemcpyAsync(xh, m->gbuf, bytes, cudaMemcpyDeviceToHost, c->stream));
    P4B_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// [PETSc] SNESSolve_KSPONLY: F0 = F(u0); J y = F0; u = u0 - y.
int p4b_fish_solve_host(p4b_mg *m, const double *fh, const double *gh, double *uh, double rtol, double abstol,
                        int max_it, p4b_ksp_result *res) {
    p4b_ctx *c = m->ctx;
    cudaStream_t st = c->stream;
    Level &T = m->lev[m->top];
    const long long n = T.d.nlocal();
    const size_t bytes = sizeof(double) * (size_t)n;
    // f -> fbuf, g -> gbuf, u -> w (ghosted; halo for F(u)), F -> p, y -> t... keep x,b,t for the solver
    P4B_CUDA(cudaMemcpyAsync(m->fbuf, fh, bytes, cudaMemcpyHostToDevice, st));
    P4B_CUDA(cudaMemcpyAsync(m->gbuf, gh, bytes, cudaMemcpyHostToDevice, st));
    P4B_CUDA(cudaMemcpyAsync(m->w, uh, bytes, cudaMemcpyHostToDevice, st));
    m->pushed = nullptr;
    P4B_CHECK(halo(m, m->top, m->w));
    P4B_CHECK(halo(m, m->top, m->gbuf));
    P4B_CHECK(launch_poisson_function(st, T.d, T.g.dim, T.g.cx, m->w, m->fbuf, m->gbuf, m->p));
    // F0 must survive the solve (p and w are CG work vectors): move F0 to fbuf, u0 to gbuf
    P4B_CUDA(cudaMemcpyAsync(m->fbuf, m->p, bytes, cudaMemcpyDeviceToDevice, st));
    P4B_CUDA(cudaMemcpyAsync(m->gbuf, m->w, bytes, cudaMemcpyDeviceToDevice, st));
    double *y = nullptr;
    P4B_CUDA(cudaMallocAsync((void **)&y, bytes, st));
    int rc = p4b_cg_solve(m, P4B_PC_MG, m->fbuf, y, rtol, abstol, max_it, res);
    if (!rc) rc = launch_axpy(st, n, -1.0, y, m->gbuf);
    cudaFreeAsync(y, st);
    if (rc) return rc;
    P4B_CUDA(cudaMemcpyAsync(uh, m->gbuf, bytes, cudaMemcpyDeviceToHost, st));
    P4B_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// b = F(u0) for fish.c's built-in problems on this rank's slab (any of the outputs may be NULL):
// fish.c:186-187 (function tables), :237-238 (InitialState), SNESComputeFunction at :239.
int p4b_mg_fish_setup(p4b_mg *m, int problem, int gonboundary, double *b_out, double *u0_out, double *uexact_out) {
    p4b_ctx *c = m->ctx;
    cudaStream_t st = c->stream;
    Level &T = m->lev[m->top];
    const size_t bytes = sizeof(double) * (size_t)T.d.nlocal();
    if (problem == P4B_PROBLEM_MANUEXP && (T.g.cx != 1.0 || T.g.cy != 1.0 || T.g.cz != 1.0))
        return fail(3, "cx=cy=cz=1 required for problem MANUEXP");
    P4B_CHECK(launch_fish_sample(st, T.d, T.g.dim, problem, T.g.cx, T.g.cy, T.g.cz, m->fbuf, m->gbuf));
    P4B_CHECK(launch_initial_state(st, T.d, m->gbuf, gonboundary, m->w));
    m->pushed = nullptr;
    P4B_CHECK(halo(m, m->top, m->w));
    P4B_CHECK(halo(m, m->top, m->gbuf));
    P4B_CHECK(launch_poisson_function(st, T.d, T.g.dim, T.g.cx, m->w, m->fbuf, m->gbuf, m->p));
    if (b_out) P4B_CUDA(cudaMemcpyAsync(b_out, m->p, bytes, cudaMemcpyDeviceToDevice, st));
    if (u0_out) P4B_CUDA(cudaMemcpyAsync(u0_out, m->w, bytes, cudaMemcpyDeviceToDevice, st));
    if (uexact_out) P4B_CUDA(cudaMemcpyAsync(uexact_out, m->gbuf, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- minimal.c / pattern.c callbacks, assembled SpMV ---------------------------------------------------
int p4b_minimal_sample(p4b_ctx *c, int mx, int my, int problem, double tent_H, double catenoid_c, double *g) {
    if (mx < 3 || my < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    if (problem != 0 && problem != 1) return fail(62, "minimal problem must be 0 (tent) or 1 (catenoid)");
    if (problem == 1 && catenoid_c < 1.0) return fail(62, "catenoid_c >= 1 required (minimal.c:116)");
    return launch_minimal_sample(c->stream, mx, my, 0, my, problem, tent_H, catenoid_c, g);
}
int p4b_minimal_function(p4b_ctx *c, int mx, int my, double q, const double *u, const double *g, double *FF) {
    if (mx < 3 || my < 3) return fail(60, "grid needs at least 3 nodes per dimension");
    return launch_minimal_function(c->stream, mx, my, 0, my, q, u, g, FF);
}
int p4b_pattern_initial_state(p4b_ctx *c, int mx, int my, double L, double *Y) {
    if (mx < 3 || my < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    return launch_pattern_init(c->stream, mx, my, L, Y);
}
int p4b_pattern_initial_state_noisy(p4b_ctx *c, int mx, int my, double L, const double *noise, double level, double *Y) {
    if (mx < 3 || my < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    return launch_pattern_init(c->stream, mx, my, L, Y, noise, level);
}
// [PETSc] rander48 (the default PetscRandom type): drand48's 48-bit linear congruential generator, see p4b200.h
unsigned long long p4b_rander48_seed(unsigned long seed) {
    return (((unsigned long long)(seed & 0xffffffffUL)) << 16) | 0x330EULL;
}
int p4b_rander48_fill(unsigned long long *state, size_t n, double *out) {
    if (!state || (n && !out)) return fail(62, "null argument");
    unsigned long long x = *state & 0xFFFFFFFFFFFFULL;
    for (size_t i = 0; i < n; i++) {
        x = (0x5DEECE66DULL * x + 0xBULL) & 0xFFFFFFFFFFFFULL;
        out[i] = (double)x * (1.0 / 281474976710656.0);      // X / 2^48: exact in fp64
    }
    *state = x;
    return 0;
}
int p4b_pattern_rhsfunction(p4b_ctx *c, int mx, int my, double phi, double kappa, const double *Y, double *G) {
    return launch_pattern_rhs(c->stream, mx * my, phi, kappa, Y, G);
}
int p4b_pattern_ifunction(p4b_ctx *c, int mx, int my, double L, double Du, double Dv, const double *Y, const double *Ydot,
                          double *F) {
    if (mx < 3 || my < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    const double h = L / (double)mx;                       // pattern.c:246
    return launch_pattern_ifunction(c->stream, mx, my, Du / (6.0 * h * h), Dv / (6.0 * h * h), 0, 0.0, Y, Ydot, F);
}
int p4b_pattern_ijacobian_mult(p4b_ctx *c, int mx, int my, double L, double Du, double Dv, double shift, const double *X,
                               double *JX) {
    if (mx < 3 || my < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    const double h = L / (double)mx;
    return launch_pattern_ifunction(c->stream, mx, my, Du / (6.0 * h * h), Dv / (6.0 * h * h), 1, shift, X, nullptr, JX);
}

static int pattern_coef(int mx, int my, double L, double Du, double Dv, double *Cu, double *Cv) {
    if (mx < 3 || my < 3) return fail(60, "periodic grid needs at least 3 nodes per dimension");
    const double h = L / (double)mx;
    *Cu = Du / (6.0 * h * h);
    *Cv = Dv / (6.0 * h * h);
    return 0;
}
int p4b_pattern_jac_apply(p4b_ctx *c, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                          double shift, const double *Y, const double *X, double *out) {
    double Cu, Cv;
    P4B_CHECK(pattern_coef(mx, my, L, Du, Dv, &Cu, &Cv));
    return launch_pattern_jac(c->stream, 0, mx, my, Cu, Cv, shift, phi, kappa, Y, X, nullptr, nullptr, 0, 0, 0, 0, out);
}
int p4b_pattern_jac_lin(p4b_ctx *c, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                        double shift, const double *Y, const double *X, const double *b, const double *pm1, double ca,
                        double cb, double cg, int jacobi, double *out) {
    double Cu, Cv;
    P4B_CHECK(pattern_coef(mx, my, L, Du, Dv, &Cu, &Cv));
    return launch_pattern_jac(c->stream, 1, mx, my, Cu, Cv, shift, phi, kappa, Y, X, b, pm1, ca, cb, cg, jacobi, out);
}
int p4b_pattern_jac_gershgorin(p4b_ctx *c, int mx, int my, double L, double Du, double Dv, double phi, double kappa,
                               double shift, const double *Y, double *work, double *res) {
    double Cu, Cv;
    P4B_CHECK(pattern_coef(mx, my, L, Du, Dv, &Cu, &Cv));
    P4B_CHECK(launch_pattern_jac(c->stream, 2, mx, my, Cu, Cv, shift, phi, kappa, Y, Y, nullptr, nullptr, 0, 0, 0, 0, work));
    return p4b_vec_norminf(c, (size_t)2 * mx * my, work, res);
}
int p4b_pattern_restrict(p4b_ctx *c, int Mx, int My, const double *rf, double *bc) {
    return launch_pattern_transfer(c->stream, 0, Mx, My, rf, bc);
}
int p4b_pattern_prolong_add(p4b_ctx *c, int Mx, int My, const double *xc, double *xf) {
    return launch_pattern_transfer(c->stream, 1, Mx, My, xc, xf);
}
int p4b_pattern_inject(p4b_ctx *c, int Mx, int My, const double *yf, double *yc) {
    return launch_pattern_transfer(c->stream, 2, Mx, My, yf, yc);
}

int p4b_minimal_jacobian_fd(p4b_ctx *c, int mx, int my, double q, const double *u, const double *g, const double *F0,
                            double *vals9) {
    if (mx < 3 || my < 3) return fail(60, "minimal Jacobian: grid must be at least 3 x 3");
    double *tmp = nullptr, unorm = 0.0;
    const size_t N = (size_t)mx * my;
    P4B_CHECK(p4b_vec_norm2(c, N, u, &unorm));                 // the "wp" differencing step depends on ||u||_2
    P4B_CUDA(cudaMallocAsync((void **)&tmp, sizeof(double) * 2 * N, c->stream));
    const int rc = fd_jacobian_minimal(c->stream, mx, my, q, unorm, u, g, F0, vals9, tmp, tmp + N);
    cudaFreeAsync(tmp, c->stream);
    return rc;
}
int p4b_poisson_stencil9(p4b_ctx *c, int mx, int my, double Lx, double Ly, double cx, double cy, double *vals9) {
    if (mx < 3 || my < 3) return fail(60, "Poisson matrix: grid must be at least 3 x 3");
    return launch_poisson_stencil9(c->stream, mx, my, Lx, Ly, cx, cy, vals9);
}
int p4b_stencil9_apply(p4b_ctx *c, int mx, int my, const double *vals9, const double *x, double *y) {
    return launch_stencil9_apply(c->stream, mx, my, vals9, x, y);
}
int p4b_stencil9_lin(p4b_ctx *c, int mx, int my, const double *vals9, const double *u, const double *b, const double *pm1,
                     double ca, double cb, double cg, int jacobi, double *out) {
    return launch_stencil9_lin(c->stream, mx, my, vals9, u, b, pm1, ca, cb, cg, jacobi, out);
}
int p4b_dense_matvec(p4b_ctx *c, int n, const double *Ainv, const double *b, double *x) {
    return launch_dense_matvec(c->stream, n, Ainv, b, x);
}
int p4b_stencil9_gershgorin(p4b_ctx *c, int mx, int my, const double *vals9, double *work, double *res) {
    P4B_CHECK(launch_stencil9_rowratio(c->stream, mx, my, vals9, work));
    return p4b_vec_norminf(c, (size_t)mx * my, work, res);
}
int p4b_inject2d(p4b_ctx *c, int cmx, int cmy, const double *uf, double *uc) {
    return launch_inject2d(c->stream, cmx, cmy, 2 * cmx - 1, uf, uc);
}
int p4b_vec_axpby(p4b_ctx *c, size_t n, double a, const double *x, double b, const double *y, double *out) {
    return launch_axpby_out(c->stream, (long long)n, a, x, b, y, out);
}
int p4b_vi_inactive_mask(p4b_ctx *c, size_t n, const double *u, const double *lower, const double *F, double *mask) {
    return launch_vi_mask(c->stream, (long long)n, u, lower, F, mask);
}
int p4b_vec_pointwise_mult(p4b_ctx *c, size_t n, const double *x, const double *y, double *out) {
    return launch_pointwise(c->stream, (long long)n, 0, x, y, out);
}
int p4b_vec_pointwise_max(p4b_ctx *c, size_t n, const double *x, const double *y, double *out) {
    return launch_pointwise(c->stream, (long long)n, 1, x, y, out);
}
int p4b_vec_copy(p4b_ctx *c, size_t n, const double *x, double *y) {
    P4B_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

struct p4b_sell {
    p4b_ctx *ctx;
    Sell *A;
};
int p4b_sell_create(p4b_ctx *c, int nrows, const int *rowptr, const int *colind, const double *vals, p4b_sell **out) {
    Sell *A = nullptr;
    P4B_CHECK(sell_build(c->stream, nrows, rowptr, colind, vals, &A));
    *out = new p4b_sell{c, A};
    return 0;
}
int p4b_sell_spmv(p4b_sell *S, const double *x, double *y) { return sell_spmv(S->ctx->stream, S->A, x, y); }
int p4b_sell_info(p4b_sell *S, int *nrows, long long *nnz, long long *padded) {
    sell_info(S->A, nrows, nnz, padded);
    return 0;
}
int p4b_sell_destroy(p4b_sell *S) {
    if (!S) return 0;
    cudaStreamSynchronize(S->ctx->stream);
    sell_free(S->A);
    delete S;
    return 0;
}

// ---- profiler -----------------------------------------------------------------------------------------
int p4b_profile_enable(p4b_mg *m, int on) { m->prof.on = on < 0 ? 0 : (on > 2 ? 2 : on); return 0; }
int p4b_profile_get_level(p4b_mg *m, int level, int cls, p4b_kernel_stat *out) {
    if (cls < 0 || cls >= P4B_K_NCLASSES || level < 0 || level > m->top) return fail(62, "bad level / kernel class");
    *out = m->prof.lstat[level][cls];
    return 0;
}
int p4b_profile_reset(p4b_mg *m) {
    memset(m->prof.stat, 0, sizeof m->prof.stat);
    memset(m->prof.lstat, 0, sizeof m->prof.lstat);
    return 0;
}
int p4b_profile_get(p4b_mg *m, int cls, p4b_kernel_stat *out) {
    if (cls < 0 || cls >= P4B_K_NCLASSES) return fail(62, "kernel class out of range");
    *out = m->prof.stat[cls];
    return 0;
}
long long p4b_launch_count(void) { return g_launch_count; }

}  // extern "C"
