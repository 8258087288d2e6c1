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
};

struct PeerTable {
    int rank, nranks;
    Mailbox *mbox[MAX_RANKS];                   // mbox[rank] is the local one
};

struct GatherTable {
    double *base[MAX_RANKS];                    // arena base of every rank (peer mapped)
};

int launch_halo_push(cudaStream_t st, const double *lo_src, double *lo_dst, const double *hi_src, double *hi_dst,
                     long long plane, unsigned long long *flag_prev, unsigned long long *flag_next,
                     const unsigned long long *my_flags, LocalSync *sync);
int launch_allreduce(cudaStream_t st, double *vals, int nv, int op_max, const PeerTable &peers, LocalSync *sync);
int launch_barrier(cudaStream_t st, int all, const PeerTable &peers, LocalSync *sync);
int launch_gather_push(cudaStream_t st, const double *src, long long n, long long off_doubles, const GatherTable &dst,
                       int rank, int nranks);

}  // namespace p4b
