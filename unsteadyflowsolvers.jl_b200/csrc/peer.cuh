// peer.cuh -- NVLink peer-memory plumbing for the fused "finish + exchange" step of the multi-GPU stepper.
//
// One process per GPU.  Every rank maps the other ranks' device buffers into its own address space (CUDA IPC
// handles, exchanged once through the NCCL communicator the handle already owns), so that a kernel can read the
// peers' partial sums and write the result into the peers' state directly over NVLink / NVSwitch -- no staging,
// no collective library call in the step.  Synchronisation between the ranks' streams is a flag barrier in
// peer memory (k_peer_barrier: system-scope release stores into every peer's flag array, acquire spin on the own
// one, bounded by a watchdog so that a dead peer turns into an error instead of a hung GPU).
#pragma once
#include "common.cuh"
#include "induce.cuh"
#include "nccl_dl.h"

namespace unsflow {

constexpr int kMaxPeers = 8;       // one NVSwitch node
constexpr int kPeerBufs = 2;       // buffers shared per rank: [0] packed partial sums, [1] the wake slabs

struct PeerGroup {
    int rank = 0, nranks = 1;
    bool ready = false;
    void *bufs[kMaxPeers][kPeerBufs] = {};            // bufs[q][b]: rank q's buffer b in THIS process' address space
    unsigned long long *flags[kMaxPeers] = {};       // flags[q]: rank q's flag array [kMaxPeers]
    unsigned long long *my_flags = nullptr;           // own allocation (cudaMalloc), IPC-exported
    int *err = nullptr;                               // device flag set by a barrier that timed out
    unsigned long long epoch = 0;                     // barriers passed so far (identical on all ranks)
    bool opened[kMaxPeers][kPeerBufs + 1] = {};
};

// exchange the IPC handles of local_bufs[kPeerBufs] (cudaMalloc base pointers) and of a fresh flag array; returns 0 and
// sets g->ready on success on ALL ranks, otherwise leaves g->ready false everywhere (the caller keeps its NCCL path)
int peer_setup(PeerGroup *g, ncclComm_t comm, cudaStream_t st, int rank, int nranks, void *const *local_bufs);
void peer_destroy(PeerGroup *g);
// enqueue a barrier over all ranks on `st`: everything the ranks enqueued before it is visible to all after it
int launch_peer_barrier(PeerGroup *g, cudaStream_t st);
// non-zero when a barrier of this group ever timed out (checked at synchronisation points)
int peer_check(PeerGroup *g, cudaStream_t st);

// fused finish + exchange of the multi-GPU stepper (k_peer_finish, peer.cu)
struct PeerFinishArgs {
    const Regions *reg;              // live target regions (local copy; identical on all ranks)
    int nreg;
    int rank, nranks;
    long long stride;                // packed layout of the partial sums: [u(stride) | w(stride)] by live ordinal
    const double *sums[kMaxPeers];   // every rank's partial sums
    double *slab[kMaxPeers];         // every rank's wake slabs (base pointer of the 7-slab allocation)
    long long off_x, off_z, off_vx, off_vz;   // slab offsets in doubles
    const double *fs;                // (u_inf, w_inf), local
    double dt;
};
int launch_peer_finish(const PeerFinishArgs &a, long long ntargets_ub, cudaStream_t st);

}  // namespace unsflow
