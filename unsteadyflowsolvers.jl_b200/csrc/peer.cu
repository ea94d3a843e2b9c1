// peer.cu -- CUDA-IPC peer mapping, flag barrier and the fused finish + exchange kernel (see peer.cuh)
#include "peer.cuh"
#include "induce_sym.cuh"
#include "runtime.h"
#include <cstdlib>
#include <cstring>
#include <vector>

namespace unsflow {

struct PeerFlagPtrs {
    unsigned long long *f[kMaxPeers];
};

// thread q signals rank q (release store of the epoch into slot `rank` of q's flag array), then waits until rank q
// has signalled this rank; ~40 s watchdog (clock64) so that a dead peer cannot hang the GPU for ever
__global__ void k_peer_barrier(PeerFlagPtrs p, int rank, int nranks, unsigned long long epoch, int *err) {
    const int q = threadIdx.x;
    if (q < nranks) {
        __threadfence_system();
        unsigned long long *dst = p.f[q] + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
        const unsigned long long *src = p.f[rank] + q;
        const long long t0 = clock64();
        unsigned long long v = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 80000000000LL) { *err = 1; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

struct PeerPack {
    cudaIpcMemHandle_t h[kPeerBufs + 1];
    int ok;
    int pad[15];
};

static void peer_close(PeerGroup *g) {
    for (int q = 0; q < g->nranks; q++) {
        if (q == g->rank) continue;
        for (int b = 0; b < kPeerBufs; b++)
            if (g->opened[q][b] && g->bufs[q][b]) { cudaIpcCloseMemHandle(g->bufs[q][b]); g->opened[q][b] = false; g->bufs[q][b] = nullptr; }
        if (g->opened[q][kPeerBufs] && g->flags[q]) { cudaIpcCloseMemHandle(g->flags[q]); g->opened[q][kPeerBufs] = false; g->flags[q] = nullptr; }
    }
    (void)cudaGetLastError();
}

int peer_setup(PeerGroup *g, ncclComm_t comm, cudaStream_t st, int rank, int nranks, void *const *local_bufs) {
    g->rank = rank; g->nranks = nranks; g->ready = false;
    static const bool off = [] { const char *e = getenv("UNSFLOW_NO_PEER"); return e && atoi(e) != 0; }();
    const NcclApi *nc = nccl_api();
    if (!nc || nranks < 2) return 0;
    int ok = (!off && nranks <= kMaxPeers) ? 1 : 0;
    if (ok && (cudaMalloc((void **)&g->my_flags, sizeof(unsigned long long) * kMaxPeers) != cudaSuccess ||
               cudaMalloc((void **)&g->err, sizeof(int)) != cudaSuccess))
        ok = 0;
    if (ok) {
        cudaMemsetAsync(g->my_flags, 0, sizeof(unsigned long long) * kMaxPeers, st);
        cudaMemsetAsync(g->err, 0, sizeof(int), st);
    }
    PeerPack mine;
    memset(&mine, 0, sizeof(mine));
    for (int b = 0; b < kPeerBufs && ok; b++)
        if (cudaIpcGetMemHandle(&mine.h[b], local_bufs[b]) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine.h[kPeerBufs], g->my_flags) != cudaSuccess) ok = 0;
    (void)cudaGetLastError();
    mine.ok = ok;
    // every rank takes part in both exchanges whatever its own state, so that nobody is left waiting
    DevBuf<PeerPack> all;
    if (all.alloc(nranks)) return 2;
    std::vector<PeerPack> hall((size_t)nranks);
    UF_CUDA(cudaMemcpyAsync(all.p + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    if (nc->AllGather(all.p + rank, all.p, sizeof(PeerPack), ncclChar, comm, st) != ncclSuccess) { set_error("peer_setup: ncclAllGather failed"); return 4; }
    UF_CUDA(cudaMemcpyAsync(hall.data(), all.p, sizeof(PeerPack) * nranks, cudaMemcpyDeviceToHost, st));
    UF_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < nranks; q++) ok = ok && hall[q].ok;
    if (ok) {
        for (int q = 0; q < nranks && ok; q++) {
            if (q == rank) {
                for (int b = 0; b < kPeerBufs; b++) g->bufs[q][b] = local_bufs[b];
                g->flags[q] = g->my_flags;
                continue;
            }
            for (int b = 0; b <= kPeerBufs && ok; b++) {
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, hall[q].h[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
                g->opened[q][b] = true;
                if (b < kPeerBufs) g->bufs[q][b] = ptr; else g->flags[q] = (unsigned long long *)ptr;
            }
        }
        (void)cudaGetLastError();
    }
    // second agreement: did everybody manage to map everybody?
    DevBuf<int> oks;
    if (oks.alloc(nranks)) return 2;
    std::vector<int> hoks((size_t)nranks, 0);
    UF_CUDA(cudaMemcpyAsync(oks.p + rank, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
    if (nc->AllGather(oks.p + rank, oks.p, sizeof(int), ncclChar, comm, st) != ncclSuccess) { set_error("peer_setup: ncclAllGather failed"); return 4; }
    UF_CUDA(cudaMemcpyAsync(hoks.data(), oks.p, sizeof(int) * nranks, cudaMemcpyDeviceToHost, st));
    UF_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < nranks; q++) ok = ok && hoks[q];
    if (!ok) { peer_close(g); return 0; }
    g->ready = true;
    g->epoch = 0;
    return 0;
}

int launch_peer_barrier(PeerGroup *g, cudaStream_t st) {
    PeerFlagPtrs p;
    for (int q = 0; q < kMaxPeers; q++) p.f[q] = q < g->nranks ? g->flags[q] : nullptr;
    g->epoch++;
    k_peer_barrier<<<1, 32, 0, st>>>(p, g->rank, g->nranks, g->epoch, g->err);
    UF_CUDA(cudaGetLastError());
    return 0;
}

int peer_check(PeerGroup *g, cudaStream_t st) {
    if (!g->ready) return 0;
    int e = 0;
    UF_CUDA(cudaMemcpyAsync(&e, g->err, sizeof(int), cudaMemcpyDeviceToHost, st));
    UF_CUDA(cudaStreamSynchronize(st));
    if (e) { set_error("peer barrier timed out: a rank of the multi-GPU stepper stopped responding"); return 4; }
    return 0;
}

void peer_destroy(PeerGroup *g) {
    if (g->ready) peer_close(g);
    g->ready = false;
    if (g->my_flags) cudaFree(g->my_flags);
    if (g->err) cudaFree(g->err);
    g->my_flags = nullptr;
    g->err = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused finish + exchange.  Every rank holds, in its own `sums` buffer, the plain partial velocities of ITS share of
// the pair tiles for all live vortices (packed by live ordinal: [u(stride) | w(stride)]).  Rank r owns the ordinals
// [ntot r / G, ntot (r+1) / G): it adds the G partial sums of those vortices (read from the peers over NVLink, in rank
// order: deterministic), applies 1/2pi + free stream, convects the vortex (explicit Euler, calcs.jl:252-291) and writes
// velocity and new position into EVERY rank's state -- all ranks end the step with bit-identical wakes.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peer_finish(const PeerFinishArgs a) {
    const Regions reg = *a.reg;
    long long ntot = 0;
    for (int r = 0; r < a.nreg; r++) ntot += reg.end[r] > reg.begin[r] ? reg.end[r] - reg.begin[r] : 0;
    long long o0, o1;
    split_range(ntot, a.rank, a.nranks, &o0, &o1);
    const long long ord = o0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ord >= o1) return;
    long long lin = ord, i = -1;
    for (int r = 0; r < a.nreg; r++) {
        const long long n = reg.end[r] > reg.begin[r] ? reg.end[r] - reg.begin[r] : 0;
        if (lin < n) { i = reg.begin[r] + lin; break; }
        lin -= n;
    }
    if (i < 0) return;
    double su = 0.0, sw = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxPeers; q++)
        if (q < a.nranks) {
            su += __ldcv(a.sums[q] + ord);
            sw += __ldcv(a.sums[q] + a.stride + ord);
        }
    const double vx = su * kInvTwoPi + a.fs[0], vz = sw * kInvTwoPi + a.fs[1];
    const double x = a.slab[a.rank][a.off_x + i] + a.dt * vx, z = a.slab[a.rank][a.off_z + i] + a.dt * vz;
#pragma unroll
    for (int q = 0; q < kMaxPeers; q++)
        if (q < a.nranks) {
            double *s = a.slab[q];
            s[a.off_vx + i] = vx; s[a.off_vz + i] = vz; s[a.off_x + i] = x; s[a.off_z + i] = z;
        }
    __threadfence_system();
}

int launch_peer_finish(const PeerFinishArgs &a, long long ntargets_ub, cudaStream_t st) {
    const long long mine = (ntargets_ub + a.nranks - 1) / a.nranks + 1;
    const int blocks = (int)((mine + 255) / 256);
    k_peer_finish<<<blocks, 256, 0, st>>>(a);
    UF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace unsflow
