// ldvm.cu -- device-resident LDVM time stepper (ldvm / ldvm2DOF / ldvmLin / lautat).
//
// One time step of the reference's `for istep = 1:nsteps` body (src/lowOrder2D/solvers.jl
// :178-257, :697-778, :488-644, :76-131) is five launches on one stream, no host round trip:
//
//   k_step_head   (1 CTA)  kinematics row / 2DOF Adams-Bashforth, update_boundpos, place_tev
//   k_induce_direct        bound-point sweep   update_indbound          (70 x N)
//   k_step_solve  (1 CTA)  Kelvin solve, LESP test, place_lev, Kelvin-Kutta solve, Fourier
//                          coefficients + rates, update_bv, Spalart merge, calc_forces, mat row
//   k_induce_direct        wake roll-up        mutual_ind + ind_vel(bv) (N x (N+69))
//   k_induce_finish        1/2pi, free stream, explicit-Euler convection
//
// Wakes below ~1600 vortices (single surface, ndiv <= 72, one rank) take k_cluster_steps instead: 64 steps per launch of
// one 16-CTA thread-block cluster, the same phases separated by cluster barriers (step_cluster.cuh).
//
// Vortex counts live in device memory (Regions); grids are sized from host-side upper bounds
// and surplus blocks exit.  Vortex families are slabs of one SoA allocation with a moving
// head (popfirst! of the Spalart merge = zero the strength and advance the head).
#include "../../include/unsflow.h"
#include "induce_sym.cuh"
#include "sweep.cuh"
#include "nccl_dl.h"
#include "peer.cuh"
#include "runtime.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "step_single.cuh"
#include "step_multi.cuh"
#include "step_lve.cuh"
#include "step_ensemble.cuh"
#include "step_cluster.cuh"

namespace unsflow {

__global__ void __launch_bounds__(kSolveThreads) k_step_head(const LdvmDev d) { step_head(d); }
// shared-memory copy of the cos / sin tables, filled by TMA bulk copies issued by one thread at kernel entry
__device__ __forceinline__ void tabs_to_smem(const LdvmDev &d, double *dst, uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const uint32_t bytes = (uint32_t)(sizeof(double) * 2 * (d.naterm + 1) * d.ndiv);
        mbar_expect_tx(bar, bytes);
        for (uint32_t o = 0; o < bytes; o += 16384)
            tma_load_1d(reinterpret_cast<char *>(dst) + o, reinterpret_cast<const char *>(d.costab) + o, min(16384u, bytes - o), bar);
    }
    __syncthreads();   // the initialised barrier is visible to every waiter
}

__global__ void __launch_bounds__(kSolveThreads) k_step_solve(const LdvmDev d, int tabs_in_smem) {
    // dynamic shared memory: [step_solve workspace | cos / sin tables (optional)]
    extern __shared__ __align__(128) double solve_dyn[];
    __shared__ __align__(8) uint64_t tab_bar;
    double *tab_s = solve_dyn + (solve_ws_doubles(d.ndiv, d.naterm) + 15) / 16 * 16;
    if (tabs_in_smem) {
        tabs_to_smem(d, tab_s, &tab_bar);
        step_solve(d, tab_s, tab_s + (d.naterm + 1) * d.ndiv, &tab_bar, solve_dyn);
    } else {
        step_solve(d, d.costab, d.sintab, nullptr, solve_dyn);
    }
}

// small wakes: a chunk of whole time steps in ONE launch of one thread-block cluster (step_cluster.cuh)
// StepState <-> global, word by word (all threads of the CTA)
__device__ __forceinline__ void copy_state(StepState *dst, const StepState *src) {
    static_assert(sizeof(StepState) % 8 == 0, "StepState is copied as 8-byte words");
    const long long *s8 = reinterpret_cast<const long long *>(src);
    long long *d8 = reinterpret_cast<long long *>(dst);
    for (int k = threadIdx.x; k < (int)(sizeof(StepState) / 8); k += blockDim.x) d8[k] = s8[k];
}

template <bool UNIFORM>
__global__ void __launch_bounds__(kSolveThreads, 1) k_cluster_steps(const ClusterArgs a) {
    // dynamic shared memory: [step_solve workspace | cos / sin tables (optional) | surface copy | source staging of the roll-up]
    extern __shared__ __align__(128) double cl_dyn[];
    __shared__ __align__(8) uint64_t tab_bar;
    __shared__ StepState sst;
    const unsigned rank = cl_rank(), C = cl_size();
    const LdvmDev &d = a.d;
    const int tid = threadIdx.x, nd = d.ndiv, na = d.naterm, ndp = (nd + 7) / 8 * 8, nap = (na + 7) / 8 * 8;
    double *ws = cl_dyn;
    double *tab_s = ws + (solve_ws_doubles(nd, na) + 15) / 16 * 16;
    double *surf_s = tab_s + (a.tabs_in_smem ? (2 * (na + 1) * nd + 15) / 16 * 16 : 0);
    double *stg = surf_s + (cluster_surf_doubles(nd, na) + 15) / 16 * 16;
    if (rank == 0 && a.tabs_in_smem) {      // once per chunk, not once per step
        tabs_to_smem(d, tab_s, &tab_bar);
        mbar_wait(&tab_bar, 0);
    }
    // CTA 0 keeps StepState and the surface arrays in shared memory for the whole chunk (its single-thread sections walk
    // shared-memory latencies instead of L2 round trips, as an ensemble member does) and publishes what the other CTAs
    // read -- the live ranges / free stream (StepState) and the bound points -- to global memory before each barrier
    // (a per-thread copy of the descriptor with the redirected pointers: after inlining, the untouched fields are still
    // read from the kernel-parameter constant bank -- a descriptor selected at run time or kept in shared memory made
    // every field access of step_solve a load: 15.8 -> 17.1 us per C1 step)
    const bool local = rank == 0;
    LdvmDev d0 = d;
    d0.st = &sst;
    d0.bnd_x = surf_s; d0.bnd_z = surf_s + ndp; d0.uind = surf_s + 2 * ndp; d0.wind = surf_s + 3 * ndp;
    d0.downwash = surf_s + 4 * ndp; d0.theta = surf_s + 5 * ndp; d0.x = surf_s + 6 * ndp; d0.cam = surf_s + 7 * ndp;
    d0.cam_slope = surf_s + 8 * ndp;
    d0.aterm = surf_s + 9 * ndp; d0.adot = surf_s + 9 * ndp + nap; d0.aprev = surf_s + 9 * ndp + 2 * nap;
    if (local) {
        double *const g[9] = {d.bnd_x, d.bnd_z, d.uind, d.wind, d.downwash, d.theta, d.x, d.cam, d.cam_slope};
        for (int q = 0; q < 9; q++)
            for (int k = tid; k < nd; k += blockDim.x) surf_s[q * ndp + k] = g[q][k];
        double *const ga[3] = {d.aterm, d.adot, d.aprev};
        for (int q = 0; q < 3; q++)
            for (int k = tid; k < na; k += blockDim.x) surf_s[9 * ndp + q * nap + k] = ga[q][k];
        copy_state(&sst, d.st);
        __syncthreads();
    }
    const double *ct = a.tabs_in_smem ? tab_s : d.costab;
    const double *stb = a.tabs_in_smem ? tab_s + (na + 1) * nd : d.sintab;
    __shared__ ClIndex s_ix;      // live ranges as published after the last solve (what the next bound sweep walks)
    __shared__ double hs[4];      // trailing edge + u, from step_head_a to step_head_b
    const bool lautat = d.driver == UNSFLOW_DRIVER_LAUTAT;
    // prologue: first half of the first step's head; the helpers note the live ranges before anything is published
    if (tid == 0) s_ix.load(d.st);
    if (local) {
        step_head_a(d0, hs);
        __syncthreads();
        for (int k = tid; k < nd; k += blockDim.x) { d.bnd_x[k] = surf_s[k]; d.bnd_z[k] = surf_s[ndp + k]; }
    }
    cl_sync();
    long long pc[4] = {0, 0, 0, 0}, t0 = clock64();
    for (long long is = 0; is < a.nsteps; is++) {
        // ---- bound sweep on the helpers; CTA 0 places the TEV and publishes the step's state meanwhile ----
        if (local) {
            step_head_b(d0, hs);
            __syncthreads();
            copy_state(d.st, &sst);
        } else if (!lautat) {
            cl_sweep<UNIFORM>(a, s_ix, rank - 1, C - 1, stg);
        }
        cl_sync();
        long long t1 = clock64(); pc[1] += t1 - t0;
        // (counts as the head left them: identical on every CTA)
        const long long n1_pre = d.st->wake.end[1] - d.st->wake.begin[1];
        long long nf_pre = n1_pre;
        for (int r = 0; r < 3; r += 2) { const long long c = d.st->wake.end[r] - d.st->wake.begin[r]; nf_pre += c > 0 ? c : 0; }
        const bool overlap = a.overlap && cl_overlap_ok(d, C, nf_pre);
        ClEarly es;
        if (local) {
            step_solve(d0, ct, stb, nullptr, ws);
            __syncthreads();
            copy_state(d.st, &sst);
        } else if (overlap) cl_rollup_early<UNIFORM>(a, rank, C, stg, es);      // wake-on-wake part, while CTA 0 solves
        cl_sync();
        t0 = clock64(); pc[2] += t0 - t1;
        if (overlap) cl_rollup_late<UNIFORM>(a, rank, C, stg, es, n1_pre);
        else cl_rollup<UNIFORM>(a, rank, C, stg);
        if (tid == 0) s_ix.load(d.st);      // as the solve left them; stable until CTA 0 publishes the next head
        // ---- CTA 0: first half of the NEXT step's head while the others finish the roll-up ----
        if (local && is + 1 < a.nsteps) {
            step_head_a(d0, hs);
            __syncthreads();
            for (int k = tid; k < nd; k += blockDim.x) { d.bnd_x[k] = surf_s[k]; d.bnd_z[k] = surf_s[ndp + k]; }
        }
        cl_sync();
        t1 = clock64(); pc[3] += t1 - t0; t0 = t1;
    }
    if (local) {      // the surface back to global memory (StepState was published after the last solve)
        __syncthreads();
        double *const g[5] = {d.bnd_x, d.bnd_z, d.uind, d.wind, d.downwash};
        for (int q = 0; q < 5; q++)
            for (int k = tid; k < nd; k += blockDim.x) g[q][k] = surf_s[q * ndp + k];
        double *const ga[3] = {d.aterm, d.adot, d.aprev};
        for (int q = 0; q < 3; q++)
            for (int k = tid; k < na; k += blockDim.x) ga[q][k] = surf_s[9 * ndp + q * nap + k];
    }
    if (a.prof && rank == 0 && threadIdx.x == 0) {
        for (int q = 0; q < 4; q++) atomicAdd((unsigned long long *)a.prof + q, (unsigned long long)pc[q]);
        atomicAdd((unsigned long long *)a.prof + 4, (unsigned long long)a.nsteps);
    }
}

__global__ void __launch_bounds__(kSolveThreads) k_mstep_head(const LdvmDev d) { mstep_head(d); }
__global__ void __launch_bounds__(kSolveThreads) k_mstep_solve(const LdvmDev d) { mstep_solve(d); }


__global__ void __launch_bounds__(kSolveThreads) k_lve_head(const LdvmDev d) { lve_head(d); }
__global__ void __launch_bounds__(kSolveThreads) k_lve_solve(const LdvmDev d) {
    extern __shared__ __align__(16) double lve_dyn[];
    lve_solve(d, lve_dyn);
}

// OCC = member CTAs per SM the register budget is cut for (3: 85 registers, 2: 128); UN = independent source chains
// per thread in the member roll-up and bound sweep (what the register budget lets the compiler keep interleaved)
template <int THREADS, int OCC, int UN, int SYM>
__global__ void __launch_bounds__(THREADS, OCC) k_ensemble(const EnsArgs a) {
    extern __shared__ __align__(16) double ens_smem[];
    __shared__ LdvmDev d;
    __shared__ StepState sst;
    __shared__ int s_member;
    const int tid = threadIdx.x, nd = a.proto.ndiv, na = a.proto.naterm;
    const long long tot = a.tot, ndp = (nd + 7) / 8 * 8, nap = (na + 7) / 8 * 8;
    // layout: [x | z | g | vx | vz](tot each) [bnd_x | bnd_z | uind | wind | downwash](ndp each) [aterm | adot | aprev](nap each)
    //         [p2u | p2w](groups * ndp each) [solve workspace]
    double *wk = ens_smem, *sm_surf = wk + 5 * tot, *sm_at = sm_surf + 5 * ndp, *sm_p2 = sm_at + 3 * nap;
    double *ws = sm_p2 + 2 * (long long)a.groups * ndp;
    long long prof[6] = {0, 0, 0, 0, 0, 0};      // cycles: head, sweep, solve, roll-up, set-up; [5] = roll-up interactions
    for (;;) {
        long long c0 = clock64();
        __syncthreads();                                  // everyone is done with the previous member's shared state
        if (tid == 0) s_member = atomicAdd(a.next, 1);
        __syncthreads();
        const int slot = s_member;
        if (slot >= a.M) break;
        const long long m = a.order[slot];
        // ---- member set-up: empty wake, bound vortices and surface state from the shared image ----
        for (long long k = tid; k < 5 * tot; k += blockDim.x) wk[k] = 0.0;
        __syncthreads();
        for (int k = tid; k < nd; k += blockDim.x)
            for (int q = 0; q < 5; q++) sm_surf[q * ndp + k] = a.img[q * nd + k];
        for (int k = tid; k < na; k += blockDim.x)
            for (int q = 0; q < 3; q++) sm_at[q * nap + k] = a.img[5 * nd + q * na + k];
        for (int k = tid; k < nd - 1; k += blockDim.x)
            for (int q = 0; q < 3; q++) wk[q * tot + a.off_b + k] = a.img[5 * nd + 3 * na + q * (nd - 1) + k];
        if (tid == 0) {
            sst = a.st[m];
            d = a.proto;
            d.st = &sst;
            d.lespcrit = a.lespcrit ? a.lespcrit[m] : a.proto.lespcrit;
            d.bnd_x = sm_surf; d.bnd_z = sm_surf + ndp; d.uind = sm_surf + 2 * ndp; d.wind = sm_surf + 3 * ndp; d.downwash = sm_surf + 4 * ndp;
            d.aterm = sm_at; d.adot = sm_at + nap; d.aprev = sm_at + 2 * nap;
            d.wx = wk; d.wz = wk + tot; d.wg = wk + 2 * tot; d.wvx = wk + 3 * tot; d.wvz = wk + 4 * tot;
            d.wvc4 = nullptr; d.wvc = nullptr;             // scalar core radius (d.vc4_new)
            d.kin = a.kin + m * a.nsteps * 9; d.kin_rows = a.nsteps;
            d.mat = a.mat + m * a.nsteps * 8;
            d.p2u = sm_p2; d.p2w = sm_p2 + (long long)a.groups * ndp; d.p2stride = ndp; d.S2 = 1;
        }
        __syncthreads();
        long long c1 = clock64();
        prof[4] += c1 - c0;
        for (long long is = 0; is < a.nsteps; is++) {
            step_head(d);
            __syncthreads();
            c0 = clock64(); prof[0] += c0 - c1;
            if (d.driver != UNSFLOW_DRIVER_LAUTAT) {
                int S2;
                ens_sweep<UN>(d, sm_p2, sm_p2 + (long long)a.groups * ndp, &S2, wk, tot);
                if (tid == 0) d.S2 = S2;
                __syncthreads();
            }
            c1 = clock64(); prof[1] += c1 - c0;
            step_solve(d, d.costab, d.sintab, nullptr, ws);
            __syncthreads();
            c0 = clock64(); prof[2] += c0 - c1;
            ens_rollup<UN, SYM>(d, wk, tot, a.sym_min);
            c1 = clock64(); prof[3] += c1 - c0;
            if (a.prof && tid == 0) {
                const long long nfree = (sst.wake.end[0] - sst.wake.begin[0]) + (sst.wake.end[1] - sst.wake.begin[1]);
                prof[5] += nfree * (nfree + (sst.wake.end[3] - sst.wake.begin[3]));
            }
        }
        if (tid == 0) a.st[m] = sst;
    }
    if (a.prof && tid == 0)
        for (int q = 0; q < 6; q++) a.prof[(long long)blockIdx.x * 8 + q] = prof[q];
}

}  // namespace unsflow

using namespace unsflow;

static thread_local float g_batch_ms = 0.f;   // CUDA-event time of the last k_ensemble launch

struct unsflow_ldvm {
    unsflow_opts opts{};
    int ndiv = 0, naterm = 0;
    int nsurf = 1, mat_cols = 8;
    DevBuf<SurfState> ms;
    DevBuf<LveState> lves;    // LVE driver only
    DevBuf<double> lvea, lve_binv;
    int64_t cap_t = 0, cap_l = 0, cap_e = 0, bvpad = 0, tot = 0;
    int64_t off_t = 0, off_l = 0, off_e = 0, off_b = 0;
    int64_t ntev0 = 0, nlev0 = 0, nextv = 0;
    int64_t steps_done = 0;          // since create
    int64_t mat_cap = 0;
    bool uniform = true;
    double vc4u = 1.0;
    // tightened count knowledge (async read-back)
    int64_t known_step = 0, known_ntev_end = 0, known_nlev_end = 0;
    DevBuf<StepState> st;
    DevBuf<double> surf;      // 9*ndiv + 3*naterm
    DevBuf<double> tabs;      // costab, sintab
    DevBuf<double> wake;      // 7 slabs
    DevBuf<double> kin;       // rows x 9
    DevBuf<double> mat;       // mat_cap x 8
    DevBuf<long long> clprof; // optional phase counters of the cluster stepper
    bool cluster_off = false; // a cluster launch was refused on this device: five-kernel path only
    DevBuf<double> p2;        // bound-sweep partials [2][S2max][ndivpad]
    DevBuf<double> p1;        // roll-up partials [2][S1max][tot]
    DevBuf<double> planes;    // symmetric kernel partial planes [ctas][2][tot] (allocated on first use)
    int ctas = 0;
    int S2max = 1, S1max = 1;
    int64_t ndivpad = 0;
    int64_t kin_rows = 0;
    LdvmDev dev{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc = nullptr;
    StepState *pinned = nullptr;
    bool readback_pending = false;
    int64_t readback_step = 0;
    int64_t launches = 0;
    float last_ms = 0;
    // multi-GPU: pair tiles of the roll-up split over ranks, everything else replicated bit-identically
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    DevBuf<double> vsum;      // [2][tot] raw plane sums exchanged by all-reduce / read by the peers
    PeerGroup peer;           // NVLink peer mapping of vsum and the wake slabs (fused finish + exchange)
    // CUDA-graph cache for the launch-bound regime
    cudaGraphExec_t gexec = nullptr;
    int64_t gkey_t = -1, gkey_l = -1;
    const double *gkey_kin = nullptr;
};

static int ldvm_sync_counts(unsflow_ldvm *h, StepState *out) {
    UF_CUDA(cudaMemcpyAsync(out, h->st.p, sizeof(StepState), cudaMemcpyDeviceToHost, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" {

int unsflow_ldvm_destroy(unsflow_ldvm *h) {
    if (!h) return 0;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->peer.ready) {
        // nobody may unmap / free while a peer kernel can still touch the memory: barrier, then close the imports,
        // then a host-side rendezvous over NCCL before the exported buffers are freed
        launch_peer_barrier(&h->peer, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    peer_destroy(&h->peer);
    if (h->comm && nccl_api() && h->vsum.p && h->nranks > 1) {
        nccl_api()->AllReduce(h->vsum.p, h->vsum.p, 1, ncclDouble, ncclSum, h->comm, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    if (h->comm && nccl_api()) nccl_api()->CommDestroy(h->comm);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evc) cudaEventDestroy(h->evc);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

#define UNSFLOW_DRIVER_LVE_INTERNAL 4
static int ldvm_create_impl(unsflow_ldvm **out, int nsurf, const unsflow_surf *sfs, const unsflow_field *fl, const unsflow_opts *op,
                            bool allow_lve = false) {
    UF_CHECK(out != nullptr, "unsflow_ldvm_create: null handle pointer");
    *out = nullptr;
    UF_CHECK(sfs && fl && op, "unsflow_ldvm_create: null argument");
    UF_CHECK(nsurf >= 1 && nsurf <= kMaxSurf, "unsflow_ldvm_create: nsurf=%d outside [1,%d]", nsurf, kMaxSurf);
    const unsflow_surf *sf = &sfs[0];
    for (int q = 1; q < nsurf; q++) {
        UF_CHECK(sfs[q].ndiv == sf->ndiv && sfs[q].naterm == sf->naterm, "unsflow_ldvm_create: all surfaces must share ndiv and naterm");
        UF_CHECK(sfs[q].theta && sfs[q].x && sfs[q].cam && sfs[q].cam_slope && sfs[q].bnd_x && sfs[q].bnd_z && sfs[q].uind && sfs[q].wind && sfs[q].downwash && sfs[q].aterm && sfs[q].adot && sfs[q].aprev, "unsflow_ldvm_create: null surf array");
        UF_CHECK(sfs[q].bv.n == sf->ndiv - 1, "unsflow_ldvm_create: surf.bv must hold ndiv-1 vortices");
        for (int i = 0; i < sf->ndiv; i++) UF_CHECK(sfs[q].theta[i] == sf->theta[i], "unsflow_ldvm_create: all surfaces must share the theta grid");
    }
    UF_CHECK(nsurf == 1 || op->driver == UNSFLOW_DRIVER_LDVM, "unsflow_ldvm_create: several surfaces are supported by the ldvm driver only");
    UF_CHECK(sf->ndiv >= 3 && sf->ndiv <= kNdMax, "unsflow_ldvm_create: ndiv=%d outside [3,%d]", sf->ndiv, kNdMax);
    UF_CHECK(sf->naterm >= 3 && sf->naterm <= kNaMax, "unsflow_ldvm_create: naterm=%d outside [3,%d]", sf->naterm, kNaMax);
    UF_CHECK(op->driver >= 0 && (op->driver <= 3 || (allow_lve && op->driver == UNSFLOW_DRIVER_LVE_INTERNAL)), "unsflow_ldvm_create: bad driver %d", op->driver);
    UF_CHECK(op->solve_mode == 0 || op->solve_mode == 1, "unsflow_ldvm_create: bad solve_mode %d", op->solve_mode);
    UF_CHECK(op->delvort_kind == 0 || op->delvort_kind == 1, "unsflow_ldvm_create: bad delvort_kind %d", op->delvort_kind);
    UF_CHECK(op->max_steps >= 0 && op->dt > 0, "unsflow_ldvm_create: need max_steps >= 0 and dt > 0");
    UF_CHECK(sf->theta && sf->x && sf->cam && sf->cam_slope && sf->bnd_x && sf->bnd_z && sf->uind && sf->wind &&
                 sf->downwash && sf->aterm && sf->adot && sf->aprev,
             "unsflow_ldvm_create: null surf array");
    UF_CHECK(sf->bv.n == sf->ndiv - 1, "unsflow_ldvm_create: surf.bv must hold ndiv-1 vortices");
    if (int rc = require_device()) return rc;
    if (op->device >= 0) UF_CUDA(cudaSetDevice(op->device));
    unsflow_ldvm *h = new (std::nothrow) unsflow_ldvm();
    UF_CHECK(h != nullptr, "out of host memory");
    auto fail = [&](int rc) { unsflow_ldvm_destroy(h); return rc; };
    h->opts = *op;
    const int nd = h->ndiv = sf->ndiv, na = h->naterm = sf->naterm;
    const int ns = h->nsurf = nsurf;
    h->mat_cols = op->driver == UNSFLOW_DRIVER_LVE_INTERNAL ? 9 : (nsurf == 1 ? 8 : 1 + 7 * nsurf);
    h->ntev0 = fl->tev.n; h->nlev0 = fl->lev.n; h->nextv = fl->extv.n;
    // slabs start on pair-tile boundaries (kSymTB is a multiple of the direct kernel's kTS)
    h->cap_t = round_up(h->ntev0 + ns * op->max_steps + ns, kSymTB);
    h->cap_l = round_up(h->nlev0 + ns * op->max_steps + ns, kSymTB);
    h->cap_e = round_up(std::max<int64_t>(h->nextv, 1), kSymTB);
    h->bvpad = round_up(ns * (nd - 1), kSymTB);
    h->off_t = 0; h->off_l = h->cap_t; h->off_e = h->off_l + h->cap_l; h->off_b = h->off_e + h->cap_e;
    h->tot = h->off_b + h->bvpad;
    h->mat_cap = std::max<int64_t>(op->max_steps, 1);
    h->ndivpad = round_up(ns * nd, 8);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(2); }
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->evc);
    if (cudaMallocHost((void **)&h->pinned, sizeof(StepState)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(2); }
    const int64_t tot = h->tot;
    h->S2max = std::max(256, 2 * sm_count());
    h->S1max = (int)std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)2e9 / (16 * tot)));
    if (h->st.alloc(1) || h->ms.alloc(ns) || h->surf.alloc(ns * (9 * nd + 3 * na)) || h->tabs.alloc(2 * (na + 1) * nd) || h->wake.alloc(7 * tot) ||
        h->mat.alloc(h->mat_cols * h->mat_cap) || h->p2.alloc(2 * (int64_t)h->S2max * h->ndivpad) ||
        h->p1.alloc(2 * (int64_t)h->S1max * tot))
        return fail(2);
    if (h->wake.zero(h->stream) || h->mat.zero(h->stream)) return fail(2);

    // ---- host staging -------------------------------------------------------------------------
    // surface arrays laid out [array][surface][ndiv] (+ [array][surface][naterm])
    std::vector<double> hs((size_t)ns * (9 * nd + 3 * na));
    for (int q = 0; q < ns; q++) {
        const unsflow_surf *sq = &sfs[q];
        const double *sarr[9] = {sq->theta, sq->x, sq->cam, sq->cam_slope, sq->bnd_x, sq->bnd_z, sq->uind, sq->wind, sq->downwash};
        for (int k = 0; k < 9; k++) memcpy(&hs[((size_t)k * ns + q) * nd], sarr[k], sizeof(double) * nd);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)0 * ns + q) * na], sq->aterm, sizeof(double) * na);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)1 * ns + q) * na], sq->adot, sizeof(double) * na);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)2 * ns + q) * na], sq->aprev, sizeof(double) * na);
    }
    std::vector<double> ht((size_t)(2 * (na + 1) * nd));
    for (int n = 0; n <= na; n++)
        for (int i = 0; i < nd; i++) {
            // same expressions as the reference: cos.(ia*theta) calcs.jl:4, sin(ia*theta) :209
            ht[(size_t)n * nd + i] = n == 0 ? 1.0 : std::cos(n * sf->theta[i]);
            ht[(size_t)(na + 1) * nd + (size_t)n * nd + i] = std::sin(n * sf->theta[i]);
        }
    std::vector<double> hw((size_t)(7 * tot), 0.0);
    double *wx = &hw[0], *wz = &hw[tot], *wg = &hw[2 * tot], *wv4 = &hw[3 * tot], *wvc = &hw[4 * tot], *wvx = &hw[5 * tot], *wvz = &hw[6 * tot];
    const double vc_new = 0.02 * sf->c;
    for (int64_t i = 0; i < tot; i++) { wv4[i] = vc_new * vc_new * vc_new * vc_new; wvc[i] = vc_new; }
    bool uni = true;
    auto put = [&](const unsflow_vorts &v, int64_t off) -> int {
        UF_CHECK(v.n == 0 || (v.x && v.z && v.s && v.vc), "unsflow_ldvm_create: null vortex array");
        for (int64_t i = 0; i < v.n; i++) {
            wx[off + i] = v.x[i]; wz[off + i] = v.z[i]; wg[off + i] = v.s[i];
            wvc[off + i] = v.vc[i]; wv4[off + i] = v.vc[i] * v.vc[i] * v.vc[i] * v.vc[i];
            wvx[off + i] = v.vx ? v.vx[i] : 0.0; wvz[off + i] = v.vz ? v.vz[i] : 0.0;
            uni = uni && v.vc[i] == vc_new;
        }
        return 0;
    };
    if (put(fl->tev, h->off_t) || put(fl->lev, h->off_l) || put(fl->extv, h->off_e)) return fail(1);
    for (int q = 0; q < ns; q++)
        if (put(sfs[q].bv, h->off_b + (int64_t)q * (nd - 1))) return fail(1);
    h->uniform = uni;
    h->vc4u = vc_new * vc_new * vc_new * vc_new;

    StepState s0;
    memset(&s0, 0, sizeof(s0));
    s0.wake.begin[0] = h->off_t; s0.wake.end[0] = h->off_t + h->ntev0;
    s0.wake.begin[1] = h->off_l; s0.wake.end[1] = h->off_l + h->nlev0;
    s0.wake.begin[2] = h->off_e; s0.wake.end[2] = h->off_e + h->nextv;
    s0.wake.begin[3] = h->off_b; s0.wake.end[3] = h->off_b + (int64_t)ns * (nd - 1);
    s0.bnd.begin[0] = 0; s0.bnd.end[0] = (int64_t)ns * nd;
    s0.levflag = sf->levflag;
    for (int k = 0; k < 6; k++) s0.kin[k] = sf->kinem[k];
    s0.fs[0] = fl->u; s0.fs[1] = fl->w;
    s0.a0 = sf->a0; s0.a0dot = sf->a0dot; s0.a0prev = sf->a0prev;
    *h->pinned = s0;
    h->known_step = 0; h->known_ntev_end = h->ntev0; h->known_nlev_end = h->nlev0;

    std::vector<SurfState> hms((size_t)ns);
    for (int q = 0; q < ns; q++) {
        SurfState &m = hms[q];
        memset(&m, 0, sizeof(m));
        const unsflow_surf *sq = &sfs[q];
        m.c = sq->c; m.uref = sq->uref; m.pvt = sq->pvt; m.lespcrit = sq->lespcrit;
        m.vc_new = 0.02 * sq->c; m.vc4_new = m.vc_new * m.vc_new * m.vc_new * m.vc_new;
        for (int k = 0; k < 6; k++) m.kin[k] = sq->kinem[k];
        m.a0 = sq->a0; m.a0dot = sq->a0dot; m.a0prev = sq->a0prev; m.levflag = sq->levflag;
        if (q > 0 && m.vc_new != vc_new) uni = false;
    }
    h->uniform = uni;
    cudaStream_t st = h->stream;
    if (cudaMemcpyAsync(h->ms.p, hms.data(), sizeof(SurfState) * ns, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->st.p, &s0, sizeof(s0), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->surf.p, hs.data(), sizeof(double) * hs.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->tabs.p, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->wake.p, hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("unsflow_ldvm_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(2);
    }

    LdvmDev &d = h->dev;
    d.st = h->st.p;
    d.ndiv = nd; d.naterm = na; d.driver = op->driver; d.solve_mode = op->solve_mode; d.del_kind = op->delvort_kind;
    d.del_limit = op->del_limit; d.del_dist = op->del_dist; d.del_tol = op->del_tol;
    d.c = sf->c; d.uref = sf->uref; d.pvt = sf->pvt; d.lespcrit = sf->lespcrit; d.dt = op->dt;
    d.vc_new = vc_new; d.vc4_new = h->vc4u;
    double *sp = h->surf.p;
    const int64_t snd = (int64_t)ns * nd, sna = (int64_t)ns * na;
    d.ms = h->ms.p; d.nsurf = ns;
    d.theta = sp; d.x = sp + snd; d.cam = sp + 2 * snd; d.cam_slope = sp + 3 * snd; d.bnd_x = sp + 4 * snd; d.bnd_z = sp + 5 * snd;
    d.uind = sp + 6 * snd; d.wind = sp + 7 * snd; d.downwash = sp + 8 * snd;
    d.aterm = sp + 9 * snd; d.adot = sp + 9 * snd + sna; d.aprev = sp + 9 * snd + 2 * sna;
    d.costab = h->tabs.p; d.sintab = h->tabs.p + (int64_t)(na + 1) * nd;
    double *wp = h->wake.p;
    d.wx = wp; d.wz = wp + tot; d.wg = wp + 2 * tot; d.wvc4 = wp + 3 * tot; d.wvc = wp + 4 * tot; d.wvx = wp + 5 * tot; d.wvz = wp + 6 * tot;
    d.kin = nullptr; d.kin_rows = 0;
    d.mat = h->mat.p;
    d.p2u = h->p2.p; d.p2w = h->p2.p + (int64_t)h->S2max * h->ndivpad; d.p2stride = h->ndivpad; d.S2 = 1;
    *out = h;
    return 0;
}

int unsflow_ldvm_create(unsflow_ldvm **out, const unsflow_surf *sf, const unsflow_field *fl, const unsflow_opts *op) {
    return ldvm_create_impl(out, 1, sf, fl, op);
}

int unsflow_ldvm_multi_create(unsflow_ldvm **out, int32_t nsurf, const unsflow_surf *surfs, const unsflow_field *fl,
                              const unsflow_opts *op) {
    return ldvm_create_impl(out, nsurf, surfs, fl, op);
}

int unsflow_lve_create(unsflow_ldvm **out, const unsflow_surf *sf, const unsflow_field *ifr_field, const unsflow_opts *op,
                       double rho, double lev_stop, double ifr_xend, double ifr_camend, double ifr_u0, double ifr_hdot0) {
    UF_CHECK(out && sf && ifr_field && op, "unsflow_lve_create: null argument");
    UF_CHECK(sf->ndiv >= 3 && (size_t)sf->ndiv * (sf->ndiv + 1) * sizeof(double) <= 200 * 1024, "unsflow_lve_create: ndiv=%d too large for the in-CTA dense solve (max 159)", sf->ndiv);
    unsflow_opts o = *op;
    o.driver = UNSFLOW_DRIVER_LVE_INTERNAL;
    o.delvort_kind = UNSFLOW_DEL_NONE;
    if (int rc = ldvm_create_impl(out, 1, sf, ifr_field, &o, true)) return rc;
    unsflow_ldvm *h = *out;
    auto fail = [&](int rc) { unsflow_ldvm_destroy(h); *out = nullptr; return rc; };
    const int nd = h->ndiv, np = nd - 1;
    if (h->lves.alloc(1) || h->lvea.alloc((int64_t)kLveArrays * nd)) return fail(2);
    std::vector<double> a((size_t)kLveArrays * nd, 0.0);
    const double d2r = kPi / 180.0;
    for (int j = 0; j < np; j++) {                         // LVE.jl:29-47
        const double dx = sf->x[j] - sf->x[j + 1], dc = sf->cam[j] - sf->cam[j + 1];
        const double ds = std::sqrt(dx * dx + dc * dc);
        const double csd = std::asin((sf->cam[j + 1] - sf->cam[j]) / ds) / d2r;
        const double sd = std::sin(csd * d2r), cd = std::cos(csd * d2r);
        a[LV_DS * nd + j] = ds; a[LV_CSD * nd + j] = csd;
        a[LV_NX * nd + j] = -sd; a[LV_NZ * nd + j] = cd; a[LV_TX * nd + j] = cd; a[LV_TZ * nd + j] = sd;
        a[LV_VX * nd + j] = sf->x[j] + .25 * ds * cd; a[LV_VZ * nd + j] = sf->cam[j] + .25 * ds * sd;
        a[LV_CX * nd + j] = sf->x[j] + .75 * ds * cd; a[LV_CZ * nd + j] = sf->cam[j] + .75 * ds * sd;
    }
    // constant part B of influence_coeff (calcs.jl:756-787): bound-vortex columns (singular kernel, body
    // frame), Kelvin row of ones, and the unit vector e_np in place of the moving wake column; inverted
    // once (Gauss-Jordan, partial pivoting, long double) for the per-step Sherman-Morrison update
    std::vector<long double> B((size_t)nd * nd, 0.0L), Bi((size_t)nd * nd, 0.0L);
    for (int i = 0; i < np; i++)
        for (int j = 0; j < np; j++) {
            const double xd = a[LV_VX * nd + j] - a[LV_CX * nd + i], zd = a[LV_VZ * nd + j] - a[LV_CZ * nd + i];
            const double dsq = xd * xd + zd * zd;
            const double uu = -zd / (kTwoPi * dsq), ww = xd / (kTwoPi * dsq);
            B[(size_t)i * nd + j] = uu * a[LV_NX * nd + i] + ww * a[LV_NZ * nd + i];
        }
    for (int j = 0; j < nd; j++) B[(size_t)np * nd + j] = 1.0L;
    for (int i = 0; i < nd; i++) a[LV_A1 * nd + i] = (double)B[(size_t)i * nd];      // first column of a (calcs.jl:817)
    for (int i = 0; i < nd; i++) Bi[(size_t)i * nd + i] = 1.0L;
    for (int c = 0; c < nd; c++) {
        int piv = c;
        for (int r = c + 1; r < nd; r++) if (fabsl(B[(size_t)r * nd + c]) > fabsl(B[(size_t)piv * nd + c])) piv = r;
        if (fabsl(B[(size_t)piv * nd + c]) == 0.0L) { set_error("unsflow_lve_create: singular panel influence matrix"); return fail(1); }
        if (piv != c) for (int j = 0; j < nd; j++) { std::swap(B[(size_t)c * nd + j], B[(size_t)piv * nd + j]); std::swap(Bi[(size_t)c * nd + j], Bi[(size_t)piv * nd + j]); }
        const long double inv = 1.0L / B[(size_t)c * nd + c];
        for (int j = 0; j < nd; j++) { B[(size_t)c * nd + j] *= inv; Bi[(size_t)c * nd + j] *= inv; }
        for (int r = 0; r < nd; r++) {
            if (r == c) continue;
            const long double f = B[(size_t)r * nd + c];
            if (f == 0.0L) continue;
            for (int j = 0; j < nd; j++) { B[(size_t)r * nd + j] -= f * B[(size_t)c * nd + j]; Bi[(size_t)r * nd + j] -= f * Bi[(size_t)c * nd + j]; }
        }
    }
    std::vector<double> binv((size_t)nd * nd);
    for (size_t k = 0; k < binv.size(); k++) binv[k] = (double)Bi[k];
    if (h->lve_binv.alloc((int64_t)nd * nd)) return fail(2);
    if (cudaMemcpy(h->lve_binv.p, binv.data(), sizeof(double) * binv.size(), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("unsflow_lve_create: upload failed"); return fail(2); }
    LveState L;
    memset(&L, 0, sizeof(L));
    const double x = 1 - 2 * a[LV_DS * nd] / sf->c;        // LVE.jl:50-52
    L.gamma_crit = sf->lespcrit / 1.13 * (sf->kinem[4] * sf->c * (std::acos(x) + std::sin(std::acos(x))));
    L.rho = rho; L.lev_stop = lev_stop; L.te_x = ifr_xend; L.te_cam = ifr_camend; L.le_x = sf->x[0]; L.le_cam = sf->cam[0];
    L.ifr_u0 = ifr_u0; L.ifr_hdot0 = ifr_hdot0;
    if (cudaMemcpy(h->lves.p, &L, sizeof(L), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->lvea.p, a.data(), sizeof(double) * a.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("unsflow_lve_create: upload failed");
        return fail(2);
    }
    const size_t dyn = sizeof(double) * nd * nd;
    // static + dynamic shared memory exceed the 48 KB default already at ndiv = 70: always opt in
    if (cudaFuncSetAttribute(k_lve_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
        set_error("unsflow_lve_create: cannot reserve %zu bytes of shared memory", dyn);
        return fail(2);
    }
    h->dev.lve = h->lves.p;
    h->dev.lvea = h->lvea.p;
    h->dev.lve_binv = h->lve_binv.p;
    return 0;
}

/* LVE extras of the state: circ (ndiv: panel circulations + shed TEV strength), circChange, prevCirc (ndiv-1 used) */
int unsflow_lve_get_circ(unsflow_ldvm *h, double *circ_ndiv, double *gamma_crit) {
    UF_CHECK(h && h->lvea.p && circ_ndiv, "unsflow_lve_get_circ: not an LVE handle");
    UF_CUDA(cudaStreamSynchronize(h->stream));
    UF_CUDA(cudaMemcpy(circ_ndiv, h->lvea.p + (int64_t)LV_CIRC * h->ndiv, sizeof(double) * h->ndiv, cudaMemcpyDeviceToHost));
    if (gamma_crit) {
        LveState L;
        UF_CUDA(cudaMemcpy(&L, h->lves.p, sizeof(L), cudaMemcpyDeviceToHost));
        *gamma_crit = L.gamma_crit;
    }
    return 0;
}

int unsflow_ldvm_set_kinematics(unsflow_ldvm *h, int64_t nrows, const double *table) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_set_kinematics: null handle");
    UF_CHECK(nrows > 0 && table != nullptr, "unsflow_ldvm_set_kinematics: empty table");
    UF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->kin.alloc(9 * nrows * h->nsurf)) return 2;
    UF_CUDA(cudaMemcpyAsync(h->kin.p, table, sizeof(double) * 9 * nrows * h->nsurf, cudaMemcpyHostToDevice, h->stream));
    // rows are consumed from the start of the new table
    StepState *zero_step = h->pinned;
    (void)zero_step;
    long long z = 0;
    UF_CUDA(cudaMemcpyAsync(&h->st.p->step, &z, sizeof(z), cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // kernel params are baked into the graph
    h->kin_rows = nrows;
    h->dev.kin = h->kin.p;
    h->dev.kin_rows = nrows;
    return 0;
}

int unsflow_ldvm_set_parallel(unsflow_ldvm *h, int rank, int nranks, const void *nccl_unique_id) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_set_parallel: null handle");
    UF_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "unsflow_ldvm_set_parallel: bad rank %d / %d", rank, nranks);
    UF_CHECK(h->steps_done == 0 && h->comm == nullptr, "unsflow_ldvm_set_parallel: call once, before the first run");
    if (nranks == 1) return 0;
    UF_CHECK(nccl_unique_id != nullptr, "unsflow_ldvm_set_parallel: nccl_unique_id required");
    UF_CHECK(h->uniform, "unsflow_ldvm_set_parallel: the multi-GPU roll-up needs a uniform core radius");
    const NcclApi *nc = nccl_api();
    if (!nc) return 4;
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    if (nc->CommInitRank(&h->comm, nranks, id, rank) != ncclSuccess) { set_error("ncclCommInitRank failed"); return 4; }
    if (h->vsum.alloc(2 * h->tot) || h->vsum.zero(h->stream)) return 2;
    UF_CUDA(cudaStreamSynchronize(h->stream));
    h->rank = rank; h->nranks = nranks;
    // NVLink peer mapping for the fused finish + exchange; when it cannot be set up on every rank the step keeps
    // the NCCL all-reduce (both are GPU paths; all ranks take the same one)
    void *bufs[kPeerBufs] = {h->vsum.p, h->wake.p};
    if (int rc = peer_setup(&h->peer, h->comm, h->stream, rank, nranks, bufs)) return rc;
    return 0;
}

int unsflow_ldvm_set_2dof(unsflow_ldvm *h, const double *sp11, const double *k22) {
    UF_CHECK(h && sp11 && k22, "unsflow_ldvm_set_2dof: null argument");
    UF_CUDA(cudaMemcpyAsync(h->st.p->sp, sp11, sizeof(double) * 11, cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaMemcpyAsync(h->st.p->k2, k22, sizeof(double) * 22, cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int unsflow_ldvm_get_2dof(unsflow_ldvm *h, double *k22) {
    UF_CHECK(h && k22, "unsflow_ldvm_get_2dof: null argument");
    UF_CUDA(cudaMemcpyAsync(k22, h->st.p->k2, sizeof(double) * 22, cudaMemcpyDeviceToHost, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// upper bounds of the live end offsets at `ahead` steps from now (known counts + steps since)
static void step_bounds(const unsflow_ldvm *h, int64_t ahead, int64_t *ntev_ub, int64_t *nlev_ub) {
    const int64_t since = (h->steps_done - h->known_step + ahead) * h->nsurf;
    *ntev_ub = std::min(h->cap_t, h->known_ntev_end + since);
    *nlev_ub = std::min(h->cap_l, h->known_nlev_end + since);
}

// UNSFLOW_STEP_TRACE=1: CUDA events between the launches of the LAST non-graph step of a run; the segment times are
// printed to stderr by unsflow_ldvm_run (names the segments of a step, also on several ranks where ncu cannot go)
static bool step_trace_enabled() {
    static bool v = [] { const char *e = getenv("UNSFLOW_STEP_TRACE"); return e && atoi(e) != 0; }();
    return v;
}
struct StepTrace {
    cudaEvent_t ev[12] = {};
    const char *name[12] = {};
    int n = 0;
    bool armed = false;
    void mark(const char *nm, cudaStream_t st) {
        if (!armed || n >= 12) return;
        if (!ev[n]) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        name[n++] = nm;
    }
};
static thread_local StepTrace g_trace;

// one time step = 5 launches on the handle's stream (also used under stream capture)
static int enqueue_step(unsflow_ldvm *h, int64_t ntev_ub, int64_t nlev_ub, cudaStream_t st) {
    const int driver = h->opts.driver, smc = sm_count(), order = rsqrt_order();
    const int64_t tot = h->tot, nwake_ub = ntev_ub + nlev_ub + h->nextv;
    const bool multi = h->nsurf > 1;
    const bool lve = driver == UNSFLOW_DRIVER_LVE_INTERNAL;
    g_trace.n = 0;
    g_trace.mark("start", st);
    if (lve) k_lve_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    else if (multi) k_mstep_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    else k_step_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    h->launches++;
    g_trace.mark("head", st);
    if (driver != UNSFLOW_DRIVER_LAUTAT) {
        // bound-point sweep (update_indbound, calcs.jl:35-38): targets = bound points
        InducePlan p2 = plan_induce((long long)h->ndiv * h->nsurf, 1, nwake_ub, 3, smc, h->S2max);
        InduceArgs a;
        a.sx = h->dev.wx; a.sz = h->dev.wz; a.sg = h->dev.wg; a.svc4 = h->dev.wvc4;
        a.tx = h->dev.bnd_x; a.tz = h->dev.bnd_z;
        a.sreg = &h->st.p->wake; a.treg = &h->st.p->bnd; a.nsreg = 3; a.ntreg = 1;
        a.vc4_uniform = h->vc4u;
        a.pu = h->p2.p; a.pw = h->p2.p + (int64_t)h->S2max * h->ndivpad; a.tstride = h->ndivpad;
        if (lve || (!multi && sweep_supported(h->ndiv))) {
            // dedicated few-targets kernel (sweep.cuh): one partial plane per block
            const int nt2 = lve ? h->ndiv - 1 : h->ndiv;
            p2.S = sweep_splits(nwake_ub, 3, h->S2max);
            if (int rc = launch_sweep(a, nt2, p2.S, h->uniform, st)) return rc;
        } else {
            if (int rc = launch_induce_direct(a, p2, h->uniform, order, st)) return rc;
        }
        h->dev.S2 = p2.S;
        h->launches++;
        g_trace.mark("sweep", st);
    }
    if (lve) k_lve_solve<<<1, kSolveThreads, sizeof(double) * h->ndiv * h->ndiv, st>>>(h->dev);
    else if (multi) k_mstep_solve<<<1, kSolveThreads, 0, st>>>(h->dev);
    else {
        // tables in shared memory when they fit beside the workspace (default 70 x 35: 40 KB)
        const size_t tab_bytes = sizeof(double) * 2 * (size_t)(h->naterm + 1) * h->ndiv;
        const size_t ws_bytes = sizeof(double) * (size_t)((solve_ws_doubles(h->ndiv, h->naterm) + 15) / 16 * 16);
        const int in_smem = tab_bytes <= 160 * 1024;
        const size_t dyn = ws_bytes + (in_smem ? tab_bytes : 0);
        static size_t attr_bytes[64] = {0};      // per device; the attribute only ever needs to grow
        int dev = 0;
        UF_CUDA(cudaGetDevice(&dev));
        if (dyn > attr_bytes[dev & 63]) {
            UF_CUDA(cudaFuncSetAttribute(k_step_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            attr_bytes[dev & 63] = dyn;
        }
        k_step_solve<<<1, kSolveThreads, dyn, st>>>(h->dev, in_smem);
    }
    h->launches++;
    g_trace.mark("solve", st);
    // wake roll-up (calcs.jl:222-294): targets = free vortices, sources = free + bound
    const bool sym = h->uniform && nwake_ub >= sym_threshold();
    InducePlan p1 = plan_induce(nwake_ub, 3, nwake_ub + (long long)h->nsurf * (h->ndiv - 1), 4, smc, h->S1max);
    InduceArgs a;
    a.sx = h->dev.wx; a.sz = h->dev.wz; a.sg = h->dev.wg; a.svc4 = h->dev.wvc4;
    a.tx = h->dev.wx; a.tz = h->dev.wz;
    a.sreg = &h->st.p->wake; a.treg = &h->st.p->wake; a.nsreg = 4; a.ntreg = 3;
    a.vc4_uniform = h->vc4u;
    a.pu = h->p1.p; a.pw = h->p1.p + (int64_t)h->S1max * tot; a.tstride = tot;
    if (sym) {
        const int symR = sym_rows_per_thread(nwake_ub, h->nranks);
        UF_CHECK(h->planes.p != nullptr, "internal: pair-tile planes not allocated before the step was enqueued");
        h->ctas = sym_ctas(symR);
        SymArgs sa;
        sa.x = h->dev.wx; sa.z = h->dev.wz; sa.g = h->dev.wg;
        sa.reg = &h->st.p->wake; sa.nreg = 3; sa.bv_reg = 3;
        sa.vc4 = h->vc4u; sa.P = h->planes.p; sa.n = tot; sa.rank = h->rank; sa.nranks = h->nranks;
        if (int rc = launch_induce_sym(sa, symR, order, st)) return rc;
        g_trace.mark("pair tiles", st);
        // sum of the planes that touched each block (+ 1/2pi, free stream, explicit Euler when single-rank)
        SymFinishArgs fs;
        memset(&fs, 0, sizeof(fs));
        fs.P = h->planes.p; fs.n = tot; fs.ctas = h->ctas; fs.R = symR;
        fs.reg = &h->st.p->wake; fs.nreg = 3;
        fs.vx = h->dev.wvx; fs.vz = h->dev.wvz; fs.x = h->dev.wx; fs.z = h->dev.wz;
        fs.fs = h->st.p->fs; fs.dt = h->opts.dt;
        if (h->nranks == 1) {
            if (int rc = launch_sym_finish(fs, nwake_ub, st)) return rc;
            g_trace.mark("finish", st);
            h->launches += 2;
            return 0;
        }
        // every rank holds the partial velocities of ITS share of the pair tiles for all vortices:
        // plain plane sums, packed [u | w] by live-vortex ordinal -> ONE NCCL all-reduce -> 1/2pi, free stream,
        // convection.  The all-reduce returns identical bits on every rank, so the replicated state stays in lockstep.
        fs.raw = 1; fs.packed = 1; fs.packed_stride = nwake_ub; fs.vx = h->vsum.p;
        if (int rc = launch_sym_finish(fs, nwake_ub, st)) return rc;
        g_trace.mark("plane sums", st);
        if (h->peer.ready) {
            // fused finish + exchange over NVLink peer memory: barrier (all partial sums written) -> every rank sums,
            // convects and broadcasts ITS slice of the vortices into all ranks' slabs -> barrier (all slabs complete)
            if (int rc = launch_peer_barrier(&h->peer, st)) return rc;
            PeerFinishArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.reg = &h->st.p->wake; pa.nreg = 3; pa.rank = h->rank; pa.nranks = h->nranks; pa.stride = nwake_ub;
            for (int q = 0; q < h->nranks; q++) { pa.sums[q] = (const double *)h->peer.bufs[q][0]; pa.slab[q] = (double *)h->peer.bufs[q][1]; }
            pa.off_x = 0; pa.off_z = tot; pa.off_vx = 5 * tot; pa.off_vz = 6 * tot;
            pa.fs = h->st.p->fs; pa.dt = h->opts.dt;
            if (int rc = launch_peer_finish(pa, nwake_ub, st)) return rc;
            if (int rc = launch_peer_barrier(&h->peer, st)) return rc;
            g_trace.mark("peer exchange", st);
            h->launches += 5;
            return 0;
        }
        const NcclApi *nc = nccl_api();
        if (nc->AllReduce(h->vsum.p, h->vsum.p, (size_t)(2 * nwake_ub), ncclDouble, ncclSum, h->comm, st) != ncclSuccess) { set_error("ncclAllReduce failed"); return 4; }
        FinishArgs f2;
        memset(&f2, 0, sizeof(f2));
        f2.pu = h->vsum.p; f2.pw = h->vsum.p + nwake_ub; f2.tstride = 0; f2.S = 1; f2.packed = 1;
        f2.treg = &h->st.p->wake; f2.ntreg = 3;
        f2.vx = h->dev.wvx; f2.vz = h->dev.wvz; f2.x = h->dev.wx; f2.z = h->dev.wz;
        f2.fs = h->st.p->fs; f2.dt = h->opts.dt;
        if (int rc = launch_induce_finish(f2, nwake_ub, st)) return rc;
        g_trace.mark("all-reduce + finish", st);
        h->launches += 3;
        return 0;
    }
    if (int rc = launch_induce_direct(a, p1, h->uniform, order, st)) return rc;
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.pu = a.pu; f.pw = a.pw; f.tstride = tot; f.S = p1.S;
    f.treg = &h->st.p->wake; f.ntreg = 3;
    f.vx = h->dev.wvx; f.vz = h->dev.wvz; f.x = h->dev.wx; f.z = h->dev.wz;
    f.fs = h->st.p->fs; f.dt = h->opts.dt;
    if (int rc = launch_induce_finish(f, nwake_ub, st)) return rc;
    g_trace.mark("roll-up + finish", st);
    h->launches += 2;
    return 0;
}

// partial planes of the pair-tile kernel, allocated on first need and OUTSIDE stream capture
static int ensure_planes(unsflow_ldvm *h, int64_t nwake_ub) {
    if (h->planes.p || !h->uniform || nwake_ub < sym_threshold()) return 0;
    return h->planes.alloc(sym_plane_doubles(sym_ctas(sym_rows_per_thread(nwake_ub, h->nranks)), h->tot)) ? 2 : 0;
}

static void poll_readback(unsflow_ldvm *h) {
    if (h->readback_pending && cudaEventQuery(h->evc) == cudaSuccess) {
        h->readback_pending = false;
        h->known_step = h->readback_step;
        h->known_ntev_end = h->pinned->wake.end[0] - h->off_t;
        h->known_nlev_end = h->pinned->wake.end[1] - h->off_l;
    }
}

// Smallest wakes: chunks of whole steps inside one thread-block cluster (k_cluster_steps).  Single surface, the four
// single-surface drivers, ndiv <= 72 (9 bound points per sweep lane), one rank; UNSFLOW_CLUSTER_SIZE (default 8; 16 =
// non-portable size), UNSFLOW_CLUSTER_MAX_N (free vortices up to which the cluster path is taken), UNSFLOW_NO_CLUSTER=1.
constexpr int64_t kClusterChunk = 64;
static size_t cluster_dyn_probe() {      // shared memory of the default discretisation (cluster-size probe)
    return sizeof(double) * cluster_smem_doubles(8 * kClRT, 35, true, true);
}
// cluster size: 16 CTAs (non-portable size) when the device can co-schedule such a cluster, else 8; the env forces it
static int cluster_size() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 8;
    int &c = cached[dev & 63];
    if (c) return c;
    const char *e = getenv("UNSFLOW_CLUSTER_SIZE");
    const int want = e ? atoi(e) : 0;
    if (want == 2 || want == 4 || want == 8) return c = want;
    // 16: ask the driver whether one such cluster fits
    void (*kern)(const ClusterArgs) = k_cluster_steps<true>;
    const size_t dyn = cluster_dyn_probe();
    int n16 = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) == cudaSuccess &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(16); cfg.blockDim = dim3(kSolveThreads); cfg.dynamicSmemBytes = dyn;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n16, kern, &cfg) != cudaSuccess) n16 = 0;
    }
    cudaGetLastError();
    return c = (n16 >= 1 ? 16 : 8);
}
// free vortices (host-side upper bound at the end of a chunk) up to which the cluster path is taken: the crossover with
// the five-kernel stepper measured on a growing C2 wake (profiles/r02_cluster_stepper.md) -- the roll-up of a cluster is
// 8 or 16 SMs' worth of FP64, the five-kernel path has all 148
static int64_t cluster_max_free() {
    const char *off = getenv("UNSFLOW_NO_CLUSTER");
    if (off && atoi(off) != 0) return -1;
    const char *e = getenv("UNSFLOW_CLUSTER_MAX_N");
    const int64_t n = e ? atoll(e) : (cluster_size() >= 16 ? 1664 : 1024);
    return std::min<int64_t>(n, kClMaxFree);
}
static bool cluster_usable(const unsflow_ldvm *h) {
    const int dr = h->opts.driver;
    return !h->cluster_off && cluster_max_free() >= 0 && h->nsurf == 1 && h->nranks == 1 && h->ndiv <= 8 * kClRT &&
           (dr == UNSFLOW_DRIVER_LDVM || dr == UNSFLOW_DRIVER_LDVM2DOF || dr == UNSFLOW_DRIVER_LDVMLIN || dr == UNSFLOW_DRIVER_LAUTAT);
}
static int launch_cluster_steps(unsflow_ldvm *h, int64_t nsteps, cudaStream_t st) {
    const int C = cluster_size();
    const size_t tab_bytes = sizeof(double) * 2 * (size_t)(h->naterm + 1) * h->ndiv;
    const bool tabs = tab_bytes <= 64 * 1024;
    const size_t dyn = sizeof(double) * cluster_smem_doubles(h->ndiv, h->naterm, tabs, h->uniform);
    h->dev.S2 = C - 1;      // one bound-sweep plane per helper CTA
    ClusterArgs a;
    a.d = h->dev;
    a.planes_u = h->p2.p; a.planes_w = h->p2.p + (int64_t)h->S2max * h->ndivpad;
    a.nsteps = nsteps; a.tabs_in_smem = tabs ? 1 : 0; a.vc4u = h->vc4u; a.dt = h->opts.dt;
    static const bool no_overlap = [] { const char *e = getenv("UNSFLOW_CLUSTER_NO_OVERLAP"); return e && atoi(e) != 0; }();
    a.overlap = no_overlap ? 0 : 1;
    // UNSFLOW_CLUSTER_PROFILE=1: cycles of CTA 0 per phase, printed by unsflow_ldvm_run
    static const bool want_prof = [] { const char *e = getenv("UNSFLOW_CLUSTER_PROFILE"); return e && atoi(e) != 0; }();
    a.prof = nullptr;
    if (want_prof) {
        if (!h->clprof.p) { if (h->clprof.alloc(8) || h->clprof.zero(st)) return 2; }
        a.prof = h->clprof.p;
    }
    void (*kern)(const ClusterArgs) = h->uniform ? k_cluster_steps<true> : k_cluster_steps<false>;
    static size_t attr_bytes[64][2] = {{0}};      // per device and instantiation; the attribute only ever needs to grow
    int dev = 0;
    UF_CUDA(cudaGetDevice(&dev));
    size_t &ab = attr_bytes[dev & 63][h->uniform ? 1 : 0];
    if (dyn > ab) {
        UF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        if (C > 8) UF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        ab = dyn;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(kSolveThreads); cfg.dynamicSmemBytes = dyn; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    UF_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return 0;
}

// Launch-bound regime (small wakes): replay a CUDA graph of kGraphChunk steps whose grids are
// sized for a quantised upper bound at the END of the chunk; the graph is re-captured only when
// that quantised bound (the key) changes.
constexpr int64_t kGraphChunk = 32;
constexpr int64_t kGraphQuant = 512;
static bool graphs_enabled() {
    static bool v = [] { const char *e = getenv("UNSFLOW_NO_GRAPH"); return !(e && atoi(e) != 0); }();
    return v;
}

int unsflow_ldvm_run(unsflow_ldvm *h, int64_t nsteps, double *mat_out) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_run: null handle");
    UF_CHECK(nsteps >= 0, "unsflow_ldvm_run: negative nsteps");
    UF_CHECK(nsteps == 0 || mat_out != nullptr, "unsflow_ldvm_run: null mat_out");
    UF_CHECK(h->steps_done + nsteps <= h->opts.max_steps, "unsflow_ldvm_run: %lld steps exceed opts.max_steps=%lld",
             (long long)(h->steps_done + nsteps), (long long)h->opts.max_steps);
    UF_CHECK(h->dev.kin != nullptr, "unsflow_ldvm_run: call unsflow_ldvm_set_kinematics first");
    if (nsteps == 0) return 0;
    cudaStream_t st = h->stream;
    // mat rows of this call start at row 0 of the device buffer
    long long z = 0;
    UF_CUDA(cudaMemcpyAsync(&h->st.p->mat_row, &z, sizeof(z), cudaMemcpyHostToDevice, st));
    h->launches = 0;
    UF_CUDA(cudaEventRecord(h->ev0, st));
    int64_t is = 0;
    while (is < nsteps) {
        poll_readback(h);
        int64_t tev_ub, lev_ub;
        step_bounds(h, kGraphChunk, &tev_ub, &lev_ub);
        const bool small = tev_ub + lev_ub + h->nextv < 16384;
        bool in_cluster = false;
        if (cluster_usable(h)) {
            const int64_t chunk = std::min<int64_t>(nsteps - is, kClusterChunk);
            int64_t ct_ub, cl_ub;
            step_bounds(h, chunk, &ct_ub, &cl_ub);
            in_cluster = ct_ub + cl_ub + h->nextv <= cluster_max_free();
            if (in_cluster) {
                if (launch_cluster_steps(h, chunk, st) != 0) {
                    // the device would not take the cluster launch (no room for a co-scheduled cluster, e.g. a partitioned
                    // GPU): nothing was enqueued -- this handle keeps to the five-kernel GPU path from here on
                    cudaGetLastError();
                    h->cluster_off = true;
                    in_cluster = false;
                } else {
                    h->launches += 1;
                    h->steps_done += chunk;
                    is += chunk;
                }
            }
        }
        if (in_cluster) {
        } else if (graphs_enabled() && small && nsteps - is >= kGraphChunk) {
            const int64_t qt = std::min(h->cap_t, round_up(tev_ub, kGraphQuant)), ql = std::min(h->cap_l, round_up(lev_ub, kGraphQuant));
            if (int rc = ensure_planes(h, qt + ql + h->nextv)) return rc;
            if (!h->gexec || h->gkey_t != qt || h->gkey_l != ql || h->gkey_kin != h->dev.kin) {
                if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
                cudaGraph_t g = nullptr;
                const int64_t l0 = h->launches;
                UF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int64_t k = 0; k < kGraphChunk && rc == 0; k++) rc = enqueue_step(h, qt, ql, st);
                cudaError_t ce = cudaStreamEndCapture(st, &g);
                h->launches = l0;
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                UF_CUDA(ce);
                UF_CUDA(cudaGraphInstantiate(&h->gexec, g, 0));
                cudaGraphDestroy(g);
                h->gkey_t = qt; h->gkey_l = ql; h->gkey_kin = h->dev.kin;
            }
            UF_CUDA(cudaGraphLaunch(h->gexec, st));
            h->launches += kGraphChunk * (h->opts.driver == UNSFLOW_DRIVER_LAUTAT ? 4 : 5);
            h->steps_done += kGraphChunk;
            is += kGraphChunk;
        } else {
            step_bounds(h, 1, &tev_ub, &lev_ub);
            if (int rc = ensure_planes(h, tev_ub + lev_ub + h->nextv)) return rc;
            g_trace.armed = step_trace_enabled() && is == nsteps - 1;
            if (int rc = enqueue_step(h, tev_ub, lev_ub, st)) return rc;
            g_trace.armed = false;
            UF_CUDA(cudaGetLastError());
            h->steps_done++;
            is++;
        }
        if (h->nranks > 1) {
            // grid sizes must be identical on all ranks (they fix the grouping of partial sums, hence the
            // bits): no timing-dependent read-back; a synchronous one at fixed step counts instead
            if (h->steps_done - h->readback_step >= 256) {
                StepState sx;
                if (int rc = ldvm_sync_counts(h, &sx)) return rc;
                h->known_step = h->readback_step = h->steps_done;
                h->known_ntev_end = sx.wake.end[0] - h->off_t;
                h->known_nlev_end = sx.wake.end[1] - h->off_l;
            }
        } else if (!h->readback_pending && h->steps_done - h->readback_step >= 64) {
            UF_CUDA(cudaMemcpyAsync(h->pinned, h->st.p, sizeof(StepState), cudaMemcpyDeviceToHost, st));
            UF_CUDA(cudaEventRecord(h->evc, st));
            h->readback_pending = true;
            h->readback_step = h->steps_done;
        }
    }
    UF_CUDA(cudaEventRecord(h->ev1, st));
    const int64_t tot = h->tot;
    (void)tot;
    // mat: device row-major [nsteps][8] -> caller column-major nsteps x 8
    const int mc = h->mat_cols;
    std::vector<double> hm((size_t)(mc * nsteps));
    UF_CUDA(cudaMemcpyAsync(hm.data(), h->mat.p, sizeof(double) * mc * nsteps, cudaMemcpyDeviceToHost, st));
    UF_CUDA(cudaStreamSynchronize(st));
    h->readback_pending = false;
    for (int64_t i = 0; i < nsteps; i++)
        for (int j = 0; j < mc; j++) mat_out[i + j * nsteps] = hm[(size_t)(mc * i + j)];
    UF_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (int rc = peer_check(&h->peer, st)) return rc;
    if (h->clprof.p) {
        long long pc[5] = {0, 0, 0, 0, 0};
        UF_CUDA(cudaMemcpy(pc, h->clprof.p, sizeof(pc), cudaMemcpyDeviceToHost));
        if (pc[4] > 0) {
            int khz = 0;
            cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
            const double us = 1e3 / (double)khz / (double)pc[4];
            fprintf(stderr, "[unsflow cluster profile] %lld steps; us per step on CTA 0: sweep | place_tev %.2f, solve | early roll-up %.2f, late roll-up | next head %.2f\n",
                    pc[4], pc[1] * us, pc[2] * us, pc[3] * us);
        }
        UF_CUDA(cudaMemset(h->clprof.p, 0, sizeof(pc)));
    }
    if (step_trace_enabled() && g_trace.n > 1) {
        fprintf(stderr, "[unsflow step trace] rank %d/%d last step:", h->rank, h->nranks);
        for (int k = 1; k < g_trace.n; k++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, g_trace.ev[k - 1], g_trace.ev[k]);
            fprintf(stderr, " %s %.1f us;", g_trace.name[k], 1e3 * ms);
        }
        fprintf(stderr, "\n");
        g_trace.n = 0;
    }
    // exact counts are known now
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    h->known_step = h->steps_done;
    h->known_ntev_end = s.wake.end[0] - h->off_t;
    h->known_nlev_end = s.wake.end[1] - h->off_l;
    return 0;
}

int unsflow_ldvm_counts(unsflow_ldvm *h, int64_t *ntev, int64_t *nlev, int64_t *nextv) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_counts: null handle");
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    if (ntev) *ntev = s.wake.end[0] - s.wake.begin[0];
    if (nlev) *nlev = s.wake.end[1] - s.wake.begin[1];
    if (nextv) *nextv = s.wake.end[2] - s.wake.begin[2];
    return 0;
}

static int ldvm_get_state_impl(unsflow_ldvm *h, int nsurf, unsflow_surf *sfs, unsflow_field *fl) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_get_state: null handle");
    UF_CHECK(sfs == nullptr || nsurf == h->nsurf, "unsflow_ldvm_get_state: handle has %d surface(s), caller passed %d", h->nsurf, nsurf);
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    cudaStream_t st = h->stream;
    const int nd = h->ndiv, na = h->naterm, ns = h->nsurf;
    const int64_t tot = h->tot;
    auto get = [&](unsflow_vorts *v, int64_t b, int64_t n) -> int {
        if (!v) return 0;
        UF_CHECK(v->n >= n, "unsflow_ldvm_get_state: vortex array capacity %lld < count %lld", (long long)v->n, (long long)n);
        double *dst[6] = {v->x, v->z, v->s, v->vc, v->vx, v->vz};
        const int slab[6] = {0, 1, 2, 4, 5, 6};
        for (int k = 0; k < 6; k++)
            if (dst[k] && n > 0)
                UF_CUDA(cudaMemcpyAsync(dst[k], h->wake.p + slab[k] * tot + b, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        v->n = n;
        return 0;
    };
    if (sfs) {
        std::vector<double> hs((size_t)ns * (9 * nd + 3 * na));
        std::vector<SurfState> hms((size_t)ns);
        UF_CUDA(cudaMemcpyAsync(hs.data(), h->surf.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, st));
        UF_CUDA(cudaMemcpyAsync(hms.data(), h->ms.p, sizeof(SurfState) * ns, cudaMemcpyDeviceToHost, st));
        UF_CUDA(cudaStreamSynchronize(st));
        for (int q = 0; q < ns; q++) {
            unsflow_surf *sf = &sfs[q];
            UF_CHECK(sf->ndiv == nd && sf->naterm == na, "unsflow_ldvm_get_state: surf dimensions differ from the handle's");
            double *sarr[9] = {nullptr, nullptr, nullptr, nullptr, sf->bnd_x, sf->bnd_z, sf->uind, sf->wind, sf->downwash};
            for (int k = 4; k < 9; k++) if (sarr[k]) memcpy(sarr[k], &hs[((size_t)k * ns + q) * nd], sizeof(double) * nd);
            if (sf->aterm) memcpy(sf->aterm, &hs[(size_t)9 * ns * nd + ((size_t)0 * ns + q) * na], sizeof(double) * na);
            if (sf->adot) memcpy(sf->adot, &hs[(size_t)9 * ns * nd + ((size_t)1 * ns + q) * na], sizeof(double) * na);
            if (sf->aprev) memcpy(sf->aprev, &hs[(size_t)9 * ns * nd + ((size_t)2 * ns + q) * na], sizeof(double) * na);
            if (ns == 1) {
                sf->levflag = s.levflag;
                for (int k = 0; k < 6; k++) sf->kinem[k] = s.kin[k];
                sf->a0 = s.a0; sf->a0dot = s.a0dot; sf->a0prev = s.a0prev;
            } else {
                sf->levflag = hms[q].levflag;
                for (int k = 0; k < 6; k++) sf->kinem[k] = hms[q].kin[k];
                sf->a0 = hms[q].a0; sf->a0dot = hms[q].a0dot; sf->a0prev = hms[q].a0prev;
            }
            if (int rc = get(&sf->bv, s.wake.begin[3] + (int64_t)q * (nd - 1), nd - 1)) return rc;
        }
    }
    if (fl) {
        for (int r = 0; r < 3; r++) {
            unsflow_vorts *v = r == 0 ? &fl->tev : (r == 1 ? &fl->lev : &fl->extv);
            if (int rc = get(v, s.wake.begin[r], s.wake.end[r] - s.wake.begin[r])) return rc;
        }
        fl->u = s.fs[0]; fl->w = s.fs[1];
    }
    UF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int unsflow_ldvm_get_state(unsflow_ldvm *h, unsflow_surf *sf, unsflow_field *fl) { return ldvm_get_state_impl(h, 1, sf, fl); }

int unsflow_ldvm_multi_get_state(unsflow_ldvm *h, int32_t nsurf, unsflow_surf *surfs, unsflow_field *fl) {
    return ldvm_get_state_impl(h, nsurf, surfs, fl);
}

int unsflow_ldvm_parallel_info(unsflow_ldvm *h, int *rank, int *nranks, int *peer_exchange) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_parallel_info: null handle");
    if (rank) *rank = h->rank;
    if (nranks) *nranks = h->nranks;
    if (peer_exchange) *peer_exchange = h->peer.ready ? 1 : 0;
    return 0;
}

int unsflow_ldvm_timing(unsflow_ldvm *h, float *ms_total, int64_t *n_launches) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_timing: null handle");
    if (ms_total) *ms_total = h->last_ms;
    if (n_launches) *n_launches = h->launches;
    return 0;
}

}  // extern "C"

// shared body of the two ensemble entry points; strpar/kinem2dof non-null <=> ldvm2DOF members
static thread_local std::vector<long long> g_batch_prof;   // per-CTA phase cycle sums of the last run (UNSFLOW_ENS_PROFILE=1)
static int ens_auto_threads(int64_t M) {
    return 256;
}
static int batch_run_impl(const unsflow_surf *sf, const unsflow_opts *op, int64_t nmember, int64_t nsteps,
                          const double *tables, const double *lespcrit, const double *strpar, const double *kinem2dof,
                          double *mat_out, double *kinem2dof_out, int64_t *ntev_out, int64_t *nlev_out) {
    UF_CHECK(sf && op && tables && mat_out, "unsflow_ldvm_batch_run: null argument");
    UF_CHECK(nmember > 0 && nsteps > 0, "unsflow_ldvm_batch_run: need nmember > 0 and nsteps > 0");
    UF_CHECK(sf->ndiv >= 3 && sf->ndiv <= kNdMax && sf->naterm >= 3 && sf->naterm <= kNaMax, "unsflow_ldvm_batch_run: bad ndiv/naterm");
    if (strpar) {
        UF_CHECK(op->driver == UNSFLOW_DRIVER_LDVM2DOF && kinem2dof, "unsflow_ldvm2dof_batch_run: driver must be UNSFLOW_DRIVER_LDVM2DOF and kinem2dof given");
    } else {
        UF_CHECK(op->driver == UNSFLOW_DRIVER_LDVM || op->driver == UNSFLOW_DRIVER_LDVMLIN || op->driver == UNSFLOW_DRIVER_LAUTAT,
                 "unsflow_ldvm_batch_run: prescribed-motion drivers only (ldvm, ldvmLin, lautat); use unsflow_ldvm2dof_batch_run for ldvm2DOF");
    }
    UF_CHECK(op->dt > 0, "unsflow_ldvm_batch_run: dt must be > 0");
    UF_CHECK(op->solve_mode == 0 || op->solve_mode == 1, "unsflow_ldvm_batch_run: bad solve_mode %d", op->solve_mode);
    UF_CHECK(sf->bv.n == sf->ndiv - 1, "unsflow_ldvm_batch_run: surf.bv must hold ndiv-1 vortices");
    UF_CHECK(sf->theta && sf->x && sf->cam && sf->cam_slope && sf->bnd_x && sf->bnd_z && sf->uind && sf->wind && sf->downwash &&
                 sf->aterm && sf->adot && sf->aprev && sf->bv.x && sf->bv.z && sf->bv.s && sf->bv.vc,
             "unsflow_ldvm_batch_run: null surf array");
    const int nd = sf->ndiv, na = sf->naterm;
    const int64_t M = nmember;
    // per-member slabs (shared memory): [TEV | LEV | EXTV(empty) | BV]
    const int64_t cap_t = round_up(nsteps + 1, 8), cap_l = round_up(nsteps + 1, 8), cap_e = 8, bvp = round_up(nd - 1, 8);
    const int64_t off_l = cap_t, off_e = off_l + cap_l, off_b = off_e + cap_e, tot = off_b + bvp;
    // member CTA width: 256 threads x 3 per SM when there are members enough to fill the GPU several times over; wider,
    // fewer CTAs when a GPU only holds a few hundred members (the tail of the longest member would dominate)
    static const int env_threads = [] { const char *e = getenv("UNSFLOW_ENS_THREADS"); return e ? atoi(e) : 0; }();
    const int threads = env_threads == 512 || env_threads == 1024 || env_threads == 256 ? env_threads : ens_auto_threads(M);
    const int groups = std::max(1, threads / nd);      // source groups of ens_sweep = partial planes it writes
    const size_t ens_dyn = sizeof(double) * (size_t)ens_smem_doubles(nd, na, tot, groups);
    UF_CHECK(ens_dyn <= 200 * 1024, "unsflow_ldvm_batch_run: %lld steps need %zu bytes of shared memory per member (max 200 KB); use unsflow_ldvm_run", (long long)nsteps, ens_dyn);
    // the member kernels use the scalar core radius 0.02c of every reference-created vortex (calcs.jl:385,423)
    for (int64_t i = 0; i < nd - 1; i++)
        UF_CHECK(sf->bv.vc[i] == 0.02 * sf->c, "unsflow_ldvm_batch_run: surf.bv[%lld].vc must be 0.02c (uniform core radius)", (long long)i);
    if (int rc = require_device()) return rc;
    if (op->device >= 0) UF_CUDA(cudaSetDevice(op->device));
    const int64_t ngeo = 4 * nd, nimg = 5 * nd + 3 * na + 3 * (nd - 1);
    DevBuf<StepState> st;
    DevBuf<double> geo, img, tabs, kin, mat;
    DevBuf<int> order, next;
    DevBuf<long long> prof;
    DevEvent ev0, ev1;
    if (st.alloc(M) || geo.alloc(ngeo) || img.alloc(nimg) || tabs.alloc(2 * (na + 1) * nd) || kin.alloc(M * nsteps * 9) ||
        mat.alloc(M * nsteps * 8) || order.alloc(M) || next.alloc(1) || ev0.create() || ev1.create())
        return 2;
    cudaStream_t stream = 0;
    // ---- host staging --------------------------------------------------------------------------
    std::vector<double> hg((size_t)ngeo), hi((size_t)nimg);
    const double *garr[4] = {sf->theta, sf->x, sf->cam, sf->cam_slope};
    for (int k = 0; k < 4; k++) memcpy(&hg[(size_t)k * nd], garr[k], sizeof(double) * nd);
    const double *sarr[5] = {sf->bnd_x, sf->bnd_z, sf->uind, sf->wind, sf->downwash};
    for (int k = 0; k < 5; k++) memcpy(&hi[(size_t)k * nd], sarr[k], sizeof(double) * nd);
    memcpy(&hi[(size_t)5 * nd], sf->aterm, sizeof(double) * na);
    memcpy(&hi[(size_t)5 * nd + na], sf->adot, sizeof(double) * na);
    memcpy(&hi[(size_t)5 * nd + 2 * na], sf->aprev, sizeof(double) * na);
    memcpy(&hi[(size_t)5 * nd + 3 * na], sf->bv.x, sizeof(double) * (nd - 1));
    memcpy(&hi[(size_t)5 * nd + 3 * na + (nd - 1)], sf->bv.z, sizeof(double) * (nd - 1));
    memcpy(&hi[(size_t)5 * nd + 3 * na + 2 * (nd - 1)], sf->bv.s, sizeof(double) * (nd - 1));
    std::vector<double> ht((size_t)(2 * (na + 1) * nd));
    for (int n = 0; n <= na; n++)
        for (int i = 0; i < nd; i++) {
            ht[(size_t)n * nd + i] = n == 0 ? 1.0 : std::cos(n * sf->theta[i]);
            ht[(size_t)(na + 1) * nd + (size_t)n * nd + i] = std::sin(n * sf->theta[i]);
        }
    const double vc_new = 0.02 * sf->c, vc4 = vc_new * vc_new * vc_new * vc_new;
    std::vector<StepState> hst((size_t)M);
    StepState s0;
    memset(&s0, 0, sizeof(s0));
    s0.wake.begin[0] = 0; s0.wake.end[0] = 0;
    s0.wake.begin[1] = off_l; s0.wake.end[1] = off_l;
    s0.wake.begin[2] = off_e; s0.wake.end[2] = off_e;
    s0.wake.begin[3] = off_b; s0.wake.end[3] = off_b + nd - 1;
    s0.bnd.begin[0] = 0; s0.bnd.end[0] = nd;
    s0.levflag = sf->levflag;
    for (int k = 0; k < 6; k++) s0.kin[k] = sf->kinem[k];
    s0.a0 = sf->a0; s0.a0dot = sf->a0dot; s0.a0prev = sf->a0prev;
    for (int64_t m = 0; m < M; m++) {
        hst[m] = s0;
        if (strpar) {
            memcpy(hst[m].sp, strpar + 11 * m, sizeof(double) * 11);
            memcpy(hst[m].k2, kinem2dof + 22 * m, sizeof(double) * 22);
        }
    }
    // Queue order: most expensive members first.  A member's cost is dominated by its wake size, i.e. by whether it
    // sheds LEVs every step: rank by the largest quasi-steady leading-edge loading the table asks for, relative to the
    // member's LESP threshold (a heuristic -- the order only affects load balance, never results).
    std::vector<int> hord((size_t)M);
    {
        std::vector<double> score((size_t)M, 0.0);
        for (int64_t m = 0; m < M; m++) {
            const double *tb = tables + m * nsteps * 9;
            double peak = 0.0;
            for (int64_t k = 0; k < nsteps; k++) {
                const double *r = tb + 9 * k;
                const double u = std::fabs(r[5]) > 1e-12 ? std::fabs(r[5]) : 1.0;
                peak = std::max(peak, std::fabs(r[1]) + std::fabs(r[3]) * sf->c / u + std::fabs(r[4]) / u);
            }
            const double lc = lespcrit ? lespcrit[m] : sf->lespcrit;
            score[m] = peak / std::max(lc, 1e-6);
        }
        for (int64_t m = 0; m < M; m++) hord[m] = (int)m;
        std::stable_sort(hord.begin(), hord.end(), [&](int x, int y) { return score[x] > score[y]; });
    }
    EnsArgs ea;
    memset(&ea, 0, sizeof(ea));
    LdvmDev &d = ea.proto;
    d.ndiv = nd; d.naterm = na; d.driver = op->driver; d.solve_mode = op->solve_mode; d.del_kind = op->delvort_kind;
    d.del_limit = op->del_limit; d.del_dist = op->del_dist; d.del_tol = op->del_tol;
    d.c = sf->c; d.uref = sf->uref; d.pvt = sf->pvt; d.lespcrit = sf->lespcrit; d.dt = op->dt;
    d.vc_new = vc_new; d.vc4_new = vc4; d.nsurf = 1;
    d.theta = geo.p; d.x = geo.p + nd; d.cam = geo.p + 2 * nd; d.cam_slope = geo.p + 3 * nd;
    d.costab = tabs.p; d.sintab = tabs.p + (int64_t)(na + 1) * nd;
    ea.img = img.p; ea.st = st.p; ea.kin = kin.p; ea.mat = mat.p; ea.order = order.p;
    ea.M = M; ea.nsteps = nsteps; ea.tot = tot; ea.off_l = off_l; ea.off_e = off_e; ea.off_b = off_b; ea.groups = groups;
    DevBuf<double> lesp;
    if (lespcrit) {
        if (lesp.alloc(M)) return 2;
        UF_CUDA(cudaMemcpyAsync(lesp.p, lespcrit, sizeof(double) * M, cudaMemcpyHostToDevice, stream));
        ea.lespcrit = lesp.p;
    }
    UF_CUDA(cudaMemcpyAsync(st.p, hst.data(), sizeof(StepState) * M, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(geo.p, hg.data(), sizeof(double) * ngeo, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(img.p, hi.data(), sizeof(double) * nimg, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(tabs.p, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(kin.p, tables, sizeof(double) * M * nsteps * 9, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(order.p, hord.data(), sizeof(int) * M, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemsetAsync(next.p, 0, sizeof(int), stream));
    ea.next = next.p;
    // UNSFLOW_ENS_SYM: 0 = one-sided roll-up, r = pair-symmetric with up to r rows per lane; UNSFLOW_ENS_OCC: CTAs per SM
    static const int want_sym = [] { const char *e = getenv("UNSFLOW_ENS_SYM"); return e ? atoi(e) : 0; }();
    // 3 member CTAs per SM (85 registers, 4 source chains per thread) give the best throughput; when the GPU holds
    // fewer than ~1.5 members per such slot the run time is the LONGEST member's, and 2 CTAs per SM (128 registers,
    // 8 chains) finish a single member sooner (512 members: 72.9 vs 81.5 ms, profiles/r02_ensemble_few_members.md)
    static const int env_occ = [] { const char *e = getenv("UNSFLOW_ENS_OCC"); return e ? atoi(e) : 0; }();
    const int want_occ = env_occ == 2 || env_occ == 3 ? env_occ : (M <= (int64_t)sm_count() * 9 / 2 ? 2 : 3);
    static const int sym_min = [] { const char *e = getenv("UNSFLOW_ENS_SYM_MIN"); return e ? atoi(e) : kEnsSymMin; }();
    ea.sym_min = std::max(sym_min, 64);
    void (*kern)(const EnsArgs) = nullptr;
    if (threads == 256 && want_occ == 2) kern = want_sym == 0 ? k_ensemble<256, 2, 8, 0> : want_sym == 2 ? k_ensemble<256, 2, 8, 2> : want_sym == 3 ? k_ensemble<256, 2, 8, 3> : k_ensemble<256, 2, 8, 4>;
    else if (threads == 256) kern = want_sym == 0 ? k_ensemble<256, 3, 4, 0> : want_sym == 1 ? k_ensemble<256, 3, 4, 1> : want_sym == 2 ? k_ensemble<256, 3, 4, 2> : k_ensemble<256, 3, 4, 3>;
    else if (threads == 512) kern = want_sym ? k_ensemble<512, 1, 4, 4> : k_ensemble<512, 1, 4, 0>;
    else kern = want_sym ? k_ensemble<1024, 1, 4, 2> : k_ensemble<1024, 1, 4, 0>;
    UF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ens_dyn));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, ens_dyn) != cudaSuccess || occ < 1) occ = 1;
    static const int cap_occ = [] { const char *e = getenv("UNSFLOW_ENS_MAXOCC"); return e ? atoi(e) : 0; }();
    if (cap_occ > 0) occ = std::min(occ, cap_occ);
    const int grid = (int)std::min<int64_t>(M, (int64_t)occ * sm_count());
    static const bool want_prof = [] { const char *e = getenv("UNSFLOW_ENS_PROFILE"); return e && atoi(e) != 0; }();
    if (want_prof) {
        if (prof.alloc((int64_t)grid * 8) || prof.zero(stream)) return 2;
        ea.prof = prof.p;
    }
    UF_CUDA(cudaEventRecord(ev0.e, stream));
    kern<<<grid, threads, ens_dyn, stream>>>(ea);
    UF_CUDA(cudaGetLastError());
    UF_CUDA(cudaEventRecord(ev1.e, stream));
    std::vector<double> hm((size_t)(M * nsteps * 8));
    UF_CUDA(cudaMemcpyAsync(hm.data(), mat.p, sizeof(double) * hm.size(), cudaMemcpyDeviceToHost, stream));
    UF_CUDA(cudaMemcpyAsync(hst.data(), st.p, sizeof(StepState) * M, cudaMemcpyDeviceToHost, stream));
    UF_CUDA(cudaStreamSynchronize(stream));
    UF_CUDA(cudaEventElapsedTime(&g_batch_ms, ev0.e, ev1.e));
    if (want_prof) {
        g_batch_prof.assign((size_t)grid * 8, 0);
        UF_CUDA(cudaMemcpy(g_batch_prof.data(), prof.p, sizeof(long long) * grid * 8, cudaMemcpyDeviceToHost));
        long long tot_c[6] = {0, 0, 0, 0, 0, 0};
        for (int c = 0; c < grid; c++)
            for (int q = 0; q < 6; q++) tot_c[q] += g_batch_prof[(size_t)c * 8 + q];
        fprintf(stderr, "[unsflow ensemble profile] grid %d (occupancy %d/SM), %.3f ms; cycles per CTA (mean): head %.3e sweep %.3e solve %.3e rollup %.3e setup %.3e; roll-up interactions %.4e total = %.3f per CTA-cycle in the roll-up phase\n",
                grid, occ, g_batch_ms, (double)tot_c[0] / grid, (double)tot_c[1] / grid, (double)tot_c[2] / grid, (double)tot_c[3] / grid, (double)tot_c[4] / grid,
                (double)tot_c[5], (double)tot_c[5] / (double)tot_c[3]);
    }
    for (int64_t m = 0; m < M; m++) {
        double *mo = mat_out + m * nsteps * 8;
        const double *mi = &hm[(size_t)(m * nsteps * 8)];
        for (int64_t i = 0; i < nsteps; i++)
            for (int j = 0; j < 8; j++) mo[i + j * nsteps] = mi[8 * i + j];
        if (ntev_out) ntev_out[m] = hst[m].wake.end[0] - hst[m].wake.begin[0];
        if (nlev_out) nlev_out[m] = hst[m].wake.end[1] - hst[m].wake.begin[1];
        if (kinem2dof_out) memcpy(kinem2dof_out + 22 * m, hst[m].k2, sizeof(double) * 22);
    }
    return 0;
}

extern "C" {

int unsflow_ldvm_batch_run(const unsflow_surf *sf, const unsflow_opts *op, int64_t nmember, int64_t nsteps,
                           const double *tables, const double *lespcrit, double *mat_out, int64_t *ntev_out,
                           int64_t *nlev_out) {
    return batch_run_impl(sf, op, nmember, nsteps, tables, lespcrit, nullptr, nullptr, mat_out, nullptr, ntev_out, nlev_out);
}

int unsflow_ldvm2dof_batch_run(const unsflow_surf *sf, const unsflow_opts *op, int64_t nmember, int64_t nsteps,
                               const double *tables, const double *lespcrit, const double *strpar11, const double *kinem22,
                               double *mat_out, double *kinem22_out, int64_t *ntev_out, int64_t *nlev_out) {
    UF_CHECK(strpar11 && kinem22, "unsflow_ldvm2dof_batch_run: null strpar / kinem");
    return batch_run_impl(sf, op, nmember, nsteps, tables, lespcrit, strpar11, kinem22, mat_out, kinem22_out, ntev_out, nlev_out);
}

int unsflow_ldvm_batch_timing(float *ms_kernel) {
    UF_CHECK(ms_kernel != nullptr, "unsflow_ldvm_batch_timing: null argument");
    *ms_kernel = g_batch_ms;
    return 0;
}

}  // extern "C"
