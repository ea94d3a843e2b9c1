// induce.cu -- launch plumbing for the induction kernels (see induce.cuh).
#define UNSFLOW_INDUCE_KERNELS
#include "induce_sym.cuh"
#include "sweep.cuh"
#include "runtime.h"
#include <algorithm>
#include <cstdlib>

#include <cstdio>
namespace unsflow {

InducePlan plan_induce(long long nt_ub, int ntreg, long long ns_ub, int nsreg, int sm_count, int max_S) {
    InducePlan p;
    const long long slots = (long long)sm_count * 2;
    auto tsrc_of = [&](int ts) { return std::max<long long>(1, (ns_ub + ts - 1) / ts + nsreg); };
    auto tiles = [&](int R) { return (nt_ub + (long long)kThreads * R - 1) / ((long long)kThreads * R) + ntreg; };
    int R = 1;
    if (tiles(4) * tsrc_of(kTS) >= 4 * slots) R = 4;
    else if (tiles(2) * tsrc_of(kTS) >= 4 * slots) R = 2;
    p.R = R;
    // latency regime: few targets (bound sweep) or a small wake -> 64-source tiles, more CTAs
    p.ts = (R == 1 && (nt_ub <= kThreads || ns_ub <= 16384)) ? 64 : kTS;
    long long mult = 8;
    // experiments: UNSFLOW_PLAN="R,ts,mult" forces the tiling of the all-pairs launches (nt > 256)
    static int forced[3] = {0, 0, 0};
    static bool parsed = [] { const char *e = getenv("UNSFLOW_PLAN"); if (e) sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]); return true; }();
    (void)parsed;
    if (forced[0] && nt_ub > kThreads) { R = p.R = forced[0]; p.ts = forced[1] == 64 && R == 1 ? 64 : kTS; if (forced[2]) mult = forced[2]; }
    const long long tsrc = tsrc_of(p.ts);
    p.ntiles = (int)std::max<long long>(1, tiles(R));
    long long S = (mult * slots + p.ntiles - 1) / p.ntiles;
    S = std::min<long long>(S, tsrc);
    S = std::min<long long>(S, max_S);
    S = std::min<long long>(S, 65535);
    p.S = (int)std::max<long long>(1, S);
    return p;
}

template <int R, bool U, int O, int TS>
static void launch_one(const InduceArgs &a, const InducePlan &p, cudaStream_t st) {
    dim3 grid(p.ntiles, p.S);
    k_induce_direct<R, U, O, TS><<<grid, kThreads, 0, st>>>(a);
}

template <int R, int TS>
static void launch_ru(const InduceArgs &a, const InducePlan &p, bool uniform, int order, cudaStream_t st) {
    if (uniform) { if (order == 2) launch_one<R, true, 2, TS>(a, p, st); else launch_one<R, true, 3, TS>(a, p, st); }
    else         { if (order == 2) launch_one<R, false, 2, TS>(a, p, st); else launch_one<R, false, 3, TS>(a, p, st); }
}

int launch_induce_direct(const InduceArgs &a, const InducePlan &p, bool uniform, int order, cudaStream_t st) {
    if (p.ts == 64) {
        if (p.R != 1) { set_error("launch_induce_direct: 64-source tiles need R=1"); return 1; }
        launch_ru<1, 64>(a, p, uniform, order, st);
    } else {
        switch (p.R) {
            case 1: launch_ru<1, kTS>(a, p, uniform, order, st); break;
            case 2: launch_ru<2, kTS>(a, p, uniform, order, st); break;
            case 4: launch_ru<4, kTS>(a, p, uniform, order, st); break;
            default: set_error("launch_induce_direct: unsupported R=%d", p.R); return 1;
        }
    }
    UF_CUDA(cudaGetLastError());
    return 0;
}

int sweep_supported(int ndiv) { return ndiv >= 1 && ndiv <= 256; }

template <int RT>
static void launch_sweep_rt(const InduceArgs &a, int ndiv, int S, bool uniform, cudaStream_t st) {
    if (uniform) k_sweep<RT, true><<<S, kThreads, 0, st>>>(a, ndiv);
    else k_sweep<RT, false><<<S, kThreads, 0, st>>>(a, ndiv);
}

int sweep_splits(long long nsrc_ub, int nsreg, int max_S) {
    // chunks of 32 sources; a region adds at most two partial ones.  One CTA per SM for wakes below ~5*10^4 sources,
    // two per SM from there on (profiles/r02_sweep_split_scan.md: 18.5 -> 15.0 us at 10^5, 114 -> 77 us at 10^6 -- eight
    // warps per SM leave the FP64 pipe 46 % busy, sixteen 73 %; below 5*10^4 the second CTA only adds planes);
    // env UNSFLOW_SWEEP_M forces the CTAs per SM (experiments)
    static int forced = [] { const char *e = getenv("UNSFLOW_SWEEP_M"); return e ? atoi(e) : 0; }();
    const long long chunks = std::max<long long>(1, (nsrc_ub + kSweepGroups - 1) / kSweepGroups + 2LL * nsreg);
    const long long smc = sm_count();
    const long long m = forced > 0 ? forced : (chunks >= smc * 10 ? 2 : 1);
    return (int)std::max<long long>(1, std::min<long long>(std::min<long long>(chunks, m * smc), max_S));
}

int launch_sweep(const InduceArgs &a, int ndiv, int S, bool uniform, cudaStream_t st) {
    if (!sweep_supported(ndiv)) { set_error("launch_sweep: ndiv=%d not supported", ndiv); return 1; }
    if (ndiv <= 9 * kSweepTG) launch_sweep_rt<9>(a, ndiv, S, uniform, st);
    else if (ndiv <= 16 * kSweepTG) launch_sweep_rt<16>(a, ndiv, S, uniform, st);
    else launch_sweep_rt<32>(a, ndiv, S, uniform, st);
    UF_CUDA(cudaGetLastError());
    return 0;
}

int launch_induce_finish(const FinishArgs &a, long long nt_ub, cudaStream_t st) {
    if (nt_ub <= 0) return 0;
    const int blocks = (int)((nt_ub + 255) / 256);
    k_induce_finish<<<blocks, 256, 0, st>>>(a);
    UF_CUDA(cudaGetLastError());
    return 0;
}

int launch_sym_finish(const SymFinishArgs &a, long long nt_ub, cudaStream_t st) {
    if (nt_ub <= 0) return 0;
    if (a.ctas > 512) { set_error("launch_sym_finish: %d planes exceed the list capacity", a.ctas); return 1; }
    const long long tb = (long long)kThreads * a.R;
    const long long blocks = ((nt_ub + tb - 1) / tb + a.nreg) * a.R;      // every region may add one partial block
    switch (a.R) {
        case 8: k_sym_finish<8><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
        case 4: k_sym_finish<4><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
        case 2: k_sym_finish<2><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
        case 1: k_sym_finish<1><<<(unsigned)blocks, kThreads, 0, st>>>(a); break;
        default: set_error("launch_sym_finish: unsupported R=%d", a.R); return 1;
    }
    UF_CUDA(cudaGetLastError());
    return 0;
}

size_t sym_smem_bytes(int R) { return sizeof(double) * (size_t)(2 * 3 + 2) * kThreads * R + 2 * sizeof(uint64_t) + 16; }

int sym_rows_per_thread(long long n, int nranks) {
    // env UNSFLOW_SYM_R in {1,2,4,8} forces a tile size (experiments).  Default: 2048-wide tiles (R = 8, one
    // CTA per SM).  With the work split in 1/32-tile units the widest tile wins at every N
    // (profiles/r01_sym_tile_width_scan.md); kept a pure function of (n, nranks) so that all ranks pick the
    // same tiling and unsflow_partition works without a device.
    static int forced = [] { const char *e = getenv("UNSFLOW_SYM_R"); int v = e ? atoi(e) : 0; return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 0; }();
    (void)n; (void)nranks;
    return forced ? forced : 8;
}

long long sym_threshold() {
    // free vortices from which the pair-tile kernel beats the direct one (same scan: 0.21 vs 0.26 ms at
    // N = 16000, 0.21 vs 0.16 ms at 12000); env UNSFLOW_SYM_THRESHOLD overrides (tests)
    static long long v = [] { const char *e = getenv("UNSFLOW_SYM_THRESHOLD"); return e ? atoll(e) : 16384LL; }();
    return v;
}

template <int R, int O>
static int launch_sym_one(const SymArgs &a, int ctas, cudaStream_t st) {
    static bool attr_set[64] = {false};      // per device (a process may drive several through unsflow_set_device)
    const size_t smem = sym_smem_bytes(R);
    int dev = 0;
    UF_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        UF_CUDA(cudaFuncSetAttribute(k_induce_sym<R, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    k_induce_sym<R, O><<<ctas, kThreads, smem, st>>>(a);
    UF_CUDA(cudaGetLastError());
    return 0;
}

template <int R>
static int sym_occ() {
    int nb = 0;
    cudaFuncSetAttribute(k_induce_sym<R, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sym_smem_bytes(R));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_induce_sym<R, 3>, kThreads, sym_smem_bytes(R)) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

int sym_ctas(int R) {
    static int occ[9] = {0};
    if (!occ[R]) occ[R] = R == 8 ? sym_occ<8>() : (R == 4 ? sym_occ<4>() : (R == 2 ? sym_occ<2>() : sym_occ<1>()));
    return occ[R] * sm_count();
}

int sym_ctas_max() { return std::max(std::max(sym_ctas(8), sym_ctas(4)), std::max(sym_ctas(2), sym_ctas(1))); }

int launch_induce_sym(const SymArgs &a_in, int R, int order, cudaStream_t st) {
    const int ctas = sym_ctas(R);
    // timing probe only (results are then partial sums): UNSFLOW_SYM_SHARE="rank,nranks" makes a single GPU run the
    // pair-tile share of one rank of a multi-GPU split (profiles/r02_stepper_rank_shares.md)
    static int share[2] = {0, 0};
    static bool parsed = [] { const char *e = getenv("UNSFLOW_SYM_SHARE"); if (e) sscanf(e, "%d,%d", &share[0], &share[1]); return true; }();
    (void)parsed;
    SymArgs a = a_in;
    if (share[1] > 1 && share[0] >= 0 && share[0] < share[1]) { a.rank = share[0]; a.nranks = share[1]; }
    if (R == 8) return order == 2 ? launch_sym_one<8, 2>(a, ctas, st) : launch_sym_one<8, 3>(a, ctas, st);
    if (R == 4) return order == 2 ? launch_sym_one<4, 2>(a, ctas, st) : launch_sym_one<4, 3>(a, ctas, st);
    if (R == 2) return order == 2 ? launch_sym_one<2, 2>(a, ctas, st) : launch_sym_one<2, 3>(a, ctas, st);
    return order == 2 ? launch_sym_one<1, 2>(a, ctas, st) : launch_sym_one<1, 3>(a, ctas, st);
}

}  // namespace unsflow
