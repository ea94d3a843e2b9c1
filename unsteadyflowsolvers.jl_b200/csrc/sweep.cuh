// sweep.cuh -- bound-point sweep kernel: few targets (ndiv ~ 70) x many sources (the wake).
// Replaces update_indbound (src/lowOrder2D/calcs.jl:35-38) for ndiv <= 256.
//
// A warp is cut into 32/TG source groups of TG = 8 lanes; lane l of a group owns the RT targets
// l, l+8, l+16, ... (RT = ceil(ndiv/8): 9 for ndiv = 70, 72 slots, 97 % useful) in registers, and
// every group walks its own sources of the TMA-staged tile (the 8 lanes of a group read the same
// shared-memory word: broadcast; the 4 groups of a warp read 4 consecutive words).  RT
// independent pair chains per lane keep the FP64 pipe busy; all 32 lanes do useful work, which
// the generic kernel (one target per thread, 70 of 256 threads live) cannot.
// Per-block partial sums are reduced across the 32 groups in shared memory in a fixed order and
// written to one plane per block; the consumer adds the planes in order (deterministic).
#pragma once
#include "induce.cuh"

namespace unsflow {

constexpr int kSweepTG = 8;                       // lanes per source group
constexpr int kSweepGroups = kThreads / kSweepTG; // 32 source groups per block

int sweep_supported(int ndiv);                    // 1 when a k_sweep instantiation covers ndiv
// grid.x = number of source splits = number of partial planes written (a.pu/pw, stride a.tstride)
int launch_sweep(const InduceArgs &a, int ndiv, int S, bool uniform, cudaStream_t st, int ts = kTS);
int sweep_tile(int ndiv, long long nsrc_ub);     // 64 for small wakes (ndiv <= 72), else 512

#if defined(__CUDACC__) && defined(UNSFLOW_INDUCE_KERNELS)

// TS = sources per tile: 512 (throughput) or 64 (latency: small wakes spread over more CTAs)
template <int RT, bool UNIFORM, int TS>
__global__ void __launch_bounds__(kThreads) k_sweep(const InduceArgs a, int ndiv) {
    constexpr int NARR = UNIFORM ? 3 : 4;
    constexpr int kRed = (kThreads / 32) * 2 * RT * kSweepTG;          // reduction scratch (doubles)
    constexpr int kBuf = 2 * 4 * TS > kRed ? 2 * 4 * TS : kRed;         // tile buffers, reused by the reduction
    __shared__ __align__(128) double s_raw[kBuf];
    double (*s_buf)[4][TS] = reinterpret_cast<double (*)[4][TS]>(s_raw);   // [2][4][TS]
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane8 = tid & (kSweepTG - 1), grp = tid / kSweepTG;
    const Regions sreg = *a.sreg;
    long long T = 0;
    for (int r = 0; r < a.nsreg; r++) T += tiles_in(sreg.begin[r], sreg.end[r], TS);
    const long long S = gridDim.x;
    const long long t0 = T * blockIdx.x / S, t1 = T * (blockIdx.x + 1) / S;

    double tx[RT], tz[RT], u[RT], w[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) {
        const int i = lane8 + r * kSweepTG;
        const bool ok = i < ndiv;
        tx[r] = ok ? a.tx[i] : 0.0;
        tz[r] = ok ? a.tz[i] : 0.0;
        u[r] = 0.0;
        w[r] = 0.0;
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](long long lin, int stage) {
        const long long e0 = src_tile_start<TS>(sreg, a.nsreg, lin);
        constexpr uint32_t bytes = TS * sizeof(double);
        mbar_expect_tx(&s_bar[stage], NARR * bytes);
        tma_load_1d(&s_buf[stage][0][0], a.sx + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][1][0], a.sz + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][2][0], a.sg + e0, bytes, &s_bar[stage]);
        if (!UNIFORM) tma_load_1d(&s_buf[stage][NARR - 1][0], a.svc4 + e0, bytes, &s_bar[stage]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);
    const double vc4u = a.vc4_uniform;
    for (long long it = t0; it < t1; it++) {
        const int k = (int)(it - t0), stage = k & 1;
        if (tid == 0 && it + 1 < t1) issue(it + 1, stage ^ 1);
        mbar_wait(&s_bar[stage], (k >> 1) & 1);
        const double *sx = s_buf[stage][0], *sz = s_buf[stage][1], *sg = s_buf[stage][2], *sv = s_buf[stage][NARR - 1];
#pragma unroll 1
        for (int j = grp; j < TS; j += kSweepGroups) {
            const double xj = sx[j], zj = sz[j], gj = sg[j];
            const double vc4 = UNIFORM ? vc4u : sv[j];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double dx = xj - tx[r], dz = zj - tz[r];
                const double d2 = fma(dz, dz, dx * dx);
                const double y = rsqrt_full<3>(fma(d2, d2, vc4));
                const double m = gj * y;
                u[r] = fma(-dz, m, u[r]);
                w[r] = fma(dx, m, w[r]);
            }
        }
        __syncthreads();
    }
    // fixed-order reduction: the 4 groups of a warp by butterfly, then the 8 warps through the
    // (now free) tile buffers
    double *s_red = s_raw;                                 // [warp][2 * RT * kSweepTG]
    const int warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < RT; r++) {
        double uu = u[r], ww = w[r];
        uu += __shfl_xor_sync(0xffffffffu, uu, 8);  ww += __shfl_xor_sync(0xffffffffu, ww, 8);
        uu += __shfl_xor_sync(0xffffffffu, uu, 16); ww += __shfl_xor_sync(0xffffffffu, ww, 16);
        if ((tid & 31) < kSweepTG) {
            s_red[warp * (2 * RT * kSweepTG) + 2 * (lane8 + r * kSweepTG)] = uu;
            s_red[warp * (2 * RT * kSweepTG) + 2 * (lane8 + r * kSweepTG) + 1] = ww;
        }
    }
    __syncthreads();
    for (int i = tid; i < ndiv; i += kThreads) {
        double su = 0.0, sw = 0.0;
#pragma unroll
        for (int g = 0; g < kThreads / 32; g++) { su += s_red[g * (2 * RT * kSweepTG) + 2 * i]; sw += s_red[g * (2 * RT * kSweepTG) + 2 * i + 1]; }
        a.pu[(long long)blockIdx.x * a.tstride + i] = su;
        a.pw[(long long)blockIdx.x * a.tstride + i] = sw;
    }
}

#endif
}  // namespace unsflow
