// sweep.cuh -- bound-point sweep kernel: few targets (ndiv ~ 70) x many sources (the wake).
// Replaces update_indbound (src/lowOrder2D/calcs.jl:35-38) for ndiv <= 256.
//
// A warp is cut into 32/TG source groups of TG = 8 lanes; lane l of a group owns the RT targets
// l, l+8, l+16, ... (RT = ceil(ndiv/8): 9 for ndiv = 70, 72 slots, 97 % useful) in registers, and
// every group walks its own sources of the TMA-staged tile (the 8 lanes of a group read the same
// shared-memory word: broadcast; the 4 groups of a warp read 4 consecutive words).  RT
// independent pair chains per lane keep the FP64 pipe busy; all 32 lanes do useful work, which
// the generic kernel (one target per thread, 70 of 256 threads live) cannot.
//
// Work split: the live source ranges are cut into CHUNKS of 32 sources (one per source group) and the
// chunk list is cut evenly over the grid, whose size is a multiple of the SM count (or the chunk count,
// for a small wake): every CTA gets the same work to within one chunk, whatever the wake size -- with
// whole 512-source tiles a 10^5 wake was 199 CTAs of one tile each on 148 SMs (51 SMs did double work).
// A CTA stages its chunks in tiles of up to 512 sources (variable-size TMA bulk copies, double-buffered).
// Per-block partial sums are reduced across the 32 groups in shared memory in a fixed order and
// written to one plane per block; the consumer adds the planes in a fixed order (deterministic).
#pragma once
#include "induce.cuh"

namespace unsflow {

constexpr int kSweepTG = 8;                       // lanes per source group
constexpr int kSweepGroups = kThreads / kSweepTG; // 32 source groups per block = sources per chunk

int sweep_supported(int ndiv);                    // 1 when a k_sweep instantiation covers ndiv
// number of source splits = CTAs = partial planes for a wake of at most nsrc_ub sources in nsreg regions
int sweep_splits(long long nsrc_ub, int nsreg, int max_S);
// grid.x = S = number of partial planes written (a.pu/pw, stride a.tstride)
int launch_sweep(const InduceArgs &a, int ndiv, int S, bool uniform, cudaStream_t st);

#if defined(__CUDACC__) && defined(UNSFLOW_INDUCE_KERNELS)

template <int RT, bool UNIFORM>
__global__ void __launch_bounds__(kThreads) k_sweep(const InduceArgs a, int ndiv) {
    constexpr int NARR = UNIFORM ? 3 : 4;
    constexpr int TS = kTS, CH = kSweepGroups, TCH = TS / CH;            // tile capacity in sources / chunks
    constexpr int kRed = (kThreads / 32) * 2 * RT * kSweepTG;          // reduction scratch (doubles)
    constexpr int kBuf = 2 * 4 * TS > kRed ? 2 * 4 * TS : kRed;         // tile buffers, reused by the reduction
    __shared__ __align__(128) double s_raw[kBuf];
    double (*s_buf)[4][TS] = reinterpret_cast<double (*)[4][TS]>(s_raw);   // [2][4][TS]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ volatile int s_n[2];                                     // sources in the tile of each stage

    const int tid = threadIdx.x, lane8 = tid & (kSweepTG - 1), grp = tid / kSweepTG;
    const Regions sreg = *a.sreg;
    // chunk ranges of the regions (absolute CH-aligned chunks covering [begin, end)), this CTA's share [c0, c1)
    long long cb[4], cn[4], C = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const bool live = r < a.nsreg && sreg.end[r] > sreg.begin[r];
        cb[r] = live ? sreg.begin[r] / CH : 0;
        cn[r] = live ? (sreg.end[r] + CH - 1) / CH - cb[r] : 0;
        C += cn[r];
    }
    const long long c0 = C * blockIdx.x / gridDim.x, c1 = C * (blockIdx.x + 1) / gridDim.x;
    int ntile = 0;
    {
        long long base = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const long long lo = c0 > base ? c0 : base, hi = c1 < base + cn[r] ? c1 : base + cn[r];
            if (hi > lo) ntile += (int)((hi - lo + TCH - 1) / TCH);
            base += cn[r];
        }
    }

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
    long long cur = c0;                                                 // next chunk to stage (thread 0)
    auto issue = [&](int stage) {
        // the tile starting at chunk `cur`: up to TCH chunks, never across a region boundary
        long long base = 0, e0 = 0;
        int nch = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (nch == 0 && cur < base + cn[r]) {
                const long long seg_end = c1 < base + cn[r] ? c1 : base + cn[r];
                nch = (int)(seg_end - cur < TCH ? seg_end - cur : TCH);
                e0 = (cb[r] + (cur - base)) * CH;
            }
            base += cn[r];
        }
        cur += nch;
        const uint32_t bytes = (uint32_t)(nch * CH * sizeof(double));
        s_n[stage] = nch * CH;
        mbar_expect_tx(&s_bar[stage], NARR * bytes);
        tma_load_1d(&s_buf[stage][0][0], a.sx + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][1][0], a.sz + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][2][0], a.sg + e0, bytes, &s_bar[stage]);
        if (!UNIFORM) tma_load_1d(&s_buf[stage][NARR - 1][0], a.svc4 + e0, bytes, &s_bar[stage]);
    };
    if (tid == 0 && ntile > 0) issue(0);
    const double vc4u = a.vc4_uniform;
    for (int k = 0; k < ntile; k++) {
        const int stage = k & 1;
        if (tid == 0 && k + 1 < ntile) issue(stage ^ 1);
        mbar_wait(&s_bar[stage], (k >> 1) & 1);
        const int n = s_n[stage];
        const double *sx = s_buf[stage][0], *sz = s_buf[stage][1], *sg = s_buf[stage][2], *sv = s_buf[stage][NARR - 1];
#pragma unroll 1
        for (int j = grp; j < n; j += kSweepGroups) {
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
