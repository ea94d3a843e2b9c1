// induce.cuh -- Vatistas-core (n=2) Biot-Savart induction kernels for sm_100a, FP64.
//
// Replaces the reference's ind_vel (src/lowOrder2D/calcs.jl:462-477), mutual_ind (:491-510)
// and the induction + convection part of wakeroll (:222-294).
//
//   u_i = -(1/2pi) sum_j G_j (z_j - z_i) / sqrt(vc_j^4 + d_ij^4)
//   w_i = +(1/2pi) sum_j G_j (x_j - x_i) / sqrt(vc_j^4 + d_ij^4)
//
// Layout: vortex state is SoA in HBM (x[], z[], g[], vc4[]); a "region" is a live index range
// [begin,end) inside a TS-aligned slab (TEV / LEV / EXTV / bound vortices).  Dead or padding
// entries inside a processed tile carry g = 0, finite x,z and vc4 > 0, so they contribute an
// exact 0 and tiles are processed without bounds checks.
//
// k_induce_direct : one-sided target x source tiles; targets in registers (R per thread),
//                   sources staged in shared memory by 1-D TMA bulk copies (cp.async.bulk ->
//                   UBLKCP) double-buffered on mbarriers, broadcast LDS in the inner loop.
//                   13 FP64-pipe instructions per interaction with the cubic rsqrt.
// The factor 1/(2 pi) is applied once per target in k_induce_finish, not per pair.
#pragma once
#include "common.cuh"

namespace unsflow {

struct Regions {
    long long begin[4];
    long long end[4];
};

struct InduceArgs {
    const double *sx, *sz, *sg, *svc4;  // sources (absolute indexing into slabs)
    const double *tx, *tz;              // targets (absolute indexing)
    const Regions *sreg;                // device memory
    const Regions *treg;                // device memory
    int nsreg, ntreg;
    double vc4_uniform;                 // used when UNIFORM
    double *pu, *pw;                    // partial sums [S][tstride]
    long long tstride;
};

// ---------------------------------------------------------------------------------------------
// k_induce_finish: v_i = (1/2pi) * sum_s partial[s][i]  (+ free stream)  (+= into vx,vz),
// then optional explicit-Euler convection x += dt*vx (calcs.jl:252-291).  Fixed summation
// order over the split planes -> bit-reproducible.
// ---------------------------------------------------------------------------------------------
struct FinishArgs {
    const double *pu, *pw;
    long long tstride;
    int S;
    const Regions *treg;
    int ntreg;
    double *vx, *vz;        // outputs (absolute indexing)
    double *x, *z;          // positions to convect (may be null)
    const double *fs;       // device pointer to (u_inf, w_inf) or null
    double fs_u, fs_w;      // by-value free stream (added when fs == null)
    double dt;              // 0 -> no convection
    int accumulate;         // 1: vx += result (mutual_ind semantics)
    int raw;                // 1: store the plain plane sum only (no 1/2pi, free stream, convection)
    int packed;             // 1: the partials are indexed by the live-target ordinal (regions back to back), not by slab index
};


#if defined(__CUDACC__) && defined(UNSFLOW_INDUCE_KERNELS)

__device__ __forceinline__ long long tiles_in(long long b, long long e, long long ts) {
    return e > b ? (e + ts - 1) / ts - b / ts : 0;
}

// linear source-tile index -> first absolute element of that tile
template <int TS>
__device__ __forceinline__ long long src_tile_start(const Regions &rg, int nreg, long long lin) {
    for (int r = 0; r < nreg; r++) {
        const long long n = tiles_in(rg.begin[r], rg.end[r], TS);
        if (lin < n) return (rg.begin[r] / TS + lin) * TS;
        lin -= n;
    }
    return -1;
}

// TS = sources per TMA tile: 512 for throughput (large N), 64 for latency (small N and the
// 70-target bound sweep: many short CTAs, 4 independent source chains per thread)
template <int R, bool UNIFORM, int ORDER, int TS>
__global__ void __launch_bounds__(kThreads) k_induce_direct(const InduceArgs a) {
    constexpr int NARR = UNIFORM ? 3 : 4;
    constexpr int kTS = TS;
    __shared__ __align__(128) double s_buf[2][NARR][kTS];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x;
    const Regions treg = *a.treg;
    const Regions sreg = *a.sreg;

    // ---- which targets does this block own? ------------------------------------------------
    constexpr long long TT = (long long)kThreads * R;
    long long tb = blockIdx.x, tbase = -1, tend = 0;
    for (int r = 0; r < a.ntreg; r++) {
        const long long n = treg.end[r] > treg.begin[r] ? (treg.end[r] - treg.begin[r] + TT - 1) / TT : 0;
        if (tb < n) { tbase = treg.begin[r] + tb * TT; tend = treg.end[r]; break; }
        tb -= n;
    }
    if (tbase < 0) return;  // grid was sized from a host-side upper bound

    // ---- which source tiles? ----------------------------------------------------------------
    long long T = 0;
    for (int r = 0; r < a.nsreg; r++) T += tiles_in(sreg.begin[r], sreg.end[r], kTS);
    const long long S = gridDim.y;
    const long long t0 = T * blockIdx.y / S, t1 = T * (blockIdx.y + 1) / S;

    double tx[R], tz[R], u[R], w[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long i = tbase + tid + (long long)r * kThreads;
        const bool ok = i < tend;
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
        constexpr uint32_t bytes = kTS * sizeof(double);
        mbar_expect_tx(&s_bar[stage], NARR * bytes);
        tma_load_1d(&s_buf[stage][0][0], a.sx + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][1][0], a.sz + e0, bytes, &s_bar[stage]);
        tma_load_1d(&s_buf[stage][2][0], a.sg + e0, bytes, &s_bar[stage]);
        if (!UNIFORM) tma_load_1d(&s_buf[stage][NARR - 1][0], a.svc4 + e0, bytes, &s_bar[stage]);
    };

    if (tid == 0 && t0 < t1) issue(t0, 0);

    const double vc4u = a.vc4_uniform;
    for (long long it = t0; it < t1; it++) {
        const int k = (int)(it - t0);
        const int stage = k & 1;
        if (tid == 0 && it + 1 < t1) issue(it + 1, stage ^ 1);
        mbar_wait(&s_bar[stage], (k >> 1) & 1);
        const double *sx = s_buf[stage][0], *sz = s_buf[stage][1], *sg = s_buf[stage][2];
        const double *sv = s_buf[stage][NARR - 1];
#pragma unroll(TS >= 512 ? 2 : 4)
        for (int j = 0; j < kTS; j++) {
            const double xj = sx[j], zj = sz[j], gj = sg[j];
            const double vc4 = UNIFORM ? vc4u : sv[j];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double dx = xj - tx[r];
                const double dz = zj - tz[r];
                const double d2 = fma(dz, dz, dx * dx);
                const double q = fma(d2, d2, vc4);
                const double y = rsqrt_full<ORDER>(q);
                const double m = gj * y;
                u[r] = fma(-dz, m, u[r]);
                w[r] = fma(dx, m, w[r]);
            }
        }
        __syncthreads();  // everyone is done with this stage before thread 0 refills it
    }

    double *pu = a.pu + (long long)blockIdx.y * a.tstride;
    double *pw = a.pw + (long long)blockIdx.y * a.tstride;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long i = tbase + tid + (long long)r * kThreads;
        if (i < tend) { pu[i] = u[r]; pw[i] = w[r]; }
    }
}

__global__ void __launch_bounds__(256) k_induce_finish(const FinishArgs a) {
    const Regions treg = *a.treg;
    long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ord = lin;
    long long i = -1;
    for (int r = 0; r < a.ntreg; r++) {
        const long long n = treg.end[r] > treg.begin[r] ? treg.end[r] - treg.begin[r] : 0;
        if (lin < n) { i = treg.begin[r] + lin; break; }
        lin -= n;
    }
    if (i < 0) return;
    double su = 0.0, sw = 0.0;
    for (int s = 0; s < a.S; s += 16) {   // masked batches of 16 independent loads; the additions stay in plane order
        double bu[16], bw[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const bool ok = s + k < a.S;
            bu[k] = ok ? a.pu[(long long)(s + k) * a.tstride + (a.packed ? ord : i)] : 0.0;
            bw[k] = ok ? a.pw[(long long)(s + k) * a.tstride + (a.packed ? ord : i)] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) { su += bu[k]; sw += bw[k]; }
    }
    if (a.raw) { a.vx[i] = su; a.vz[i] = sw; return; }
    double vx = su * kInvTwoPi, vz = sw * kInvTwoPi;
    const double fu = a.fs ? a.fs[0] : a.fs_u, fw = a.fs ? a.fs[1] : a.fs_w;
    vx += fu;
    vz += fw;
    if (a.accumulate) { vx += a.vx[i]; vz += a.vz[i]; }
    a.vx[i] = vx;
    a.vz[i] = vz;
    if (a.dt != 0.0 && a.x) {
        a.x[i] += a.dt * vx;
        a.z[i] += a.dt * vz;
    }
}

#endif  // kernels

// host-side launch helper (induce.cu)
struct InducePlan {
    int R;        // targets per thread
    int S;        // source splits (gridDim.y)
    int ntiles;   // target tiles (gridDim.x)
    int ts;       // sources per tile: 512 (throughput) or 64 (latency)
};
// ntargets_ub / nsrc_ub: host-side upper bounds of live targets / sources; nregions of each
InducePlan plan_induce(long long ntargets_ub, int ntreg, long long nsrc_ub, int nsreg, int sm_count, int max_S);
int launch_induce_direct(const InduceArgs &a, const InducePlan &p, bool uniform, int order, cudaStream_t st);
int launch_induce_finish(const FinishArgs &a, long long ntargets_ub, cudaStream_t st);

}  // namespace unsflow
