// step_ensemble.cuh -- ensemble member loop: block-level sweep and roll-up (k_ensemble body)
#pragma once
#include "step_single.cuh"

namespace unsflow {

// ---------------------------------------------------------------------------------------------
// Ensemble kernel (config 3): one CTA per member runs the WHOLE step loop; the member's state
// (a few hundred vortices, <100 KB) stays in L1/L2, induction tiles are staged in shared memory.
// Members are independent: no inter-CTA communication, any number of GPUs by sharding members.
// ---------------------------------------------------------------------------------------------
constexpr int kEnsTile = 256;   // sources staged per shared-memory tile
constexpr int kEnsR = 4;        // targets per thread in the member roll-up

// pair kernel shared by the two block-level induction loops (13 FP64 instr, cubic rsqrt)
__device__ __forceinline__ void pair_acc(double xj, double zj, double gj, double vc4, double tx, double tz, double &u, double &w) {
    const double dx = xj - tx, dz = zj - tz;
    const double d2 = fma(dz, dz, dx * dx);
    const double y = rsqrt_full<3>(fma(d2, d2, vc4));
    const double m = gj * y;
    u = fma(-dz, m, u);
    w = fma(dx, m, w);
}

// cooperative load of sources [b, b+cnt) into a zero-strength-padded tile
__device__ __forceinline__ void ens_stage(const LdvmDev &d, long long b, int cnt, double *sx, double *sz, double *sg, double *sv) {
    for (int k = threadIdx.x; k < kEnsTile; k += blockDim.x) {
        const bool ok = k < cnt;
        sx[k] = ok ? d.wx[b + k] : 0.0;
        sz[k] = ok ? d.wz[b + k] : 0.0;
        sg[k] = ok ? d.wg[b + k] : 0.0;
        sv[k] = ok ? d.wvc4[b + k] : 1.0;
    }
}

// update_indbound (calcs.jl:35-38) inside one CTA: group g of threads owns sources j = g (mod G)
__device__ void ens_sweep(const LdvmDev &d, double *p2u, double *p2w, int *S2_out, double *tile) {
    const StepState *s = d.st;
    const int nd = d.ndiv, G = max(1, (int)blockDim.x / nd);
    const int grp = threadIdx.x / nd, i = threadIdx.x - grp * nd;
    const bool act = grp < G;
    double *sx = tile, *sz = tile + kEnsTile, *sg = tile + 2 * kEnsTile, *sv = tile + 3 * kEnsTile;
    const double tx = act ? d.bnd_x[i] : 0.0, tz = act ? d.bnd_z[i] : 0.0;
    double u = 0.0, w = 0.0;
    for (int r = 0; r < 3; r++) {
        for (long long b = s->wake.begin[r]; b < s->wake.end[r]; b += kEnsTile) {
            const int cnt = (int)min((long long)kEnsTile, s->wake.end[r] - b);
            __syncthreads();
            ens_stage(d, b, cnt, sx, sz, sg, sv);
            __syncthreads();
            if (act)
                for (int j = grp; j < cnt; j += G) pair_acc(sx[j], sz[j], sg[j], sv[j], tx, tz, u, w);
        }
    }
    if (act) { p2u[grp * d.p2stride + i] = u; p2w[grp * d.p2stride + i] = w; }
    *S2_out = G;
    __syncthreads();
}

// one chunk of RR*blockDim targets of region tr against all sources (free + bound vortices)
template <int RR>
__device__ __forceinline__ void ens_rollup_chunk(const LdvmDev &d, const StepState *s, int tr, long long tb, double *tile) {
    double *sx = tile, *sz = tile + kEnsTile, *sg = tile + 2 * kEnsTile, *sv = tile + 3 * kEnsTile;
    const long long tend = s->wake.end[tr];
    double tx[RR], tz[RR], u[RR], w[RR];
#pragma unroll
    for (int r = 0; r < RR; r++) {
        const long long i = tb + threadIdx.x + (long long)r * blockDim.x;
        const bool ok = i < tend;
        tx[r] = ok ? d.wx[i] : 0.0; tz[r] = ok ? d.wz[i] : 0.0; u[r] = 0.0; w[r] = 0.0;
    }
    // warps without a single live target skip the arithmetic (they still help staging)
    const bool warp_live = tb + (threadIdx.x & ~31) < tend;
    for (int sr = 0; sr < 4; sr++) {
        for (long long b = s->wake.begin[sr]; b < s->wake.end[sr]; b += kEnsTile) {
            const int cnt = (int)min((long long)kEnsTile, s->wake.end[sr] - b);
            __syncthreads();
            ens_stage(d, b, cnt, sx, sz, sg, sv);
            __syncthreads();
            if (warp_live) {
                const int cpad = (cnt + 1) & ~1;
#pragma unroll 2
                for (int j = 0; j < cpad; j++) {
                    const double xj = sx[j], zj = sz[j], gj = sg[j], vj = sv[j];
#pragma unroll
                    for (int r = 0; r < RR; r++) pair_acc(xj, zj, gj, vj, tx[r], tz[r], u[r], w[r]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RR; r++) {
        const long long i = tb + threadIdx.x + (long long)r * blockDim.x;
        if (i < tend) {
            d.wvx[i] = u[r] * kInvTwoPi + s->fs[0];      // calcs.jl:252-277
            d.wvz[i] = w[r] * kInvTwoPi + s->fs[1];
        }
    }
}

// wakeroll (calcs.jl:222-294) inside one CTA
__device__ void ens_rollup(const LdvmDev &d, double *tile) {
    StepState *s = d.st;
    for (int tr = 0; tr < 3; tr++) {
        for (long long tb = s->wake.begin[tr]; tb < s->wake.end[tr]; tb += (long long)kEnsR * blockDim.x) {
            const long long rem = s->wake.end[tr] - tb;
            const int nrow = (int)min((long long)kEnsR, (rem + blockDim.x - 1) / blockDim.x);   // block-uniform
            switch (nrow) {
                case 1: ens_rollup_chunk<1>(d, s, tr, tb, tile); break;
                case 2: ens_rollup_chunk<2>(d, s, tr, tb, tile); break;
                case 3: ens_rollup_chunk<3>(d, s, tr, tb, tile); break;
                default: ens_rollup_chunk<4>(d, s, tr, tb, tile); break;
            }
        }
    }
    __syncthreads();
    for (int tr = 0; tr < 3; tr++)                                // calcs.jl:280-291
        for (long long i = s->wake.begin[tr] + threadIdx.x; i < s->wake.end[tr]; i += blockDim.x) {
            d.wx[i] += d.dt * d.wvx[i];
            d.wz[i] += d.dt * d.wvz[i];
        }
    __syncthreads();
}

}  // namespace unsflow
