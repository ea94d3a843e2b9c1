// step_ensemble.cuh -- ensemble member loop: block-level sweep and roll-up (k_ensemble body)
#pragma once
#include "step_single.cuh"

namespace unsflow {

// ---------------------------------------------------------------------------------------------
// Ensemble kernel (config 3): a PERSISTENT grid of member CTAs.  Each CTA pulls the next member from an atomic
// queue (members are listed most-expensive-first by the host, so the long ones start early and the short ones
// fill the gaps: longest-processing-time-first list scheduling) and runs the member's WHOLE step loop with
// every piece of mutable member state in SHARED MEMORY: the wake slabs (x, z, strength, vx, vz), the StepState,
// the mutable TwoDSurf arrays (bnd_x/z, uind, wind, downwash, aterm, adot, aprev), the bound-sweep partial
// planes and the solve workspace.  The unchanged single-surface step functions (step_head / step_solve) run on
// an LdvmDev whose pointers aim at that shared memory, so their single-thread sections walk shared-memory
// latencies instead of L2 round trips.  Global memory per member: the kinematics table (read), the mat rows
// (written) and the StepState (initial 2DOF parameters in, final counts out) -- no per-member slabs.
// Members are independent: no inter-CTA communication, any number of GPUs by sharding members.
// All ensemble vortices share the core radius 0.02c (calcs.jl:385,423), hence the scalar vc4.
// ---------------------------------------------------------------------------------------------
struct EnsArgs {
    LdvmDev proto;            // scalars, geometry (theta, x, cam, cam_slope) and table pointers shared by all members
    const double *img;        // initial mutable state shared by all members: [bnd_x, bnd_z, uind, wind, downwash](ndiv each),
                              // [aterm, adot, aprev](naterm each), [bv_x, bv_z, bv_s](ndiv - 1 each)
    StepState *st;            // [M] in: initial state (2DOF parameters per member); out: final counts / 2DOF state
    const double *kin;        // [M][nsteps][9]
    double *mat;              // [M][nsteps][8]
    const double *lespcrit;   // [M] or null
    const int *order;         // [M] member ids, most expensive first
    int *next;                // queue head (zeroed before the launch)
    long long M, nsteps;
    long long tot, off_l, off_e, off_b;   // member slab layout [TEV | LEV | EXTV(empty) | BV]
    int groups;               // source groups of the in-CTA bound sweep (= partial planes)
    long long *prof;          // optional [gridDim.x][8] cycle counters (head, sweep, solve, roll-up, setup), or null
};

// dynamic shared memory of one member CTA, in doubles
__host__ __device__ inline long long ens_smem_doubles(int nd, int na, long long tot, int groups) {
    const long long ndp = (nd + 7) / 8 * 8, nap = (na + 7) / 8 * 8;
    return 5 * tot + 5 * ndp + 3 * nap + 2 * (long long)groups * ndp + solve_ws_doubles(nd, na);
}


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

// update_indbound (calcs.jl:35-38) inside one CTA: group g of threads owns sources j = g (mod G);
// 2 independent source chains per thread
// (sm = the member's shared-memory slabs [x | z | g | vx | vz], each `tot` long: passed as a raw
// shared pointer so the loops compile to LDS instead of generic loads)
__device__ __forceinline__ void ens_sweep(const LdvmDev &d, double *p2u, double *p2w, int *S2_out, const double *sm, long long tot) {
    const double *wx = sm, *wz = sm + tot, *wg = sm + 2 * tot;
    const StepState *s = d.st;
    const int nd = d.ndiv, G = max(1, (int)blockDim.x / nd);
    const int grp = threadIdx.x / nd, i = threadIdx.x - grp * nd;
    const bool act = grp < G;
    const double vc4 = d.vc4_new;
    if (act) {
        const double tx = d.bnd_x[i], tz = d.bnd_z[i];
        double u0 = 0.0, w0 = 0.0, u1 = 0.0, w1 = 0.0;
        for (int r = 0; r < 3; r++) {
            const long long e = s->wake.end[r];
            long long j = s->wake.begin[r] + grp;
            for (; j + G < e; j += 2 * G) {
                pair_acc(wx[j], wz[j], wg[j], vc4, tx, tz, u0, w0);
                pair_acc(wx[j + G], wz[j + G], wg[j + G], vc4, tx, tz, u1, w1);
            }
            if (j < e) pair_acc(wx[j], wz[j], wg[j], vc4, tx, tz, u0, w0);
        }
        p2u[grp * d.p2stride + i] = u0 + u1;
        p2w[grp * d.p2stride + i] = w0 + w1;
    }
    *S2_out = G;
}

// one chunk of RR*blockDim targets of region tr against all sources (free + bound vortices);
// velocities go to vx, vz -- positions are convected only after every chunk is done
template <int RR>
__device__ __forceinline__ void ens_rollup_chunk(const LdvmDev &d, const StepState *s, int tr, long long tb, double *sm, long long tot) {
    const double *wx = sm, *wz = sm + tot, *wg = sm + 2 * tot;
    double *wvx = sm + 3 * tot, *wvz = sm + 4 * tot;
    const long long tend = s->wake.end[tr];
    if (tb + (threadIdx.x & ~31) >= tend) return;      // warps without a single live target
    const double vc4 = d.vc4_new;
    double tx[RR], tz[RR], u[RR], w[RR];
#pragma unroll
    for (int r = 0; r < RR; r++) {
        const long long i = tb + threadIdx.x + (long long)r * blockDim.x;
        const bool ok = i < tend;
        tx[r] = ok ? wx[i] : 0.0; tz[r] = ok ? wz[i] : 0.0; u[r] = 0.0; w[r] = 0.0;
    }
    for (int sr = 0; sr < 4; sr++) {
        const long long b = s->wake.begin[sr], e = s->wake.end[sr];
#pragma unroll 2
        for (long long j = b; j < e; j++) {
            const double xj = wx[j], zj = wz[j], gj = wg[j];      // broadcast shared-memory reads
#pragma unroll
            for (int r = 0; r < RR; r++) pair_acc(xj, zj, gj, vc4, tx[r], tz[r], u[r], w[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < RR; r++) {
        const long long i = tb + threadIdx.x + (long long)r * blockDim.x;
        if (i < tend) {
            wvx[i] = u[r] * kInvTwoPi + s->fs[0];      // calcs.jl:252-277
            wvz[i] = w[r] * kInvTwoPi + s->fs[1];
        }
    }
}

// wakeroll (calcs.jl:222-294) inside one CTA
__device__ __forceinline__ void ens_rollup(const LdvmDev &d, double *sm, long long tot) {
    StepState *s = d.st;
    for (int tr = 0; tr < 3; tr++) {
        for (long long tb = s->wake.begin[tr]; tb < s->wake.end[tr]; tb += (long long)kEnsR * blockDim.x) {
            const long long rem = s->wake.end[tr] - tb;
            const int nrow = (int)min((long long)kEnsR, (rem + blockDim.x - 1) / blockDim.x);   // block-uniform
            switch (nrow) {
                case 1: ens_rollup_chunk<1>(d, s, tr, tb, sm, tot); break;
                case 2: ens_rollup_chunk<2>(d, s, tr, tb, sm, tot); break;
                case 3: ens_rollup_chunk<3>(d, s, tr, tb, sm, tot); break;
                default: ens_rollup_chunk<4>(d, s, tr, tb, sm, tot); break;
            }
        }
    }
    __syncthreads();
    for (int tr = 0; tr < 3; tr++)                                // calcs.jl:280-291
        for (long long i = s->wake.begin[tr] + threadIdx.x; i < s->wake.end[tr]; i += blockDim.x) {
            sm[i] += d.dt * sm[3 * tot + i];
            sm[tot + i] += d.dt * sm[4 * tot + i];
        }
    __syncthreads();
}

}  // namespace unsflow
