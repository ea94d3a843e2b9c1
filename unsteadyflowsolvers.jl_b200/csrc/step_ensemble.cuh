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
// kEnsUnroll independent source chains per thread
// (sm = the member's shared-memory slabs [x | z | g | vx | vz], each `tot` long: passed as a raw
// shared pointer so the loops compile to LDS instead of generic loads)
template <int kEnsUnroll>
__device__ __forceinline__ void ens_sweep(const LdvmDev &d, double *p2u, double *p2w, int *S2_out, const double *sm, long long tot) {
    const double *wx = sm, *wz = sm + tot, *wg = sm + 2 * tot;
    const StepState *s = d.st;
    const int nd = d.ndiv, G = max(1, (int)blockDim.x / nd);
    const int grp = threadIdx.x / nd, i = threadIdx.x - grp * nd;
    const bool act = grp < G;
    const double vc4 = d.vc4_new;
    if (act) {
        const double tx = d.bnd_x[i], tz = d.bnd_z[i];
        double u[kEnsUnroll], w[kEnsUnroll];
#pragma unroll
        for (int q = 0; q < kEnsUnroll; q++) { u[q] = 0.0; w[q] = 0.0; }
        for (int r = 0; r < 3; r++) {
            const int e = (int)s->wake.end[r];
            int j = (int)s->wake.begin[r] + grp;
            for (; j + (kEnsUnroll - 1) * G < e; j += kEnsUnroll * G) {
#pragma unroll
                for (int q = 0; q < kEnsUnroll; q++) pair_acc(wx[j + q * G], wz[j + q * G], wg[j + q * G], vc4, tx, tz, u[q], w[q]);
            }
#pragma unroll
            for (int q = 0; q < kEnsUnroll - 1; q++)      // remainder, static accumulator indices
                if (j + q * G < e) pair_acc(wx[j + q * G], wz[j + q * G], wg[j + q * G], vc4, tx, tz, u[q], w[q]);
        }
        double su = 0.0, sw = 0.0;
#pragma unroll
        for (int q = 0; q < kEnsUnroll; q++) { su += u[q]; sw += w[q]; }
        p2u[grp * d.p2stride + i] = su;
        p2w[grp * d.p2stride + i] = sw;
    }
    *S2_out = G;
}

// Free vortices of the member as ONE index space: virtual index v in [0, nt + nl) -> slab index (TEV slab first,
// then the LEV slab; members have no external vortices), so a short LEV family does not get a pass of its own.
struct EnsIndex {
    long long tb, nt, lb, nl;
    __device__ __forceinline__ long long phys(long long v) const { return v < nt ? tb + v : lb + (v - nt); }
};

// One pass of the member roll-up: ONE target per thread (virtual index v0 + t) against all sources (free + bound
// vortices) with UN INDEPENDENT source chains per thread (UN accumulator pairs, added in a fixed order at the end).
// The pair chain is 13 dependent FP64 instructions of ~8 cycles each: a member CTA only keeps the FP64 pipe busy if
// every thread has several pairs in flight, and with the register budget of a 3-CTA-per-SM kernel the compiler only
// interleaves chains that are this cheap to keep live (one target, UN sources).
template <int UN>
__device__ __forceinline__ void ens_rollup_pass(const LdvmDev &d, const StepState *s, const EnsIndex ix, long long v0, long long v1, double *sm, long long tot) {
    const double *wx = sm, *wz = sm + tot, *wg = sm + 2 * tot;
    double *wvx = sm + 3 * tot, *wvz = sm + 4 * tot;
    const long long v = v0 + threadIdx.x;
    if (v0 + (threadIdx.x & ~31) >= v1) return;      // warps without a single live target in this pass
    const bool ok = v < v1;
    const long long pi = ok ? ix.phys(v) : 0;
    const double vc4 = d.vc4_new;
    const double tx = ok ? wx[pi] : 0.0, tz = ok ? wz[pi] : 0.0;
    double u[UN], w[UN];
#pragma unroll
    for (int q = 0; q < UN; q++) { u[q] = 0.0; w[q] = 0.0; }
    for (int sr = 0; sr < 4; sr++) {
        const int b = (int)s->wake.begin[sr], e = (int)s->wake.end[sr];
        int j = b;
        for (; j + UN <= e; j += UN) {
#pragma unroll
            for (int q = 0; q < UN; q++) pair_acc(wx[j + q], wz[j + q], wg[j + q], vc4, tx, tz, u[q], w[q]);      // broadcast shared-memory reads
        }
#pragma unroll
        for (int q = 0; q < UN - 1; q++)      // remainder with STATIC accumulator indices (a dynamic one would push u[], w[] to local memory)
            if (j + q < e) pair_acc(wx[j + q], wz[j + q], wg[j + q], vc4, tx, tz, u[q], w[q]);
    }
    double su = 0.0, sw = 0.0;
#pragma unroll
    for (int q = 0; q < UN; q++) { su += u[q]; sw += w[q]; }
    if (ok) {
        wvx[pi] = su * kInvTwoPi + s->fs[0];      // calcs.jl:252-277
        wvz[pi] = sw * kInvTwoPi + s->fs[1];
    }
}

// wakeroll (calcs.jl:222-294) inside one CTA
template <int kEnsUnroll>
__device__ __forceinline__ void ens_rollup(const LdvmDev &d, double *sm, long long tot) {
    StepState *s = d.st;
    EnsIndex ix;
    ix.tb = s->wake.begin[0]; ix.nt = s->wake.end[0] - ix.tb;
    ix.lb = s->wake.begin[1]; ix.nl = s->wake.end[1] - ix.lb;
    const long long N = ix.nt + ix.nl;
    // passes of blockDim targets: the last, partly filled one runs with one warp per scheduler and finishes sooner
    // (equal-size passes were measured slower: the pass time follows the scheduler with the most warps)
    for (long long v0 = 0; v0 < N; v0 += blockDim.x) ens_rollup_pass<kEnsUnroll>(d, s, ix, v0, min(N, v0 + (long long)blockDim.x), sm, tot);
    __syncthreads();
    for (int tr = 0; tr < 2; tr++)                                // calcs.jl:280-291
        for (long long i = s->wake.begin[tr] + threadIdx.x; i < s->wake.end[tr]; i += blockDim.x) {
            sm[i] += d.dt * sm[3 * tot + i];
            sm[tot + i] += d.dt * sm[4 * tot + i];
        }
    __syncthreads();
}

}  // namespace unsflow
