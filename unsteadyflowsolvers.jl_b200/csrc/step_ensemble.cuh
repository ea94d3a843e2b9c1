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
    int sym_min;              // free vortices from which the pair-symmetric roll-up is used
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

// ---------------------------------------------------------------------------------------------
// Pair-symmetric member roll-up.  All member vortices share one core radius, so a free-vortex pair needs ONE
// 1/sqrt(d^4 + vc^4) for both directions (16 FP64 instructions per pair = 8 per interaction instead of 13).
// The N free vortices (virtual index space, EnsIndex) are cut into Bq super-blocks of SB (>= 32, any multiple of
// 8) consecutive vortices.  A warp owns one super-block at a time: lane l keeps the rows l, l+32, .. of it
// (position, strength, velocity sums) in registers -- R <= 4 rows per lane.  Work is ordered as a block circulant:
//   phase 0      the warp's own super-block and the bound vortices, one-sided (every lane reads the same source:
//                shared-memory broadcast; no column sums);
//   phase m >= 1 super-block A against C = (A + m) mod Bq, m = 1 .. (Bq-1)/2 (+ half a phase at m = Bq/2 for even
//                Bq): at step k lane l takes column (k + l) mod SB of C -- a Latin square, so the lanes of a warp
//                never touch the same column in one step -- and adds the reaction on that column to the member's
//                shared-memory velocity arrays by a plain read-modify-write (no atomics, no shuffles).
// In one phase the warps' column blocks C are pairwise different (different A, same m), so there is no race between
// warps as long as the phases are separated by a block barrier; the row sums go to shared memory after the last
// phase, between two barriers.  More super-blocks than warps (long runs) are taken in rounds.  Every sum has a fixed
// order: results are bit-reproducible run to run.
// ---------------------------------------------------------------------------------------------
struct EnsSymPlan { int SB, R, nb0, Bq; };      // super-block width, rows per lane, TEV blocks, all blocks
constexpr int kEnsSymMin = 96;      // below this many free vortices the one-sided passes are as fast

// shared-memory accesses by 32-bit address (no generic-to-shared conversion inside the loops)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }

// Super-blocks live in PHYSICAL slab space: the TEV slab is cut into nb0 blocks of SB, the LEV slab into Bq - nb0, so
// a block is always one contiguous index range (no per-element index mapping in the loops).
// The plan minimises the estimated issue time: a warp step costs ~30 issue slots of bookkeeping (column fetch,
// read-modify-write of the column sums, indices) + 33 per row pair.  RMAX = rows a lane may keep in registers.
template <int RMAX>
__device__ __forceinline__ EnsSymPlan ens_sym_plan(int nt, int nl, int nwarps) {
    constexpr int SBMAX = 32 * RMAX;
    EnsSymPlan best;
    best.SB = SBMAX; best.R = RMAX; best.nb0 = (nt + SBMAX - 1) / SBMAX; best.Bq = best.nb0 + (nl + SBMAX - 1) / SBMAX;
    if (nt + nl >= SBMAX * nwarps) return best;
    long long bc = -1;
    for (int sb = 32; sb <= SBMAX; sb += 8) {
        const int r = (sb + 31) / 32, n0 = (nt + sb - 1) / sb, be = n0 + (nl + sb - 1) / sb;
        const int rounds = (be + nwarps - 1) / nwarps;
        // per round every warp walks phase 0 (sb + bound vortices, one-sided) and be/2 symmetric phases of sb steps
        const long long cost = (long long)rounds * ((long long)(sb + 70) * (10 + 26 * r) + (long long)(be / 2) * sb * (30 + 33 * r));
        if (bc < 0 || cost < bc) { bc = cost; best.SB = sb; best.R = r; best.nb0 = n0; best.Bq = be; }
    }
    return best;
}

template <int R>
__device__ __forceinline__ void ens_sym_round(const StepState *s, const EnsSymPlan pl, int A, double vc4, double *sm, long long tot) {
    const uint32_t sx = smem_u32(sm), sz = sx + 8u * (uint32_t)tot, sg = sz + 8u * (uint32_t)tot, su = sg + 8u * (uint32_t)tot, sw = su + 8u * (uint32_t)tot;
    const int lane = threadIdx.x & 31, SB = pl.SB, Bq = pl.Bq;
    const int tb = (int)s->wake.begin[0], te = (int)s->wake.end[0], lb = (int)s->wake.begin[1], le = (int)s->wake.end[1];
    auto blk_base = [&](int C) { return C < pl.nb0 ? tb + C * SB : lb + (C - pl.nb0) * SB; };
    auto blk_cnt = [&](int C) { return C < pl.nb0 ? min(SB, te - (tb + C * SB)) : min(SB, le - (lb + (C - pl.nb0) * SB)); };
    const bool active = A < Bq;
    const int abase = active ? blk_base(A) : 0, acnt = active ? blk_cnt(A) : 0;
    double tx[R], tz[R], tg[R], u[R], w[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int o = lane + 32 * r;
        const bool ok = o < acnt;
        const uint32_t off = 8u * (uint32_t)(abase + (ok ? o : 0));
        tx[r] = ok ? lds_f64(sx + off) : 0.0;
        tz[r] = ok ? lds_f64(sz + off) : 0.0;
        tg[r] = ok ? lds_f64(sg + off) : 0.0;
        u[r] = 0.0; w[r] = 0.0;
    }
    if (active) {
        // ---- phase 0: own super-block, then the bound vortices; one-sided, broadcast reads, >= 4 chains per lane ----
        constexpr int CU = R >= 4 ? 1 : (R >= 2 ? 2 : 4);      // columns per iteration
        double u2[CU > 1 ? CU - 1 : 1][R], w2[CU > 1 ? CU - 1 : 1][R];
#pragma unroll
        for (int q = 0; q < CU - 1; q++)
#pragma unroll
            for (int r = 0; r < R; r++) { u2[q][r] = 0.0; w2[q][r] = 0.0; }
        for (int part = 0; part < 2; part++) {
            int pj = part == 0 ? abase : (int)s->wake.begin[3];
            const int e = part == 0 ? abase + acnt : (int)s->wake.end[3];
            for (; pj + CU <= e; pj += CU) {
#pragma unroll
                for (int q = 0; q < CU; q++) {
                    const uint32_t off = 8u * (uint32_t)(pj + q);
                    const double xj = lds_f64(sx + off), zj = lds_f64(sz + off), gj = lds_f64(sg + off);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (q == 0) pair_acc(xj, zj, gj, vc4, tx[r], tz[r], u[r], w[r]);
                        else pair_acc(xj, zj, gj, vc4, tx[r], tz[r], u2[q > 0 ? q - 1 : 0][r], w2[q > 0 ? q - 1 : 0][r]);
                    }
                }
            }
            for (; pj < e; pj++) {
                const uint32_t off = 8u * (uint32_t)pj;
                const double xj = lds_f64(sx + off), zj = lds_f64(sz + off), gj = lds_f64(sg + off);
#pragma unroll
                for (int r = 0; r < R; r++) pair_acc(xj, zj, gj, vc4, tx[r], tz[r], u[r], w[r]);
            }
        }
#pragma unroll
        for (int q = 0; q < CU - 1; q++)
#pragma unroll
            for (int r = 0; r < R; r++) { u[r] += u2[q][r]; w[r] += w2[q][r]; }
    }
    // ---- phases m >= 1: symmetric tiles, separated by block barriers ----
    const int mfull = (Bq - 1) / 2, mlast = Bq / 2;      // mlast > mfull <=> Bq even: half a phase
    for (int m = 1; m <= mlast; m++) {
        if (m > 1) __syncthreads();
        if (!active || (m > mfull && A >= Bq / 2)) continue;
        int C = A + m;
        if (C >= Bq) C -= Bq;
        const int cnt = blk_cnt(C);
        const uint32_t cb = 8u * (uint32_t)blk_base(C);
        const uint32_t cx = sx + cb, cz = sz + cb, cg = sg + cb, cuA = su + cb, cwA = sw + cb;
        // column of step 0 (lane < 32 <= SB); the column of step k + 1 is fetched while step k computes.  A column
        // beyond the block's live count is replaced by the last live one with zero strength and never written.
        int j = lane;
        bool okc = j < cnt;
        uint32_t oj = 8u * (uint32_t)(okc ? j : cnt - 1);
        double xj = lds_f64(cx + oj), zj = lds_f64(cz + oj), gj = okc ? lds_f64(cg + oj) : 0.0;
        for (int k = 0; k < SB; k++) {
            int jn = j + 1;
            if (jn >= SB) jn = 0;
            const bool okn = jn < cnt;
            const uint32_t on = 8u * (uint32_t)(okn ? jn : cnt - 1);
            const double xn = lds_f64(cx + on), zn = lds_f64(cz + on), gn = okn ? lds_f64(cg + on) : 0.0;
            const double a_u = lds_f64(cuA + oj), a_w = lds_f64(cwA + oj);   // last written by another lane one step ago (program order)
            double cu = 0.0, cw = 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double dx = xj - tx[r], dz = zj - tz[r];
                const double d2 = fma(dz, dz, dx * dx);
                const double y = rsqrt_full<3>(fma(d2, d2, vc4));
                const double a = gj * y, b = tg[r] * y;
                u[r] = fma(-dz, a, u[r]);
                w[r] = fma(dx, a, w[r]);
                cu = fma(dz, b, cu);
                cw = fma(-dx, b, cw);
            }
            if (okc) { sts_f64(cuA + oj, a_u + cu); sts_f64(cwA + oj, a_w + cw); }
            __syncwarp();
            j = jn; okc = okn; oj = on; xj = xn; zj = zn; gj = gn;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int o = lane + 32 * r;
        if (o < acnt) {
            const uint32_t off = 8u * (uint32_t)(abase + o);
            sts_f64(su + off, lds_f64(su + off) + u[r]);
            sts_f64(sw + off, lds_f64(sw + off) + w[r]);
        }
    }
    __syncthreads();
}

template <int RMAX>
__device__ __forceinline__ void ens_rollup_sym(const LdvmDev &d, const StepState *s, double *sm, long long tot) {
    double *au = sm + 3 * tot, *aw = sm + 4 * tot;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const int tb = (int)s->wake.begin[0], te = (int)s->wake.end[0], lb = (int)s->wake.begin[1], le = (int)s->wake.end[1];
    for (int tr = 0; tr < 2; tr++)
        for (int i = (tr ? lb : tb) + threadIdx.x; i < (tr ? le : te); i += blockDim.x) { au[i] = 0.0; aw[i] = 0.0; }
    const EnsSymPlan pl = ens_sym_plan<RMAX>(te - tb, le - lb, nwarps);
    __syncthreads();
    for (int A0 = 0; A0 < pl.Bq; A0 += nwarps) {
        const int A = A0 + warp;
        if (RMAX >= 4 && pl.R == 4) ens_sym_round<4>(s, pl, A, d.vc4_new, sm, tot);
        else if (RMAX >= 3 && pl.R == 3) ens_sym_round<3>(s, pl, A, d.vc4_new, sm, tot);
        else if (RMAX >= 2 && pl.R == 2) ens_sym_round<2>(s, pl, A, d.vc4_new, sm, tot);
        else ens_sym_round<1>(s, pl, A, d.vc4_new, sm, tot);
    }
    const double f0 = s->fs[0], f1 = s->fs[1];
    for (int tr = 0; tr < 2; tr++)      // calcs.jl:252-277
        for (int i = (tr ? lb : tb) + threadIdx.x; i < (tr ? le : te); i += blockDim.x) {
            au[i] = au[i] * kInvTwoPi + f0;
            aw[i] = aw[i] * kInvTwoPi + f1;
        }
}

// wakeroll (calcs.jl:222-294) inside one CTA
template <int kEnsUnroll, int kEnsSym>      // kEnsSym: 0 = one-sided passes, else rows per lane of the pair-symmetric roll-up
__device__ __forceinline__ void ens_rollup(const LdvmDev &d, double *sm, long long tot, int sym_min) {
    StepState *s = d.st;
    EnsIndex ix;
    ix.tb = s->wake.begin[0]; ix.nt = s->wake.end[0] - ix.tb;
    ix.lb = s->wake.begin[1]; ix.nl = s->wake.end[1] - ix.lb;
    const long long N = ix.nt + ix.nl;
    if (kEnsSym > 0 && N >= sym_min) ens_rollup_sym<(kEnsSym > 0 ? kEnsSym : 1)>(d, s, sm, tot);
    else
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
