// step_cluster.cuh -- small wakes: the WHOLE time-step loop of ldvm / ldvm2DOF / ldvmLin / lautat in one launch
// of ONE thread-block cluster (k_cluster_steps body).
//
// Below a few thousand vortices a step of the five-kernel stepper (ldvm.cu) is five dependent launches of
// 3-10 us each: latency, not arithmetic (config C1, the reference's own test: 28-31 us per step).  Here a cluster
// of C CTAs (8, or 16 with the non-portable size) stays resident for a chunk of steps and walks the same phases
// with hardware cluster barriers (barrier.cluster, release / acquire at cluster scope: ~0.2 us) in place of
// kernel boundaries:
//
//   helpers    bound sweep   update_indbound (calcs.jl:35-38): helper c takes the c-th slice of the free vortices,
//                            8-lane source groups x 9 bound points per lane as in k_sweep -> one plane per helper
//   CTA 0      step_head_b   place_tev (the second half of step_head), meanwhile
//   barrier
//   CTA 0      step_solve    Kelvin / Kelvin-Kutta, LESP, place_lev, A_n, update_bv, Spalart merge, forces,
//                            mat row (unchanged; the cos / sin tables stay in its shared memory for the whole chunk)
//   helpers    wakeroll, early part: the wake-on-wake interactions only need what is known before the solve
//   barrier
//   helpers    wakeroll, late part: the vortices shed in this step and the bound vortices as sources; barrier inside
//                            (every CTA holds its snapshot, positions may now change); P neighbouring lanes share one
//                            target and split the sources P ways, shuffle sum in a fixed order; the group's first lane
//                            adds 1/2pi and the free stream and convects (explicit Euler) straight into the slabs --
//                            no partial planes, no finish (calcs.jl:222-294)
//   CTA 0      the LEV shed in this step as a target, then step_head_a of the NEXT step (kinematics row / 2DOF,
//              update_boundpos: nothing the roll-up reads), meanwhile
//   barrier
//
// CTA 0 keeps StepState and the surface arrays in its shared memory for the whole chunk and publishes the live ranges,
// the free stream and the bound points before each barrier (k_cluster_steps, ldvm.cu).
// Everything a CTA needs from another one travels through global memory (L2) between two cluster barriers; the
// acquire side of the barrier invalidates the SM's L1.  Deterministic: fixed work split, fixed summation orders.
#pragma once
#include "step_single.cuh"

namespace unsflow {

constexpr int kClRT = 9;                      // bound points per lane of the sweep: ndiv <= 72
constexpr int kClMaxFree = 2048;              // staging capacity in free vortices (the host's upper bound stays below)
constexpr int kClCap = kClMaxFree + kNdMax + 8;   // + bound vortices (+ the two vortices shed in the step)

struct ClusterArgs {
    LdvmDev d;             // p2u / p2w / p2stride / S2 (= helper CTAs = bound-sweep planes) set by the host
    double *planes_u, *planes_w;   // the same planes, writable
    long long nsteps;
    int tabs_in_smem;      // cos / sin tables copied once into CTA 0's shared memory
    int overlap;           // 1: helper CTAs run the wake-on-wake part of the roll-up while CTA 0 solves (when possible)
    double vc4u;           // uniform core radius^4 (UNIFORM instantiation)
    double dt;
    long long *prof;       // optional [5] cycle counters of CTA 0 ([1] sweep, [2] solve, [3] late roll-up phase, [4] steps) or null
};

// CTA 0's shared-memory copy of the surface: [bnd_x, bnd_z, uind, wind, downwash, theta, x, cam, cam_slope](ndp each)
// [aterm, adot, aprev](nap each)
__host__ __device__ inline size_t cluster_surf_doubles(int nd, int na) {
    return (size_t)(9 * ((nd + 7) / 8 * 8) + 3 * ((na + 7) / 8 * 8));
}
__host__ __device__ inline size_t cluster_smem_doubles(int nd, int na, bool tabs, bool uniform) {
    const size_t ws = (size_t)((solve_ws_doubles(nd, na) + 15) / 16 * 16);
    const size_t tab = tabs ? (size_t)((2 * (na + 1) * nd + 15) / 16 * 16) : 0;
    return ws + tab + (cluster_surf_doubles(nd, na) + 15) / 16 * 16 + (size_t)(uniform ? 3 : 4) * kClCap;
}

#ifdef __CUDACC__

__device__ __forceinline__ unsigned cl_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// free vortices as one index space: virtual index v -> slab index (TEV, LEV, EXTV slabs, then the bound vortices)
struct ClIndex {
    long long b[4], n[4];
    __device__ __forceinline__ void load(const StepState *s) {
#pragma unroll
        for (int r = 0; r < 4; r++) { b[r] = s->wake.begin[r]; const long long c = s->wake.end[r] - b[r]; n[r] = c > 0 ? c : 0; }
    }
    __device__ __forceinline__ long long nfree() const { return n[0] + n[1] + n[2]; }
    __device__ __forceinline__ long long phys(long long v) const {
        if (v < n[0]) return b[0] + v;
        v -= n[0];
        if (v < n[1]) return b[1] + v;
        v -= n[1];
        if (v < n[2]) return b[2] + v;
        return b[3] + (v - n[2]);
    }
};

__device__ __forceinline__ void cl_pair(double xj, double zj, double gj, double vc4, double tx, double tz, double &u, double &w) {
    const double dx = xj - tx, dz = zj - tz;
    const double d2 = fma(dz, dz, dx * dx);
    const double y = rsqrt_full<3>(fma(d2, d2, vc4));
    const double m = gj * y;
    u = fma(-dz, m, u);
    w = fma(dx, m, w);
}

// update_indbound: slice `rank` of the free vortices against all bound points -> plane `rank`
// (ix: the live ranges as published after the previous solve -- the TEV the head is placing meanwhile has strength 0)
template <bool UNIFORM>
__device__ __forceinline__ void cl_sweep(const ClusterArgs &a, const ClIndex &ix, unsigned rank, unsigned C, double *s_red) {
    const LdvmDev &d = a.d;
    const int tid = threadIdx.x, lane8 = tid & 7, grp = tid >> 3, warp = tid >> 5, nd = d.ndiv;
    const long long Nf = ix.nfree();
    const long long c0 = Nf * rank / C, c1 = Nf * (rank + 1) / C;
    double tx[kClRT], tz[kClRT], u[kClRT], w[kClRT];
#pragma unroll
    for (int r = 0; r < kClRT; r++) {
        const int i = lane8 + 8 * r;
        const bool ok = i < nd;
        tx[r] = ok ? d.bnd_x[i] : 0.0;
        tz[r] = ok ? d.bnd_z[i] : 0.0;
        u[r] = 0.0;
        w[r] = 0.0;
    }
#pragma unroll 2
    for (long long v = c0 + grp; v < c1; v += 32) {
        const long long p = ix.phys(v);
        const double xj = d.wx[p], zj = d.wz[p], gj = d.wg[p];
        const double vc4 = UNIFORM ? a.vc4u : d.wvc4[p];
#pragma unroll
        for (int r = 0; r < kClRT; r++) cl_pair(xj, zj, gj, vc4, tx[r], tz[r], u[r], w[r]);
    }
    // the 4 groups of a warp by butterfly, then the 8 warps through shared memory, in a fixed order
    constexpr int kRow = 2 * kClRT * 8;
#pragma unroll
    for (int r = 0; r < kClRT; r++) {
        double uu = u[r], ww = w[r];
        uu += __shfl_xor_sync(0xffffffffu, uu, 8);  ww += __shfl_xor_sync(0xffffffffu, ww, 8);
        uu += __shfl_xor_sync(0xffffffffu, uu, 16); ww += __shfl_xor_sync(0xffffffffu, ww, 16);
        if ((tid & 31) < 8) {
            s_red[warp * kRow + 2 * (lane8 + 8 * r)] = uu;
            s_red[warp * kRow + 2 * (lane8 + 8 * r) + 1] = ww;
        }
    }
    __syncthreads();
    for (int i = tid; i < nd; i += blockDim.x) {
        double su = 0.0, sw = 0.0;
#pragma unroll
        for (int g = 0; g < 8; g++) { su += s_red[g * kRow + 2 * i]; sw += s_red[g * kRow + 2 * i + 1]; }
        a.planes_u[(long long)rank * d.p2stride + i] = su;
        a.planes_w[(long long)rank * d.p2stride + i] = sw;
    }
    __syncthreads();      // s_red is the staging area of the roll-up
}

// wakeroll: stage all sources, cluster barrier, P lanes per target, convect
template <bool UNIFORM>
__device__ __forceinline__ void cl_rollup(const ClusterArgs &a, unsigned rank, unsigned C, double *stg) {
    const LdvmDev &d = a.d;
    const StepState *s = d.st;
    const int tid = threadIdx.x;
    ClIndex ix;
    ix.load(s);
    const long long Nf = ix.nfree();
    const int Ns = (int)(Nf + ix.n[3]);
    double *sx = stg, *sz = stg + kClCap, *sg = stg + 2 * kClCap, *sv = stg + 3 * kClCap;
    for (int k = tid; k < Ns; k += blockDim.x) {
        const long long p = ix.phys(k);
        sx[k] = d.wx[p]; sz[k] = d.wz[p]; sg[k] = d.wg[p];
        if (!UNIFORM) sv[k] = d.wvc4[p];
    }
    const double fs0 = s->fs[0], fs1 = s->fs[1];
    __syncthreads();
    cl_sync();            // every CTA holds its snapshot of the positions: they may be overwritten from here on
    const int T = (int)C * (int)blockDim.x;
    int P = 32;
    while (P > 1 && Nf * P > T) P >>= 1;
    const int G = T / P, gid = ((int)rank * (int)blockDim.x + tid) / P, sub = tid & (P - 1);
    for (long long base = 0; base < Nf; base += G) {
        const long long v = base + gid;
        const bool ok = v < Nf;
        double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
        double tx = 0.0, tz = 0.0;
        if (ok) {
            tx = sx[v]; tz = sz[v];
            int j = sub;
            for (; j + 3 * P < Ns; j += 4 * P) {
                cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
                cl_pair(sx[j + P], sz[j + P], sg[j + P], UNIFORM ? a.vc4u : sv[j + P], tx, tz, u1, w1);
                cl_pair(sx[j + 2 * P], sz[j + 2 * P], sg[j + 2 * P], UNIFORM ? a.vc4u : sv[j + 2 * P], tx, tz, u2, w2);
                cl_pair(sx[j + 3 * P], sz[j + 3 * P], sg[j + 3 * P], UNIFORM ? a.vc4u : sv[j + 3 * P], tx, tz, u3, w3);
            }
            if (j < Ns) cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
            if (j + P < Ns) cl_pair(sx[j + P], sz[j + P], sg[j + P], UNIFORM ? a.vc4u : sv[j + P], tx, tz, u1, w1);
            if (j + 2 * P < Ns) cl_pair(sx[j + 2 * P], sz[j + 2 * P], sg[j + 2 * P], UNIFORM ? a.vc4u : sv[j + 2 * P], tx, tz, u2, w2);
        }
        double su = (u0 + u1) + (u2 + u3), sw = (w0 + w1) + (w2 + w3);
        for (int o = P >> 1; o > 0; o >>= 1) {
            su += __shfl_xor_sync(0xffffffffu, su, o);
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
        }
        if (ok && sub == 0) {
            const long long p = ix.phys(v);
            const double vx = su * kInvTwoPi + fs0, vz = sw * kInvTwoPi + fs1;      // calcs.jl:252-277
            d.wvx[p] = vx;
            d.wvz[p] = vz;
            d.wx[p] = tx + a.dt * vx;                                                // calcs.jl:280-291
            d.wz[p] = tz + a.dt * vz;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Overlapped roll-up.  The wake-on-wake part of wakeroll only needs what is known once step_head is done (old
// positions and strengths; the TEV placed by the head still has strength 0), so the C - 1 helper CTAs evaluate it
// WHILE CTA 0 runs step_solve, keeping each thread's partial sums in registers; after the solve only the "late"
// sources are left -- the TEV (now with its strength), the LEV if one was shed, the bound vortices -- plus the new
// LEV as a target, which CTA 0 (idle otherwise) takes against all sources.  Not used when the solve may move old
// vortices (Spalart merging), when the TEV is placed after the solve (ldvmLin), or when the free vortices do not
// fit one pass of the helper threads; the sequential cl_rollup runs then.
// ---------------------------------------------------------------------------------------------
// sum over the P neighbouring lanes of a target group, valid on the group's first lane; fixed order
__device__ __forceinline__ double cl_group_sum(double v, int P) {
    if ((P & (P - 1)) == 0) {
        for (int o = P >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    double t = v;
    for (int o = 1; o < P; o++) t += __shfl_down_sync(0xffffffffu, v, o);
    return t;
}

struct ClEarly {
    double su, sw, tx, tz;
    long long p;          // slab index of this thread group's target
    long long te_phys;    // slab index of the TEV placed by the head (or -1)
    int P, sub;
    bool ok;
};

__device__ __forceinline__ bool cl_overlap_ok(const LdvmDev &d, unsigned C, long long Nf) {
    return C >= 2 && d.del_kind == 0 && d.driver != UNSFLOW_DRIVER_LDVMLIN && Nf <= (long long)(C - 1) * blockDim.x && Nf > 0;
}

template <bool UNIFORM>
__device__ __forceinline__ void cl_rollup_early(const ClusterArgs &a, unsigned rank, unsigned C, double *stg, ClEarly &es) {
    const LdvmDev &d = a.d;
    const int tid = threadIdx.x;
    ClIndex ix;
    ix.load(d.st);
    const int Nf = (int)ix.nfree();
    double *sx = stg, *sz = stg + kClCap, *sg = stg + 2 * kClCap, *sv = stg + 3 * kClCap;
    const int k_te = (int)ix.n[0] - 1;      // the TEV the head just placed: strength 0 until the solve writes it (never read here)
    for (int k = tid; k < Nf; k += blockDim.x) {
        const long long p = ix.phys(k);
        sx[k] = d.wx[p]; sz[k] = d.wz[p]; sg[k] = k == k_te ? 0.0 : d.wg[p];
        if (!UNIFORM) sv[k] = d.wvc4[p];
    }
    __syncthreads();
    // P lanes per target, 32 / P targets per warp: the largest P in {32, 16, 8, 7, .. 1} whose capacity covers the Nf targets
    const int nw = (int)(C - 1) * (int)(blockDim.x >> 5), lane = tid & 31;
    int P = 1;
    {
        const int cand[10] = {32, 16, 8, 7, 6, 5, 4, 3, 2, 1};
#pragma unroll
        for (int q = 9; q >= 0; q--) if ((long long)nw * (32 / cand[q]) >= Nf) P = cand[q];
    }
    const int gpw = 32 / P, g_in_w = lane / P, sub = lane - g_in_w * P;
    const int gid = (((int)(rank - 1) * (int)blockDim.x + tid) >> 5) * gpw + g_in_w;
    es.P = P; es.sub = sub;
    es.ok = g_in_w < gpw && gid < Nf;
    es.te_phys = ix.n[0] > 0 ? ix.b[0] + ix.n[0] - 1 : -1;
    es.p = es.ok ? ix.phys(gid) : 0;
    double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
    double tx = 0.0, tz = 0.0;
    if (es.ok) {
        tx = sx[gid]; tz = sz[gid];
        int j = sub;
        for (; j + 3 * P < Nf; j += 4 * P) {
            cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
            cl_pair(sx[j + P], sz[j + P], sg[j + P], UNIFORM ? a.vc4u : sv[j + P], tx, tz, u1, w1);
            cl_pair(sx[j + 2 * P], sz[j + 2 * P], sg[j + 2 * P], UNIFORM ? a.vc4u : sv[j + 2 * P], tx, tz, u2, w2);
            cl_pair(sx[j + 3 * P], sz[j + 3 * P], sg[j + 3 * P], UNIFORM ? a.vc4u : sv[j + 3 * P], tx, tz, u3, w3);
        }
        if (j < Nf) cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
        if (j + P < Nf) cl_pair(sx[j + P], sz[j + P], sg[j + P], UNIFORM ? a.vc4u : sv[j + P], tx, tz, u1, w1);
        if (j + 2 * P < Nf) cl_pair(sx[j + 2 * P], sz[j + 2 * P], sg[j + 2 * P], UNIFORM ? a.vc4u : sv[j + 2 * P], tx, tz, u2, w2);
    }
    es.su = (u0 + u1) + (u2 + u3);
    es.sw = (w0 + w1) + (w2 + w3);
    es.tx = tx; es.tz = tz;
}

// after the solve: late sources on the helpers' partial sums, the new LEV (if any) on CTA 0; convection
template <bool UNIFORM>
__device__ __forceinline__ void cl_rollup_late(const ClusterArgs &a, unsigned rank, unsigned C, double *stg, const ClEarly &es, long long n1_pre) {
    const LdvmDev &d = a.d;
    const StepState *s = d.st;
    const int tid = threadIdx.x;
    ClIndex ix;
    ix.load(s);
    const double fs0 = s->fs[0], fs1 = s->fs[1];
    double *sx = stg, *sz = stg + kClCap, *sg = stg + 2 * kClCap, *sv = stg + 3 * kClCap;
    __shared__ double s_lev[2 * 8];
    if (rank != 0) {
        const bool new_lev = ix.n[1] > n1_pre;
        // late list: [TEV of this step | LEV of this step | bound vortices]
        const int nb = (int)ix.n[3];
        const int nl = (es.te_phys >= 0 ? 1 : 0) + (new_lev ? 1 : 0) + nb;
        for (int k = tid; k < nl; k += blockDim.x) {
            int q = k;
            long long p;
            if (es.te_phys >= 0 && q == 0) p = es.te_phys;
            else {
                if (es.te_phys >= 0) q--;
                if (new_lev && q == 0) p = ix.b[1] + ix.n[1] - 1;
                else { if (new_lev) q--; p = ix.b[3] + q; }
            }
            sx[k] = d.wx[p]; sz[k] = d.wz[p]; sg[k] = d.wg[p];
            if (!UNIFORM) sv[k] = d.wvc4[p];
        }
        __syncthreads();
        cl_sync();            // every CTA holds what it needs of the old positions
        double su = es.su, sw = es.sw;
        if (es.ok) {
            double u1 = 0.0, w1 = 0.0;
            int j = es.sub;
            for (; j + es.P < nl; j += 2 * es.P) {
                cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], es.tx, es.tz, su, sw);
                cl_pair(sx[j + es.P], sz[j + es.P], sg[j + es.P], UNIFORM ? a.vc4u : sv[j + es.P], es.tx, es.tz, u1, w1);
            }
            if (j < nl) cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], es.tx, es.tz, su, sw);
            su += u1; sw += w1;
        }
        su = cl_group_sum(su, es.P);
        sw = cl_group_sum(sw, es.P);
        if (es.ok && es.sub == 0) {
            const double vx = su * kInvTwoPi + fs0, vz = sw * kInvTwoPi + fs1;      // calcs.jl:252-277
            d.wvx[es.p] = vx;
            d.wvz[es.p] = vz;
            d.wx[es.p] = es.tx + a.dt * vx;                                          // calcs.jl:280-291
            d.wz[es.p] = es.tz + a.dt * vz;
        }
    } else {
        // CTA 0: the LEV shed by this step's solve (typedefs.jl:313-348 / calcs.jl:409-426) as a target against all sources
        const bool new_lev = ix.n[1] > n1_pre;
        const int Ns = (int)(ix.nfree() + ix.n[3]);
        if (new_lev) {
            for (int k = tid; k < Ns; k += blockDim.x) {
                const long long p = ix.phys(k);
                sx[k] = d.wx[p]; sz[k] = d.wz[p]; sg[k] = d.wg[p];
                if (!UNIFORM) sv[k] = d.wvc4[p];
            }
        }
        __syncthreads();
        cl_sync();
        if (new_lev) {
            const long long p = ix.b[1] + ix.n[1] - 1;
            const int vt = (int)(ix.n[0] + ix.n[1] - 1);      // its place in the staged order
            const double tx = sx[vt], tz = sz[vt];
            double u0 = 0.0, w0 = 0.0, u1 = 0.0, w1 = 0.0;
            int j = tid;
            for (; j + (int)blockDim.x < Ns; j += 2 * blockDim.x) {
                cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
                cl_pair(sx[j + blockDim.x], sz[j + blockDim.x], sg[j + blockDim.x], UNIFORM ? a.vc4u : sv[j + blockDim.x], tx, tz, u1, w1);
            }
            if (j < Ns) cl_pair(sx[j], sz[j], sg[j], UNIFORM ? a.vc4u : sv[j], tx, tz, u0, w0);
            double su = u0 + u1, sw = w0 + w1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                sw += __shfl_xor_sync(0xffffffffu, sw, o);
            }
            if ((tid & 31) == 0) { s_lev[2 * (tid >> 5)] = su; s_lev[2 * (tid >> 5) + 1] = sw; }
            __syncthreads();
            if (tid == 0) {
                double tu = 0.0, tw = 0.0;
                for (int q = 0; q < (int)(blockDim.x >> 5); q++) { tu += s_lev[2 * q]; tw += s_lev[2 * q + 1]; }
                const double vx = tu * kInvTwoPi + fs0, vz = tw * kInvTwoPi + fs1;
                d.wvx[p] = vx;
                d.wvz[p] = vz;
                d.wx[p] = tx + a.dt * vx;
                d.wz[p] = tz + a.dt * vz;
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace unsflow
