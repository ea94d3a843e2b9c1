// step_single.cuh -- single-surface step: ldvm / ldvm2DOF / ldvmLin / lautat (k_step_head, k_step_solve bodies)
#pragma once
#include "step_state.cuh"
#include "nlsolve_dev.cuh"

namespace unsflow {


// ---------------------------------------------------------------------------------------------
// k_step_head
// ---------------------------------------------------------------------------------------------
__device__ void dev_update_kinem2dof(const LdvmDev &d, StepState *s) {
    // update_kinem2DOF, src/lowOrder2D/calcs.jl:604-650 (k2 in KinemPar2DOF field order)
    double *k = s->k2;
    const double *p = s->sp;
    enum { alpha, h, alphadot, hdot, u, udot, alphaddot, hddot, alpha_pr, h_pr, alphadot_pr, alphadot_pr2,
           alphadot_pr3, hdot_pr, hdot_pr2, hdot_pr3, alphaddot_pr, alphaddot_pr2, alphaddot_pr3, hddot_pr,
           hddot_pr2, hddot_pr3 };
    enum { x_alpha, r_alpha, kappa, w_alpha, w_h, w_alphadot, w_hdot, cubic_h_1, cubic_h_3, cubic_alpha_1, cubic_alpha_3 };
    k[alpha_pr] = k[alpha];
    k[h_pr] = k[h];
    k[alphadot_pr3] = k[alphadot_pr2]; k[alphadot_pr2] = k[alphadot_pr]; k[alphadot_pr] = k[alphadot];
    k[hdot_pr3] = k[hdot_pr2]; k[hdot_pr2] = k[hdot_pr]; k[hdot_pr] = k[hdot];
    k[alphaddot_pr3] = k[alphaddot_pr2]; k[alphaddot_pr2] = k[alphaddot_pr];
    k[hddot_pr3] = k[hddot_pr2]; k[hddot_pr2] = k[hddot_pr];
    const double c = d.c, uref = d.uref, cl = s->cl, cm = s->cm, dt = d.dt;
    const double m11 = 2. / c;
    const double m12 = -p[x_alpha] * cos(k[alpha]);
    const double m21 = -2. * p[x_alpha] * cos(k[alpha]) / c;
    const double m22 = p[r_alpha] * p[r_alpha];
    const double R1 = 4 * p[kappa] * uref * uref * cl / (kPi * c * c)
                      - 2 * p[w_h] * p[w_h] * (p[cubic_h_1] * k[h] + p[cubic_h_3] * (k[h] * k[h] * k[h])) / c
                      - p[x_alpha] * sin(k[alpha]) * k[alphadot] * k[alphadot];
    const double R2 = 8 * p[kappa] * uref * uref * cm / (kPi * c * c)
                      - p[w_alpha] * p[w_alpha] * p[r_alpha] * p[r_alpha]
                            * (p[cubic_alpha_1] * k[alpha] + p[cubic_alpha_3] * (k[alpha] * k[alpha] * k[alpha]));
    k[hddot_pr] = (1 / (m11 * m22 - m21 * m12)) * (m22 * R1 - m12 * R2);
    k[alphaddot_pr] = (1 / (m11 * m22 - m21 * m12)) * (-m21 * R1 + m11 * R2);
    k[alphadot] = k[alphadot_pr] + (dt / 12.) * (23 * k[alphaddot_pr] - 16 * k[alphaddot_pr2] + 5 * k[alphaddot_pr3]);
    k[hdot] = k[hdot_pr] + (dt / 12) * (23 * k[hddot_pr] - 16 * k[hddot_pr2] + 5 * k[hddot_pr3]);
    k[alpha] = k[alpha_pr] + (dt / 12) * (23 * k[alphadot_pr] - 16 * k[alphadot_pr2] + 5 * k[alphadot_pr3]);
    k[h] = k[h_pr] + (dt / 12) * (23 * k[hdot_pr] - 16 * k[hdot_pr2] + 5 * k[hdot_pr3]);
    s->kin[2] = k[alphadot]; s->kin[3] = k[hdot]; s->kin[0] = k[alpha]; s->kin[1] = k[h];
}

__device__ void step_head(const LdvmDev &d) {
    // Three independent chains of dependent global loads run side by side (each costs two L2 round trips):
    // thread 0: step counter -> kinematics row;  thread 32: wake counts -> newest TEV;  all threads: geometry.
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv;      // nd <= kNdMax == blockDim.x: one bound point per thread
    __shared__ double sh_kin[6], sh_last[2], sh_te[2];
    __shared__ long long sh_be[2];
    double bx = 0.0, bz = 0.0, xi = 0.0, ci = 0.0;
    if (tid < nd) { bx = d.bnd_x[tid]; bz = d.bnd_z[tid]; xi = d.x[tid]; ci = d.cam[tid]; }
    if (tid == 32 % blockDim.x && d.driver != UNSFLOW_DRIVER_LDVMLIN) {
        const long long b = s->wake.begin[0], e = s->wake.end[0];
        sh_be[0] = b; sh_be[1] = e;
        if (e > b) { sh_last[0] = d.wx[e - 1]; sh_last[1] = d.wz[e - 1]; }
    }
    if (tid == 0) {
        const long long row = s->step < d.kin_rows ? s->step : d.kin_rows - 1;
        const double *r = d.kin + 9 * row;
        s->t = r[0];
        if (d.driver == UNSFLOW_DRIVER_LDVM2DOF) {
            s->fs[0] = r[7]; s->fs[1] = r[8];
            dev_update_kinem2dof(d, s);
        } else {
            // update_kinem, calcs.jl:97-198 (host-tabulated); update_externalvel :75-92
            s->kin[0] = r[1]; s->kin[1] = r[2]; s->kin[2] = r[3]; s->kin[3] = r[4]; s->kin[4] = r[5]; s->kin[5] = r[6];
            if (d.driver != UNSFLOW_DRIVER_LAUTAT) { s->fs[0] = r[7]; s->fs[1] = r[8]; }
        }
        for (int k = 0; k < 6; k++) sh_kin[k] = s->kin[k];
    }
    __syncthreads();
    const double alpha = sh_kin[0], alphadot = sh_kin[2], hdot = sh_kin[3], u = sh_kin[4];
    double sa, ca;
    sincos(alpha, &sa, &ca);
    // update_boundpos, calcs.jl:453-459
    if (tid < nd) {
        bx = bx + d.dt * ((d.pvt * d.c - xi) * sa * alphadot - u + ci * ca * alphadot);
        bz = bz + d.dt * (hdot + (d.pvt * d.c - xi) * ca * alphadot - ci * sa * alphadot);
        d.bnd_x[tid] = bx;
        d.bnd_z[tid] = bz;
        if (tid == nd - 1) { sh_te[0] = bx; sh_te[1] = bz; }
    }
    __syncthreads();
    if (tid == 0 && d.driver != UNSFLOW_DRIVER_LDVMLIN) {
        // place_tev, calcs.jl:375-387
        const long long b = sh_be[0], e = sh_be[1];
        double xloc, zloc;
        if (e == b) {
            xloc = sh_te[0] + 0.5 * u * d.dt;
            zloc = sh_te[1];
        } else {
            xloc = sh_te[0] + (1. / 3.) * (sh_last[0] - sh_te[0]);
            zloc = sh_te[1] + (1. / 3.) * (sh_last[1] - sh_te[1]);
        }
        d.wx[e] = xloc; d.wz[e] = zloc; d.wg[e] = 0.0;
        if (d.wvc4) { d.wvc4[e] = d.vc4_new; d.wvc[e] = d.vc_new; }      // (ensemble members: scalar core radius, no slabs)
        d.wvx[e] = 0.0; d.wvz[e] = 0.0;
        s->wake.end[0] = e + 1;
    }
}

// step_head in two halves for the cluster stepper (step_cluster.cuh), which runs the first half of step k+1 while the
// other CTAs finish the roll-up of step k.  A: kinematics row / 2DOF update, update_boundpos -- touches nothing the
// roll-up reads.  B: place_tev -- needs the convected position of the newest TEV.  hs: 4 doubles of shared memory
// (trailing-edge position and u, handed from A to B).  step_head_a + barrier + step_head_b == step_head.
__device__ void step_head_a(const LdvmDev &d, double *hs) {
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv;
    __shared__ double sh_kin[6];
    double bx = 0.0, bz = 0.0, xi = 0.0, ci = 0.0;
    if (tid < nd) { bx = d.bnd_x[tid]; bz = d.bnd_z[tid]; xi = d.x[tid]; ci = d.cam[tid]; }
    if (tid == 0) {
        const long long row = s->step < d.kin_rows ? s->step : d.kin_rows - 1;
        const double *r = d.kin + 9 * row;
        s->t = r[0];
        if (d.driver == UNSFLOW_DRIVER_LDVM2DOF) {
            s->fs[0] = r[7]; s->fs[1] = r[8];
            dev_update_kinem2dof(d, s);
        } else {
            s->kin[0] = r[1]; s->kin[1] = r[2]; s->kin[2] = r[3]; s->kin[3] = r[4]; s->kin[4] = r[5]; s->kin[5] = r[6];
            if (d.driver != UNSFLOW_DRIVER_LAUTAT) { s->fs[0] = r[7]; s->fs[1] = r[8]; }
        }
        for (int k = 0; k < 6; k++) sh_kin[k] = s->kin[k];
        hs[2] = s->kin[4];
    }
    __syncthreads();
    const double alpha = sh_kin[0], alphadot = sh_kin[2], hdot = sh_kin[3], u = sh_kin[4];
    double sa, ca;
    sincos(alpha, &sa, &ca);
    if (tid < nd) {      // update_boundpos, calcs.jl:453-459
        bx = bx + d.dt * ((d.pvt * d.c - xi) * sa * alphadot - u + ci * ca * alphadot);
        bz = bz + d.dt * (hdot + (d.pvt * d.c - xi) * ca * alphadot - ci * sa * alphadot);
        d.bnd_x[tid] = bx;
        d.bnd_z[tid] = bz;
        if (tid == nd - 1) { hs[0] = bx; hs[1] = bz; }
    }
}

__device__ void step_head_b(const LdvmDev &d, const double *hs) {
    StepState *s = d.st;
    if (threadIdx.x == 0 && d.driver != UNSFLOW_DRIVER_LDVMLIN) {
        // place_tev, calcs.jl:375-387
        const long long b = s->wake.begin[0], e = s->wake.end[0];
        double xloc, zloc;
        if (e == b) {
            xloc = hs[0] + 0.5 * hs[2] * d.dt;
            zloc = hs[1];
        } else {
            xloc = hs[0] + (1. / 3.) * (d.wx[e - 1] - hs[0]);
            zloc = hs[1] + (1. / 3.) * (d.wz[e - 1] - hs[1]);
        }
        d.wx[e] = xloc; d.wz[e] = zloc; d.wg[e] = 0.0;
        if (d.wvc4) { d.wvc4[e] = d.vc4_new; d.wvc[e] = d.vc_new; }
        d.wvx[e] = 0.0; d.wvz[e] = 0.0;
        s->wake.end[0] = e + 1;
    }
}

// ---------------------------------------------------------------------------------------------
// k_step_solve helpers
// ---------------------------------------------------------------------------------------------
// simpleTrapz(y .* cos.(n*theta), theta), utils.jl:105-115: one WARP per integral, lanes strided
// over the panels, butterfly reduction (the reference sums left to right; the reordering moves
// the result by ~1e-16 relative, far inside the 1e-8 history tolerance).  dth[i] = theta[i]-theta[i-1].
__device__ __forceinline__ double warp_trapz_cos(const double *y, const double *dth, const double *cosrow, int nd) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int i = 1 + lane; i < nd; i += 32) acc += dth[i] * (y[i] * cosrow[i] + y[i - 1] * cosrow[i - 1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc * 0.5;
}

// All integrals n = 0..na of one downwash at once: 4 consecutive lanes per integral (64 integrals per pass of a
// 256-thread block), each lane a strided quarter of the panels, two xor-shuffles to combine.  One short pass
// instead of ceil((na+1)/8) dependent warp-wide passes.
__device__ __forceinline__ void block_trapz_cos_all(const double *y, const double *dth, const double *costab, int nd, int na, double *out) {
    const int sub = threadIdx.x & 3, grp = threadIdx.x >> 2, ngrp = blockDim.x >> 2;
    for (int n0 = 0; n0 <= na; n0 += ngrp) {
        const int n = n0 + grp;
        double acc = 0.0;
        if (n <= na) {
            const double *cr = costab + (long long)n * nd;
            for (int i = 1 + sub; i < nd; i += 4) acc += dth[i] * (y[i] * cr[i] + y[i - 1] * cr[i - 1]);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (n <= na && sub == 0) out[n] = acc * 0.5;
    }
}

// unit-strength induced velocity of one vortex at (xv,zv) on bound point i, calcs.jl:469-473
__device__ __forceinline__ void unit_influence(double xv, double zv, double vc4, double bx, double bz, double *u, double *w) {
    const double xdist = xv - bx, zdist = zv - bz;
    const double distsq = xdist * xdist + zdist * zdist;
    const double den = kTwoPi * sqrt(vc4 + distsq * distsq);
    *u = -zdist / den;
    *w = xdist / den;
}

// downwash at bound point i, calcs.jl:49-54 (sa, ca = sin, cos of alpha, evaluated once per thread)
__device__ __forceinline__ double downwash_at(const LdvmDev &d, const double *kin, const double *fs, int ib, double ui, double wi,
                                              double sa, double ca) {
    const double alphadot = kin[2], hdot = kin[3], u = kin[4];
    return -(u + fs[0]) * sa - ui * sa + (hdot - fs[1]) * ca - wi * ca - alphadot * (d.x[ib] - d.pvt * d.c)
           + d.cam_slope[ib] * (ui * ca + (u + fs[0]) * ca + (hdot - fs[1]) * sa - wi * sa);
}

// one family of the Spalart merge, calcs.jl:550-572 / :577-598
__device__ void dev_spalart(const LdvmDev &d, long long *begin, long long end, double lx, double lz) {
    const long long j = *begin;
    const double gj = d.wg[j], xj = d.wx[j], zj = d.wz[j];
    const double d_j = sqrt((xj - lx) * (xj - lx) + (zj - lz) * (zj - lz));
    const double z_j = sqrt(xj * xj + zj * zj);
    for (long long i = 1; i < 20 && j + i < end; i++) {
        const long long k = j + i;
        const double gk = d.wg[k], xk = d.wx[k], zk = d.wz[k];
        const double d_k = sqrt((xk - lx) * (xk - lx) + (zk - lz) * (zk - lz));
        const double z_k = sqrt(xk * xk + zk * zk);
        const double fact = fabs(gj * gk) * fabs(z_j - z_k) / (fabs(gj + gk) * pow(d.del_dist + d_j, 1.5) * pow(d.del_dist + d_k, 1.5));
        if (fact < d.del_tol) {
            d.wx[k] = (fabs(gj) * xj + fabs(gk) * xk) / (fabs(gj + gk));
            d.wz[k] = (fabs(gj) * zj + fabs(gk) * zk) / (fabs(gj + gk));
            d.wg[k] = gk + gj;
            d.wg[j] = 0.0;          // popfirst!: dead entry contributes an exact zero
            *begin = j + 1;
            break;
        }
    }
}

// sum of the bound-sweep partial planes [k0, k1) at bound point i, in plane order: masked batches of W independent
// loads (one L2 round trip per batch)
template <int W>
__device__ __forceinline__ void plane_range_sum(const LdvmDev &d, int i, int k0, int k1, double *su_out, double *sw_out) {
    double su = 0.0, sw = 0.0;
    for (int k = k0; k < k1; k += W) {
        double bu[W], bw[W];
#pragma unroll
        for (int q = 0; q < W; q++) {
            const bool ok = k + q < k1;
            bu[q] = ok ? d.p2u[(long long)(k + q) * d.p2stride + i] : 0.0;
            bw[q] = ok ? d.p2w[(long long)(k + q) * d.p2stride + i] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < W; q++) { su += bu[q]; sw += bw[q]; }
    }
    *su_out = su;
    *sw_out = sw;
}

// costab / sintab: the [(naterm+1)][ndiv] cos(n theta) / sin(n theta) tables -- d.costab / d.sintab in global memory, or
// a shared-memory copy that a TMA bulk copy started at kernel entry is still filling (tab_bar != nullptr: wait on it
// before the first table read; the copy overlaps the partial-plane sums of phase A)
// shared-memory workspace of step_solve in doubles: 12 arrays of max(ndiv, naterm + 1) entries (rounded up to 8),
// 32 scalars, kinematics (6) + free stream (2)
__host__ __device__ inline int solve_ws_len(int nd, int na) { return ((nd > na + 1 ? nd : na + 1) + 7) / 8 * 8; }
__host__ __device__ inline int solve_ws_doubles(int nd, int na) { return 12 * solve_ws_len(nd, na) + 32 + 8; }

// ws: solve_ws_doubles(ndiv, naterm) doubles of shared memory (sized by the actual discretisation, so that an
// ensemble member CTA does not pay for kNdMax-wide arrays)
__device__ void step_solve(const LdvmDev &d, const double *costab, const double *sintab, uint64_t *tab_bar, double *ws) {
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv, na = d.naterm;
    const int wl = solve_ws_len(nd, na);
    double *uind = ws, *wind = ws + wl, *dw = ws + 2 * wl, *dw1 = ws + 3 * wl, *dw2 = ws + 4 * wl;
    double *uT = ws + 5 * wl, *wT = ws + 6 * wl, *uL = ws + 7 * wl, *wL = ws + 8 * wl, *gam = ws + 9 * wl;
    double *aterm = ws + 10 * wl;          // aterm[0] unused slot for a0-style indexing: aterm[n] = A_n
    double *dth = ws + 11 * wl;
    double *sc = ws + 12 * wl;             // scalar scratch [32]
    double *kin = sc + 32, *fs = kin + 6;
    __shared__ int lev_shed;
    const double urefpi = d.uref * kPi;
    const double kfac = d.uref * d.c * kPi;
    const bool lin = d.driver == UNSFLOW_DRIVER_LDVMLIN;
    const bool lautat = d.driver == UNSFLOW_DRIVER_LAUTAT;
    const bool twodof = d.driver == UNSFLOW_DRIVER_LDVM2DOF;
    const double eps_fd = 6.0554544523933395e-06;   // cbrt(eps(Float64))

    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    if (tid < 6) kin[tid] = s->kin[tid];
    if (tid < 2) fs[tid] = s->fs[tid];
    if (tid == 0) lev_shed = 0;
    for (int i = tid; i < nd; i += blockDim.x) dth[i] = i > 0 ? d.theta[i] - d.theta[i - 1] : 0.0;
    double sa, ca;
    sincos(s->kin[0], &sa, &ca);
    // ---- A: induced velocity at the bound points from the sweep partials (update_indbound) ----
    // More than one batch of planes (wakes from ~10^4 vortices: 148 or 296 sweep CTAs): the S2 planes are cut into Q
    // contiguous ranges summed side by side by Q groups of nd threads, then the Q range sums are added in range
    // order: fixed grouping for a given (S2, nd, block size) -> deterministic, identical on all ranks.
    {
        const int bd = blockDim.x;
        const int Q = (lautat || nd > bd || d.S2 <= 16) ? 1 : (bd / nd < 3 ? bd / nd : 3);
        if (lautat) {   // lautat never calls update_indbound: uind/wind persist (solvers.jl:76-92)
            for (int i = tid; i < nd; i += bd) { uind[i] = d.uind[i]; wind[i] = d.wind[i]; }
        } else if (Q == 1) {
            for (int i = tid; i < nd; i += bd) {
                double su, sw;
                plane_range_sum<16>(d, i, 0, d.S2, &su, &sw);
                uind[i] = su * kInvTwoPi;
                wind[i] = sw * kInvTwoPi;
            }
        } else {
            const int g = tid / nd, i = tid - g * nd;
            if (g < Q) {
                const int k0 = (int)((long long)d.S2 * g / Q), k1 = (int)((long long)d.S2 * (g + 1) / Q);
                double su, sw;
                plane_range_sum<16>(d, i, k0, k1, &su, &sw);      // (32-wide batches measured no faster)
                // range sums: group 0 -> uind / wind, 1 -> uT / wT, 2 -> uL / wL (free until phase B)
                (g == 0 ? uind : (g == 1 ? uT : uL))[i] = su;
                (g == 0 ? wind : (g == 1 ? wT : wL))[i] = sw;
            }
            __syncthreads();
            if (tid < nd) {
                double su = uind[tid] + uT[tid], sw = wind[tid] + wT[tid];
                if (Q > 2) { su += uL[tid]; sw += wL[tid]; }
                uind[tid] = su * kInvTwoPi;
                wind[tid] = sw * kInvTwoPi;
            }
        }
    }
    if (tab_bar) mbar_wait(tab_bar, 0);
    __syncthreads();

    // location of the TEV shed this step
    const long long tb = s->wake.begin[0], te = s->wake.end[0];
    double xt, zt;
    if (!lin) {
        xt = d.wx[te - 1]; zt = d.wz[te - 1];
    } else {   // ldvmLin places after the solve, solvers.jl:515-523
        if (te == tb) { xt = d.bnd_x[nd - 1] + 0.5 * kin[4] * d.dt; zt = d.bnd_z[nd - 1]; }
        else { xt = d.bnd_x[nd - 1] + (1. / 3.) * (d.wx[te - 1] - d.bnd_x[nd - 1]); zt = d.bnd_z[nd - 1] + (1. / 3.) * (d.wz[te - 1] - d.bnd_z[nd - 1]); }
    }
    const double a0prev = s->a0prev, aprev1 = d.aprev[0];

    double gt = 0.0, gl = 0.0;      // shed strengths
    if (!lin) {
        // ---- B: unit influence of the new TEV; C: affine split of the downwash ----------------
        for (int i = tid; i < nd; i += blockDim.x) {
            unit_influence(xt, zt, d.vc4_new, d.bnd_x[i], d.bnd_z[i], &uT[i], &wT[i]);
            dw[i] = downwash_at(d, kin, fs, i, uind[i], wind[i], sa, ca);
            dw1[i] = -uT[i] * sa - wT[i] * ca + d.cam_slope[i] * (uT[i] * ca - wT[i] * sa);
        }
        __syncthreads();
        // ---- D: A0, A1 as affine functions of the TEV strength --------------------------------
        for (int q = warp; q < 4; q += nwarps) {
            const double v = warp_trapz_cos((q & 2) ? dw1 : dw, dth, costab + (q & 1) * nd, nd);
            if (lane == 0) sc[q] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi;
            const double a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
            const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.);
            const double Kp = kfac * (a0prev + aprev1 / 2.);
            // Kelvin condition typedefs.jl:249-251:  K0 + g*K1 - Kp + g = 0
            double g = -(K0 - Kp) / (K1 + 1.0);
            if (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) {
                // nlsolve(not_in_place(kelv), [-0.01]) (solvers.jl:199): soln.zero is the trust-region iterate;
                // the reference is left in the state of the last finite-difference probe g - eps
                double xs[1] = {-0.01};
                auto f1 = [&](const double *x, double *r) { r[0] = K0 + x[0] * K1 - Kp + x[0]; };
                dev_nlsolve_tr<1>(f1, xs);
                g = xs[0];
            }
            sc[8] = g;
            sc[9] = (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) ? g - eps_fd * fmax(1.0, fabs(g)) : g;
        }
        __syncthreads();
        gt = sc[8];
        const double gt_applied = sc[9];
        // ---- F: functor side effects at the applied strength (typedefs.jl:240-247) --------------
        for (int i = tid; i < nd; i += blockDim.x) {
            uind[i] += gt_applied * uT[i];
            wind[i] += gt_applied * wT[i];
            dw[i] = downwash_at(d, kin, fs, i, uind[i], wind[i], sa, ca);
        }
        __syncthreads();
        // all A_n integrals of this downwash at once: A0..A3 feed update_a2a3adot (:2-12) /
        // update_atermdot (:14-24); when no LEV is shed they are also the final coefficients
        // (update_a2toan :66-72 recomputes the same numbers), so phase H is skipped then
        block_trapz_cos_all(dw, dth, costab, nd, na, aterm);
        __syncthreads();
        if (tid == 0) {
            const double a0 = -aterm[0] / urefpi;
            sc[10] = a0;
            sc[11] = (a0 - a0prev) / d.dt;                        // a0dot, calcs.jl:7
            lev_shed = (!lautat && fabs(a0) > d.lespcrit) ? 1 : 0;   // solvers.jl:208
        }
        {
            // scale every coefficient now (in place; thread 0 only reads the raw aterm[0]): without a LEV these
            // are already the final values and the barrier of phase H is saved
            const int nrate = twodof ? na : 3;
            for (int n = tid + 1; n <= na; n += blockDim.x) {
                const double an = 2. * aterm[n] / urefpi;
                aterm[n] = an;
                d.aterm[n - 1] = an;
                if (n <= nrate) d.adot[n - 1] = (an - d.aprev[n - 1]) / d.dt;     // calcs.jl:8-10 / :20-22
            }
        }
        __syncthreads();
        // (lautat saves a0prev/aprev BEFORE update_a2toan, solvers.jl:98-101: same values, a0..a3 do not change afterwards)
        if (lev_shed) {
            // ---- G: place_lev (calcs.jl:409-426) + Kelvin-Kutta 2x2 (typedefs.jl:313-348) ------
            if (tid == 0) {
                const long long lb = s->wake.begin[1], le = s->wake.end[1];
                const double le_vel_x = kin[4] - kin[2] * sa * d.pvt * d.c + uind[0];
                const double le_vel_z = -kin[2] * ca * d.pvt * d.c - kin[3] + wind[0];
                double xl, zl;
                if (s->levflag == 0 || le == lb) {
                    xl = d.bnd_x[0] + 0.5 * le_vel_x * d.dt;
                    zl = d.bnd_z[0] + 0.5 * le_vel_z * d.dt;
                } else {
                    xl = d.bnd_x[0] + (1. / 3.) * (d.wx[le - 1] - d.bnd_x[0]);
                    zl = d.bnd_z[0] + (1. / 3.) * (d.wz[le - 1] - d.bnd_z[0]);
                }
                d.wx[le] = xl; d.wz[le] = zl; d.wg[le] = 0.0;
                if (d.wvc4) { d.wvc4[le] = d.vc4_new; d.wvc[le] = d.vc_new; }
                d.wvx[le] = 0.0; d.wvz[le] = 0.0;
                s->wake.end[1] = le + 1;
                sc[12] = xl; sc[13] = zl;
            }
            __syncthreads();
            const double xl = sc[12], zl = sc[13];
            // base for the 2-D solve = current uind minus the STORED tev strength (the functor
            // subtracts ind_vel of the stored strengths, typedefs.jl:319-328)
            for (int i = tid; i < nd; i += blockDim.x) {
                unit_influence(xl, zl, d.vc4_new, d.bnd_x[i], d.bnd_z[i], &uL[i], &wL[i]);
                uind[i] -= gt * uT[i];
                wind[i] -= gt * wT[i];
                dw[i] = downwash_at(d, kin, fs, i, uind[i], wind[i], sa, ca);
                    dw2[i] = -uL[i] * sa - wL[i] * ca + d.cam_slope[i] * (uL[i] * ca - wL[i] * sa);
            }
            __syncthreads();
            for (int q = warp; q < 6; q += nwarps) {
                const double v = warp_trapz_cos(q < 2 ? dw : (q < 4 ? dw1 : dw2), dth, costab + (q & 1) * nd, nd);
                if (lane == 0) sc[q] = v;
            }
            __syncthreads();
            if (tid == 0) {
                const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi;
                const double a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
                const double a02 = -sc[4] / urefpi, a12 = 2. * sc[5] / urefpi;
                const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.), K2 = kfac * (a02 + a12 / 2.);
                const double Kp = kfac * (a0prev + aprev1 / 2.);
                const double lesp_cond = (sc[10] > 0) ? d.lespcrit : -d.lespcrit;   // typedefs.jl:340-344
                // (K1+1) gt + (K2+1) gl = Kp - K0 ;   a01 gt + a02 gl = lesp_cond - a00
                const double m11 = K1 + 1., m12 = K2 + 1., m21 = a01, m22 = a02;
                const double r1 = Kp - K0, r2 = lesp_cond - a00;
                const double det = m11 * m22 - m12 * m21;
                double g1 = (r1 * m22 - m12 * r2) / det;
                double g2 = (m11 * r2 - r1 * m21) / det;
                if (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) {
                    // nlsolve(not_in_place(kelvkutta), [-0.01; 0.01]) (solvers.jl:216).  The LESP target follows the
                    // sign of A0 at every iterate (typedefs.jl:340-344): two self-consistent roots, the iteration picks.
                    double xs[2] = {-0.01, 0.01};
                    const double lc = d.lespcrit;
                    auto f2 = [&](const double *x, double *r) {
                        const double a0x = a00 + a01 * x[0] + a02 * x[1];
                        r[0] = K0 + K1 * x[0] + K2 * x[1] - Kp + x[0] + x[1];
                        r[1] = a0x - (a0x > 0 ? lc : -lc);
                    };
                    dev_nlsolve_tr<2>(f2, xs);
                    g1 = xs[0]; g2 = xs[1];
                }
                sc[8] = g1; sc[14] = g2;
                sc[9] = g1;
                sc[15] = (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) ? g2 - eps_fd * fmax(1.0, fabs(g2)) : g2;
            }
            __syncthreads();
            gt = sc[8]; gl = sc[14];
            const double gta = sc[9], gla = sc[15];
            for (int i = tid; i < nd; i += blockDim.x) {
                uind[i] += gta * uT[i] + gla * uL[i];
                wind[i] += gta * wT[i] + gla * wL[i];
                dw[i] = downwash_at(d, kin, fs, i, uind[i], wind[i], sa, ca);
            }
            __syncthreads();
        }
        // ---- H: all Fourier coefficients from the final downwash (update_a0anda1 :57-63,
        //         update_a2toan :66-72); only the LEV branch changed the downwash since phase F ---
        if (lev_shed) {
            block_trapz_cos_all(dw, dth, costab, nd, na, aterm);
            __syncthreads();
            if (tid == 0) { sc[10] = -aterm[0] / urefpi; }
            for (int n = tid + 1; n <= na; n += blockDim.x) d.aterm[n - 1] = aterm[n] = 2. * aterm[n] / urefpi;
            __syncthreads();
        }
    } else {
        // =========================== ldvmLin, solvers.jl:505-613 =================================
        double *T1 = dw, *T2 = dw1, *T3 = dw2;
        for (int i = tid; i < nd; i += blockDim.x) {
            T1[i] = downwash_at(d, kin, fs, i, uind[i], wind[i], sa, ca);                       // :505-510
            const double xdist = d.bnd_x[i] - xt, zdist = d.bnd_z[i] - zt;             // :525-530
            const double distsq = xdist * xdist + zdist * zdist;
            T2[i] = (d.cam_slope[i] * zdist + xdist) / (kTwoPi * sqrt(distsq * distsq + d.vc4_new));
        }
        __syncthreads();
        // integrals of T1, T2 against cos(n theta), n = 0..na  -> uL (T1), wL (T2) reused as storage
        double *c1 = uL, *c2 = wL, *c3 = uT;
        block_trapz_cos_all(T1, dth, costab, nd, na, c1);
        block_trapz_cos_all(T2, dth, costab, nd, na, c2);
        __syncthreads();
        if (tid == 0) {
            // I = trapz(T.*(cos(theta)-1)) = c[1] - c[0]  (algebraically; :511, :535)
            const double I1 = d.c * (c1[1] - c1[0]);
            const double J1 = -c1[0] / urefpi;                                          // :512
            const double I2 = (c2[1] - c2[0]);
            const double J2 = -c2[0] / urefpi;                                          // :536
            const double sig_prev = -kfac * (a0prev + aprev1 / 2.);                     // :533
            const double tevstr = -(I1 + sig_prev) / (1 + I2);                          // :538
            const double a0 = J1 + J2 * tevstr;                                         // :541
            sc[0] = I1; sc[1] = J1; sc[2] = I2; sc[3] = J2; sc[4] = sig_prev;
            sc[8] = tevstr; sc[10] = a0;
            lev_shed = fabs(a0) > d.lespcrit ? 1 : 0;                                   // :553
        }
        __syncthreads();
        gt = sc[8];
        // aterm, a0dot, adot from the TEV-only solution (:542-550)
        for (int n = tid + 1; n <= na; n += blockDim.x) {
            const double an = 2. * (c1[n] + gt * c2[n]) / urefpi;
            aterm[n] = an;
            d.aterm[n - 1] = an;
            d.adot[n - 1] = (an - d.aprev[n - 1]) / d.dt;
        }
        if (tid == 0) sc[11] = (sc[10] - a0prev) / d.dt;
        __syncthreads();
        if (lev_shed) {
            if (tid == 0) {
                const long long lb = s->wake.begin[1], le = s->wake.end[1];
                const double le_vel_x = kin[4] - kin[2] * sa * d.pvt * d.c + uind[0];
                const double le_vel_z = -kin[2] * ca * d.pvt * d.c - kin[3] + wind[0];
                double xl, zl;
                if (s->levflag == 0 || le == lb) { xl = d.bnd_x[0] + 0.5 * le_vel_x * d.dt; zl = d.bnd_z[0] + 0.5 * le_vel_z * d.dt; }
                else { xl = d.bnd_x[0] + (1. / 3.) * (d.wx[le - 1] - d.bnd_x[0]); zl = d.bnd_z[0] + (1. / 3.) * (d.wz[le - 1] - d.bnd_z[0]); }
                sc[12] = xl; sc[13] = zl;
            }
            __syncthreads();
            const double xl = sc[12], zl = sc[13];
            for (int i = tid; i < nd; i += blockDim.x) {                                // :572-577
                const double xdist = d.bnd_x[i] - xl, zdist = d.bnd_z[i] - zl;
                const double distsq = xdist * xdist + zdist * zdist;
                T3[i] = (d.cam_slope[i] * zdist + xdist) / (kTwoPi * sqrt(distsq * distsq + d.vc4_new));
            }
            __syncthreads();
            block_trapz_cos_all(T3, dth, costab, nd, na, c3);
            __syncthreads();
            if (tid == 0) {
                const double I1 = sc[0], J1 = sc[1], I2 = sc[2], J2 = sc[3], sig_prev = sc[4];
                const double lesp_cond = (sc[10] >= 0.) ? d.lespcrit : -d.lespcrit;     // :554-558
                const double I3 = (c3[1] - c3[0]);                                      // :578
                const double J3 = -c3[0] / urefpi;                                      // :579
                const double det = J3 * (I2 + 1.) - J2 * (I3 + 1.);                     // :581
                const double tevstr = (-J3 * (I1 + sig_prev) + (I3 + 1) * (J1 - lesp_cond)) / det;   // :583
                const double levstr = (J2 * (I1 + sig_prev) - (I2 + 1) * (J1 - lesp_cond)) / det;    // :584
                sc[8] = tevstr; sc[14] = levstr;
                sc[10] = J1 + J2 * tevstr + J3 * levstr;                                // :587
                const long long le = s->wake.end[1];
                d.wx[le] = xl; d.wz[le] = zl; d.wg[le] = levstr;
                if (d.wvc4) { d.wvc4[le] = d.vc4_new; d.wvc[le] = d.vc_new; }
                d.wvx[le] = 0.0; d.wvz[le] = 0.0;
                s->wake.end[1] = le + 1;
            }
            __syncthreads();
            gt = sc[8]; gl = sc[14];
            for (int n = tid + 1; n <= na; n += blockDim.x) {                           // :588-601
                const double an = 2. * (c1[n] + gt * c2[n] + gl * c3[n]) / urefpi;
                aterm[n] = an;
                d.aterm[n - 1] = an;
            }
        }
        if (tid == 0) {   // push the TEV (:594 / :605)
            d.wx[te] = xt; d.wz[te] = zt; d.wg[te] = gt;
            if (d.wvc4) { d.wvc4[te] = d.vc4_new; d.wvc[te] = d.vc_new; }
            d.wvx[te] = 0.0; d.wvz[te] = 0.0;
            s->wake.end[0] = te + 1;
        }
        __syncthreads();
    }

    // ---- strengths, flags, previous values (solvers.jl:200, :217-232) ---------------------------
    if (tid == 0) {
        if (!lin) {
            d.wg[s->wake.end[0] - 1] = gt;
            if (lev_shed) d.wg[s->wake.end[1] - 1] = gl;
        }
        s->levflag = lev_shed;
        s->a0 = sc[10];
        s->a0dot = sc[11];
        s->a0prev = sc[10];
    }
    {
        const int nprev = twodof ? na : 3;
        for (int n = tid; n < nprev; n += blockDim.x) d.aprev[n] = aterm[n + 1];
    }
    for (int i = tid; i < nd; i += blockDim.x) { d.uind[i] = uind[i]; d.wind[i] = wind[i]; d.downwash[i] = dw[i]; }
    __syncthreads();

    // ---- I: update_bv, calcs.jl:204-219 -------------------------------------------------------
    const double a0 = sc[10];
    for (int ib = tid; ib < nd; ib += blockDim.x) {
        double g = (a0 * (1 + costab[nd + ib]));
        const double sth = sintab[nd + ib];
#pragma unroll 7
        for (int ia = 1; ia <= na; ia++) g = g + aterm[ia] * sintab[ia * nd + ib] * sth;
        gam[ib] = g * d.uref * d.c;
    }
    __syncthreads();
    const long long bv0 = s->wake.begin[3];
    for (int ib = tid + 1; ib < nd; ib += blockDim.x) {
        d.wg[bv0 + ib - 1] = (gam[ib] + gam[ib - 1]) * (d.theta[1] - d.theta[0]) / 2.;
        d.wx[bv0 + ib - 1] = (d.bnd_x[ib] + d.bnd_x[ib - 1]) / 2.;
        d.wz[bv0 + ib - 1] = (d.bnd_z[ib] + d.bnd_z[ib - 1]) / 2.;
    }
    __syncthreads();

    // ---- J: controlVortCount, calcs.jl:541-602 (thread 0) ; K: calc_forces postprocess.jl:1-32 --
    if (tid == 0 && d.del_kind == UNSFLOW_DEL_SPALART) {
        const int mid = (int)nearbyint(nd / 2.0) - 1;    // Int(round(ndiv/2)), solvers.jl:238
        const double lx = d.bnd_x[mid], lz = d.bnd_z[mid];
        if (s->wake.end[0] - s->wake.begin[0] > d.del_limit) {
            dev_spalart(d, &s->wake.begin[0], s->wake.end[0], lx, lz);
            if (s->wake.end[1] - s->wake.begin[1] > d.del_limit) dev_spalart(d, &s->wake.begin[1], s->wake.end[1], lx, lz);
        }
    }
    if (warp == 1 % nwarps) {
        const double hdot = kin[3], u = kin[4];
        const double a1 = aterm[1], a2 = aterm[2];
        const double a0dot = sc[11];
        const double ad1 = d.adot[0], ad2 = d.adot[1], ad3 = d.adot[2];
        const double ufac = (u + fs[0]) * ca / d.uref + (hdot - fs[1]) * sa / d.uref;
        const double cnc = 2 * kPi * ufac * (a0 + a1 / 2.);
        const double cnnc = 2 * kPi * (3 * d.c * a0dot / (4 * d.uref) + d.c * ad1 / (4 * d.uref) + d.c * ad2 / (8 * d.uref));
        const double cs = 2 * kPi * a0 * a0;
        double nonl = 0, nonl_m = 0;
        for (int ib = lane; ib < nd - 1; ib += 32) {      // postprocess.jl:15-18, lanes strided + butterfly
            const double gb = (gam[ib + 1] + gam[ib]) * (d.theta[1] - d.theta[0]) / 2.;
            const double v = (uind[ib] * ca - wind[ib] * sa);
            nonl = nonl + v * gb;
            nonl_m = nonl_m + v * d.x[ib] * gb;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nonl += __shfl_xor_sync(0xffffffffu, nonl, o);
            nonl_m += __shfl_xor_sync(0xffffffffu, nonl_m, o);
        }
        nonl = nonl * 2. / (d.uref * d.uref * d.c);
        nonl_m = nonl_m * 2. / (d.uref * d.uref * d.c * d.c);
        const double cn = cnc + cnnc + nonl;
        const double cl = cn * ca + cs * sa;
        const double cd = cn * sa - cs * ca;
        const double cm = cn * d.pvt
                          - 2 * kPi * (ufac * (a0 / 4. + a1 / 4. - a2 / 8.)
                                       + (d.c / d.uref) * (7. * a0dot / 16. + 3. * ad1 / 16. + ad2 / 16. - ad3 / 64.))
                          - nonl_m;
        if (lane == 0) {
            s->cl = cl; s->cd = cd; s->cm = cm;
            double *m = d.mat + 8 * s->mat_row;
            m[0] = s->t; m[1] = kin[0]; m[2] = kin[1]; m[3] = kin[4]; m[4] = a0; m[5] = cl; m[6] = cd; m[7] = cm;   // solvers.jl:255
            s->mat_row += 1;
            s->step += 1;
        }
    }
}

}  // namespace unsflow
