// step_multi.cuh -- multi-surface step: ldvm(surf::Vector{TwoDSurf}) (k_mstep_head, k_mstep_solve bodies)
#pragma once
#include "step_single.cuh"

namespace unsflow {

// =============================================================================================
// Multi-surface step: ldvm(surf::Vector{TwoDSurf}, ...) solvers.jl:311-434.
// Surface arrays are laid out [array][surface][ndiv]; surface q uses offset q*ndiv (q*naterm).
// The reference's Mult functors feed back only a surface's OWN newly shed vortices
// (typedefs.jl:279-287, :375-388), so every surface solves an independent 1x1 (Kelvin) or 2x2
// (Kelvin-Kutta) affine system; cross-surface influence enters through update_indbound and
// add_indbound_b (calcs.jl:40-46) with the other surfaces' bound vortices of the previous step.
// =============================================================================================
__device__ void mstep_head(const LdvmDev &d) {
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv, ns = d.nsurf;
    __shared__ double sk[kMaxSurf][6];
    if (tid == 0) {
        const long long row = s->step < d.kin_rows ? s->step : d.kin_rows - 1;
        const double *r0 = d.kin + 9 * ns * row;
        s->t = r0[0]; s->fs[0] = r0[7]; s->fs[1] = r0[8];            // solvers.jl:313-316
        for (int q = 0; q < ns; q++) {
            const double *r = r0 + 9 * q;
            SurfState &m = d.ms[q];
            m.kin[0] = r[1]; m.kin[1] = r[2]; m.kin[2] = r[3]; m.kin[3] = r[4]; m.kin[4] = r[5]; m.kin[5] = r[6];
            for (int k = 0; k < 6; k++) sk[q][k] = m.kin[k];
        }
    }
    __syncthreads();
    for (int t = tid; t < ns * nd; t += blockDim.x) {                 // update_boundpos, calcs.jl:453-459
        const int q = t / nd;
        const SurfState &m = d.ms[q];
        const double alpha = sk[q][0], alphadot = sk[q][2], hdot = sk[q][3], u = sk[q][4];
        double sa, ca;
        sincos(alpha, &sa, &ca);
        d.bnd_x[t] = d.bnd_x[t] + d.dt * ((m.pvt * m.c - d.x[t]) * sa * alphadot - u + d.cam[t] * ca * alphadot);
        d.bnd_z[t] = d.bnd_z[t] + d.dt * (hdot + (m.pvt * m.c - d.x[t]) * ca * alphadot - d.cam[t] * sa * alphadot);
    }
    __syncthreads();
    if (tid == 0) {                                                    // place_tev multi, calcs.jl:389-406
        const long long b = s->wake.begin[0], e = s->wake.end[0];
        const long long ntev = e - b;
        for (int q = 0; q < ns; q++) {
            const SurfState &m = d.ms[q];
            const double bx = d.bnd_x[q * nd + nd - 1], bz = d.bnd_z[q * nd + nd - 1];
            double xloc, zloc;
            if (ntev < ns) { xloc = bx + 0.5 * sk[q][4] * d.dt; zloc = bz; }
            else { xloc = bx + (1. / 3.) * (d.wx[e - ns + q] - bx); zloc = bz + (1. / 3.) * (d.wz[e - ns + q] - bz); }
            const long long k = e + q;
            d.wx[k] = xloc; d.wz[k] = zloc; d.wg[k] = 0.0; d.wvc4[k] = m.vc4_new; d.wvc[k] = m.vc_new; d.wvx[k] = 0.0; d.wvz[k] = 0.0;
        }
        s->wake.end[0] = e + ns;
    }
}

// downwash of surface q at its bound point i (global index t = q*nd + i), calcs.jl:49-54
__device__ __forceinline__ double mdownwash_at(const LdvmDev &d, const SurfState &m, const double *fs, int t, double ui, double wi,
                                               double sa, double ca) {
    const double alphadot = m.kin[2], hdot = m.kin[3], u = m.kin[4];
    return -(u + fs[0]) * sa - ui * sa + (hdot - fs[1]) * ca - wi * ca - alphadot * (d.x[t] - m.pvt * m.c)
           + d.cam_slope[t] * (ui * ca + (u + fs[0]) * ca + (hdot - fs[1]) * sa - wi * sa);
}

__device__ void mstep_solve(const LdvmDev &d) {
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv, na = d.naterm, ns = d.nsurf;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    __shared__ double uind[kNdMax], wind[kNdMax], dw[kNdMax], dw1[kNdMax], dw2[kNdMax];
    __shared__ double uT[kNdMax], wT[kNdMax], uL[kNdMax], wL[kNdMax], gam[kNdMax];
    __shared__ double aterm[kNaMax + 1], dth[kNdMax], sc[32], fs[2];
    __shared__ double gts[kMaxSurf], gap[kMaxSurf], gts2[kMaxSurf], gls[kMaxSurf], gla[kMaxSurf], a0s[kMaxSurf], cK[kMaxSurf][8];
    __shared__ double trx[2 * kMaxSurf], trws[trn_ws_doubles(2 * kMaxSurf)];      // trust-region iterate and work space (thread 0)
    __shared__ int shed[kMaxSurf], nshed;
    __shared__ SurfState ms[kMaxSurf];
    const double eps_fd = 6.0554544523933395e-06;
    if (tid < 2) fs[tid] = s->fs[tid];
    if (tid < ns) ms[tid] = d.ms[tid];
    for (int i = tid; i < nd; i += blockDim.x) dth[i] = i > 0 ? d.theta[i] - d.theta[i - 1] : 0.0;
    __syncthreads();
    const long long te = s->wake.end[0];          // the ns TEVs shed this step are te-ns .. te-1
    const long long bv0 = s->wake.begin[3];
    const int nb = nd - 1;

    // ---------------- pass 1a: per surface, update_indbound + add_indbound_b, Kelvin coefficients ----------
    // The residual of surface q is affine in its own TEV strength only: K0 + g K1 - Kp + g (typedefs.jl:298-300).
    for (int q = 0; q < ns; q++) {
        const SurfState &m = ms[q];
        const double urefpi = m.uref * kPi, kfac = m.uref * m.c * kPi;
        double sa, ca;
        sincos(m.kin[0], &sa, &ca);
        const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
        for (int i = tid; i < nd; i += blockDim.x) {
            const int t = q * nd + i;
            double su = 0.0, sw = 0.0;
            for (int k = 0; k < d.S2; k++) { su += d.p2u[(long long)k * d.p2stride + t]; sw += d.p2w[(long long)k * d.p2stride + t]; }
            double u = su * kInvTwoPi, w = sw * kInvTwoPi;
            // add_indbound_b (calcs.jl:40-46): bound vortices of the other surfaces (previous step)
            for (int j = 0; j < ns; j++) {
                if (j == q) continue;
                double uj = 0.0, wj = 0.0;
                for (int b = 0; b < nb; b++) {
                    const long long k = bv0 + (long long)j * nb + b;
                    const double xd = d.wx[k] - d.bnd_x[t], zd = d.wz[k] - d.bnd_z[t];
                    const double dsq = xd * xd + zd * zd;
                    const double den = kTwoPi * sqrt(d.wvc4[k] + dsq * dsq);
                    uj = uj - d.wg[k] * zd / den;
                    wj = wj + d.wg[k] * xd / den;
                }
                u += uj; w += wj;
            }
            d.uind[t] = u; d.wind[t] = w;                          // base: everything but the new TEV
            unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
            dw[i] = mdownwash_at(d, m, fs, t, u, w, sa, ca);
            dw1[i] = -uT[i] * sa - wT[i] * ca + d.cam_slope[t] * (uT[i] * ca - wT[i] * sa);
        }
        __syncthreads();
        for (int k = warp; k < 4; k += nwarps) {
            const double v = warp_trapz_cos((k & 2) ? dw1 : dw, dth, d.costab + (k & 1) * nd, nd);
            if (lane == 0) sc[k] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi, a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
            cK[q][0] = kfac * (a00 + a10 / 2.); cK[q][1] = kfac * (a01 + a11 / 2.); cK[q][2] = kfac * (m.a0prev + d.aprev[q * na] / 2.);
        }
        __syncthreads();
    }
    // ---------------- Kelvin condition of all surfaces (solvers.jl:342-343) ----------
    if (tid == 0) {
        for (int q = 0; q < ns; q++) gts[q] = gap[q] = -(cK[q][0] - cK[q][2]) / (cK[q][1] + 1.0);
        if (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) {
            // nlsolve(not_in_place(kelv), ones(nsurf)*-0.01): the trust-region iteration on the joint system; the
            // reference is left in the state of its last functor call, the probe x - eps e_n of the last Jacobian
            // column: every surface at soln.zero except the last one
            double *xs = trx;
            for (int q = 0; q < ns; q++) xs[q] = -0.01;
            auto fk = [&](const double *x, double *r) {
                for (int q = 0; q < ns; q++) r[q] = cK[q][0] + x[q] * cK[q][1] - cK[q][2] + x[q];
            };
            dev_nlsolve_tr_n(fk, ns, xs, trws);
            for (int q = 0; q < ns; q++) gts[q] = gap[q] = xs[q];
            gap[ns - 1] = xs[ns - 1] - eps_fd * fmax(1.0, fabs(xs[ns - 1]));
        }
    }
    __syncthreads();
    // ---------------- pass 1b: the functor's side effects at the applied strengths ----------
    for (int q = 0; q < ns; q++) {
        const SurfState &m = ms[q];
        const double urefpi = m.uref * kPi;
        double sa, ca;
        sincos(m.kin[0], &sa, &ca);
        const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
        const double ga = gap[q];
        for (int i = tid; i < nd; i += blockDim.x) {
            const int t = q * nd + i;
            unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
            uind[i] = d.uind[t] + ga * uT[i]; wind[i] = d.wind[t] + ga * wT[i];
            dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
            d.uind[t] = uind[i]; d.wind[t] = wind[i]; d.downwash[t] = dw[i];
        }
        __syncthreads();
        block_trapz_cos_all(dw, dth, d.costab, nd, na, aterm);
        __syncthreads();
        if (tid == 0) {
            const double a0 = -aterm[0] / urefpi;
            a0s[q] = a0;
            ms[q].a0 = a0;
            ms[q].a0dot = (a0 - m.a0prev) / d.dt;                 // update_a2a3adot, calcs.jl:7
        }
        for (int n = tid + 1; n <= na; n += blockDim.x) {
            const double an = 2. * aterm[n] / urefpi;
            d.aterm[q * na + n - 1] = an;                          // functor: update_a0anda1 + update_a2toan
            if (n <= 3) d.adot[q * na + n - 1] = (an - d.aprev[q * na + n - 1]) / d.dt;   // calcs.jl:8-10
        }
        __syncthreads();
    }
    // ---------------- LESP test and place_lev (solvers.jl:354-368, calcs.jl:428-450) ---------------
    if (tid == 0) {
        int c = 0;
        for (int q = 0; q < ns; q++) { shed[q] = fabs(a0s[q]) > ms[q].lespcrit ? 1 : 0; c += shed[q]; }
        nshed = c;
        if (c > 0) {
            const long long le = s->wake.end[1];
            for (int q = 0; q < ns; q++) {
                const long long k = le + q;
                if (shed[q]) {
                    const SurfState &m = ms[q];
                    double sa, ca;
                    sincos(m.kin[0], &sa, &ca);
                    const double bx = d.bnd_x[q * nd], bz = d.bnd_z[q * nd];
                    double xl, zl;
                    if (m.levflag == 0) {
                        const double vx = m.kin[4] - m.kin[2] * sa * m.pvt * m.c + d.uind[q * nd];
                        const double vz = -m.kin[2] * ca * m.pvt * m.c - m.kin[3] + d.wind[q * nd];
                        xl = bx + 0.5 * vx * d.dt; zl = bz + 0.5 * vz * d.dt;
                    } else {
                        xl = bx + (1. / 3.) * (d.wx[le - ns + q] - bx); zl = bz + (1. / 3.) * (d.wz[le - ns + q] - bz);
                    }
                    d.wx[k] = xl; d.wz[k] = zl; d.wg[k] = 0.0; d.wvc4[k] = m.vc4_new; d.wvc[k] = m.vc_new;
                } else {
                    // placeholder TwoDVort(0,0,0,0,0,0) (calcs.jl:446): strength 0 contributes nothing;
                    // vc4 is kept > 0 on the device so that two coincident placeholders cannot make 0*Inf
                    d.wx[k] = 0.0; d.wz[k] = 0.0; d.wg[k] = 0.0; d.wvc4[k] = 1.0; d.wvc[k] = 0.0;
                }
                d.wvx[k] = 0.0; d.wvz[k] = 0.0;
            }
            s->wake.end[1] = le + ns;
        }
    }
    __syncthreads();
    const long long le = s->wake.end[1];
    // ---------------- pass 2a: Kelvin-Kutta coefficients (only when some surface sheds) -------------------
    // Shedding surface: K0 + g1 (K1 + 1) + g2 (K2 + 1) - Kp and A0 = a00 + a01 g1 + a02 g2 (typedefs.jl:400-411);
    // a surface that does not shed re-solves its TEV on the base the first solve left (:413-417).
    if (nshed > 0) {
        for (int q = 0; q < ns; q++) {
            const SurfState &m = ms[q];
            const double urefpi = m.uref * kPi, kfac = m.uref * m.c * kPi;
            double sa, ca;
            sincos(m.kin[0], &sa, &ca);
            const double gt0 = gts[q];
            const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
            const double xl = d.wx[le - ns + q], zl = d.wz[le - ns + q];
            const bool sh = shed[q] != 0;
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
                const double ub = d.uind[t] - gt0 * uT[i], wb = d.wind[t] - gt0 * wT[i];   // functor subtracts the STORED strength
                d.uind[t] = ub; d.wind[t] = wb;
                dw[i] = mdownwash_at(d, m, fs, t, ub, wb, sa, ca);
                dw1[i] = -uT[i] * sa - wT[i] * ca + d.cam_slope[t] * (uT[i] * ca - wT[i] * sa);
                if (sh) {
                    unit_influence(xl, zl, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uL[i], &wL[i]);
                    dw2[i] = -uL[i] * sa - wL[i] * ca + d.cam_slope[t] * (uL[i] * ca - wL[i] * sa);
                }
            }
            __syncthreads();
            for (int k = warp; k < (sh ? 6 : 4); k += nwarps) {
                const double v = warp_trapz_cos(k < 2 ? dw : (k < 4 ? dw1 : dw2), dth, d.costab + (k & 1) * nd, nd);
                if (lane == 0) sc[k] = v;
            }
            __syncthreads();
            if (tid == 0) {
                const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi, a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
                cK[q][0] = kfac * (a00 + a10 / 2.); cK[q][1] = kfac * (a01 + a11 / 2.);
                cK[q][3] = kfac * (m.a0prev + d.aprev[q * na] / 2.);
                cK[q][4] = a00; cK[q][5] = a01;
                if (sh) {
                    const double a02 = -sc[4] / urefpi, a12 = 2. * sc[5] / urefpi;
                    cK[q][2] = kfac * (a02 + a12 / 2.); cK[q][6] = a02;
                }
            }
            __syncthreads();
        }
        // ---------------- Kelvin + Kutta conditions of all surfaces (solvers.jl:371-372) ----------
        if (tid == 0) {
            int lastshed = -1;
            for (int q = 0; q < ns; q++) {
                const double *c = cK[q];
                if (shed[q]) {
                    // closed form on the branch of the A0 that triggered the shedding:
                    // (K1+1) gt + (K2+1) gl = Kp - K0 ;   a01 gt + a02 gl = lesp_cond - a00
                    const double lesp_cond = (a0s[q] > 0) ? ms[q].lespcrit : -ms[q].lespcrit;
                    const double m11 = c[1] + 1., m12 = c[2] + 1., m21 = c[5], m22 = c[6], r1 = c[3] - c[0], r2 = lesp_cond - c[4];
                    const double det = m11 * m22 - m12 * m21;
                    gts2[q] = (r1 * m22 - m12 * r2) / det; gls[q] = (m11 * r2 - r1 * m21) / det;
                    lastshed = q;
                } else {
                    gts2[q] = -(c[0] - c[3]) / (c[1] + 1.0); gls[q] = 0.0;
                }
                gla[q] = gls[q];
            }
            if (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE) {
                // nlsolve(not_in_place(kelvkutta), -0.01*ones(nshed+nsurf)): unknowns = the nsurf TEV strengths, then the
                // LEV strengths of the shedding surfaces in surface order.  The LESP target follows the sign of A0 at every
                // evaluation (typedefs.jl:404-408): two self-consistent roots per shedding surface, the iteration picks.
                double *xs = trx;
                const int n = ns + nshed;
                for (int k = 0; k < n; k++) xs[k] = -0.01;
                auto fkk = [&](const double *x, double *r) {
                    int lc = 0;
                    for (int q = 0; q < ns; q++) {
                        const double *c = cK[q];
                        if (shed[q]) {
                            const double g1 = x[q], g2 = x[ns + lc];
                            const double a0x = c[4] + c[5] * g1 + c[6] * g2;
                            r[q] = c[0] + c[1] * g1 + c[2] * g2 - c[3] + g1 + g2;
                            r[ns + lc] = a0x - (a0x > 0 ? ms[q].lespcrit : -ms[q].lespcrit);
                            lc++;
                        } else {
                            r[q] = c[0] + x[q] * c[1] - c[3] + x[q];
                        }
                    }
                };
                dev_nlsolve_tr_n(fkk, n, xs, trws);
                int lc = 0;
                for (int q = 0; q < ns; q++) {
                    gts2[q] = xs[q];
                    gls[q] = gla[q] = shed[q] ? xs[ns + lc++] : 0.0;
                }
                // last functor call = probe of the last unknown = LEV strength of the last shedding surface
                gla[lastshed] = gls[lastshed] - eps_fd * fmax(1.0, fabs(gls[lastshed]));
            }
        }
        __syncthreads();
    }
    // ---------------- pass 2b: side effects of the second solve, bv, forces -------------------
    for (int q = 0; q < ns; q++) {
        const SurfState &m = ms[q];
        const double urefpi = m.uref * kPi;
        double sa, ca;
        sincos(m.kin[0], &sa, &ca);
        double gt = gts[q], gl = 0.0;
        if (nshed > 0) {
            gt = gts2[q]; gl = gls[q];
            const double glq = gla[q];
            const bool sh = shed[q] != 0;
            const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
            const double xl = d.wx[le - ns + q], zl = d.wz[le - ns + q];
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
                double u = d.uind[t] + gt * uT[i], w = d.wind[t] + gt * wT[i];
                if (sh) {
                    unit_influence(xl, zl, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uL[i], &wL[i]);
                    u += glq * uL[i]; w += glq * wL[i];
                }
                uind[i] = u; wind[i] = w;
                dw[i] = mdownwash_at(d, m, fs, t, u, w, sa, ca);
                d.uind[t] = u; d.wind[t] = w; d.downwash[t] = dw[i];
            }
        } else {
            for (int i = tid; i < nd; i += blockDim.x) { const int t = q * nd + i; uind[i] = d.uind[t]; wind[i] = d.wind[t]; dw[i] = d.downwash[t]; }
        }
        __syncthreads();
        // final Fourier coefficients of this surface from its final downwash
        block_trapz_cos_all(dw, dth, d.costab, nd, na, aterm);
        __syncthreads();
        if (tid == 0) {
            const double a0 = -aterm[0] / urefpi;
            ms[q].a0 = a0; ms[q].a0prev = a0;                           // solvers.jl:399
            ms[q].levflag = (nshed > 0 && shed[q]) ? 1 : 0;             // :377-394
            d.wg[te - ns + q] = gt;
            if (nshed > 0 && shed[q]) d.wg[le - ns + q] = gl;
            sc[10] = a0;
        }
        for (int n = tid + 1; n <= na; n += blockDim.x) {
            const double an = 2. * aterm[n] / urefpi;
            aterm[n] = an;
            d.aterm[q * na + n - 1] = an;
            if (n <= 3) d.aprev[q * na + n - 1] = an;                   // :400-402
        }
        __syncthreads();
        const double a0 = sc[10];
        for (int ib = tid; ib < nd; ib += blockDim.x) {                 // update_bv, calcs.jl:204-219
            double g = (a0 * (1 + d.costab[nd + ib]));
            const double sth = d.sintab[nd + ib];
#pragma unroll 7
            for (int ia = 1; ia <= na; ia++) g = g + aterm[ia] * d.sintab[ia * nd + ib] * sth;
            gam[ib] = g * m.uref * m.c;
        }
        __syncthreads();
        for (int ib = tid + 1; ib < nd; ib += blockDim.x) {
            const long long k = bv0 + (long long)q * nb + ib - 1;
            d.wg[k] = (gam[ib] + gam[ib - 1]) * (d.theta[1] - d.theta[0]) / 2.;
            d.wx[k] = (d.bnd_x[q * nd + ib] + d.bnd_x[q * nd + ib - 1]) / 2.;
            d.wz[k] = (d.bnd_z[q * nd + ib] + d.bnd_z[q * nd + ib - 1]) / 2.;
        }
        if (warp == 1 % nwarps) {                                       // calc_forces, postprocess.jl:1-32
            const double hdot = m.kin[3], u = m.kin[4];
            const double a1 = aterm[1], a2 = aterm[2], a0dot = ms[q].a0dot;
            const double ad1 = d.adot[q * na], ad2 = d.adot[q * na + 1], ad3 = d.adot[q * na + 2];
            const double ufac = (u + fs[0]) * ca / m.uref + (hdot - fs[1]) * sa / m.uref;
            const double cnc = 2 * kPi * ufac * (a0 + a1 / 2.);
            const double cnnc = 2 * kPi * (3 * m.c * a0dot / (4 * m.uref) + m.c * ad1 / (4 * m.uref) + m.c * ad2 / (8 * m.uref));
            const double cs = 2 * kPi * a0 * a0;
            double nonl = 0, nonl_m = 0;
            for (int ib = lane; ib < nb; ib += 32) {
                const double gb = (gam[ib + 1] + gam[ib]) * (d.theta[1] - d.theta[0]) / 2.;
                const double v = (uind[ib] * ca - wind[ib] * sa);
                nonl = nonl + v * gb;
                nonl_m = nonl_m + v * d.x[q * nd + ib] * gb;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { nonl += __shfl_xor_sync(0xffffffffu, nonl, o); nonl_m += __shfl_xor_sync(0xffffffffu, nonl_m, o); }
            nonl = nonl * 2. / (m.uref * m.uref * m.c);
            nonl_m = nonl_m * 2. / (m.uref * m.uref * m.c * m.c);
            const double cn = cnc + cnnc + nonl;
            const double cl = cn * ca + cs * sa, cd = cn * sa - cs * ca;
            const double cm = cn * m.pvt - 2 * kPi * (ufac * (a0 / 4. + a1 / 4. - a2 / 8.)
                              + (m.c / m.uref) * (7. * a0dot / 16. + 3. * ad1 / 16. + ad2 / 16. - ad3 / 64.)) - nonl_m;
            if (lane == 0) {
                ms[q].cl = cl; ms[q].cd = cd; ms[q].cm = cm;
                double *r = d.mat + (1 + 7 * ns) * s->mat_row + 1 + 7 * q;
                r[0] = m.kin[0]; r[1] = m.kin[1]; r[2] = m.kin[4]; r[3] = a0; r[4] = cl; r[5] = cd; r[6] = cm;   // solvers.jl:427-432
            }
        }
        __syncthreads();
    }
    if (tid < ns) d.ms[tid] = ms[tid];
    if (tid == 0) {
        if (d.del_kind == UNSFLOW_DEL_SPALART) {                        // solvers.jl:406-408
            double lx = 0, lz = 0;
            const int mid = (int)nearbyint(nd / 2.0) - 1;
            for (int q = 0; q < ns; q++) { lx += d.bnd_x[q * nd + mid]; lz += d.bnd_z[q * nd + mid]; }
            lx /= ns; lz /= ns;
            if (s->wake.end[0] - s->wake.begin[0] > d.del_limit) {
                dev_spalart(d, &s->wake.begin[0], s->wake.end[0], lx, lz);
                if (s->wake.end[1] - s->wake.begin[1] > d.del_limit) dev_spalart(d, &s->wake.begin[1], s->wake.end[1], lx, lz);
            }
        }
        d.mat[(1 + 7 * ns) * s->mat_row] = s->t;
        s->mat_row += 1;
        s->step += 1;
    }
}

}  // namespace unsflow
