// ldvm.cu -- device-resident LDVM time stepper (ldvm / ldvm2DOF / ldvmLin / lautat).
//
// One time step of the reference's `for istep = 1:nsteps` body (src/lowOrder2D/solvers.jl
// :178-257, :697-778, :488-644, :76-131) is five launches on one stream, no host round trip:
//
//   k_step_head   (1 CTA)  kinematics row / 2DOF Adams-Bashforth, update_boundpos, place_tev
//   k_induce_direct        bound-point sweep   update_indbound          (70 x N)
//   k_step_solve  (1 CTA)  Kelvin solve, LESP test, place_lev, Kelvin-Kutta solve, Fourier
//                          coefficients + rates, update_bv, Spalart merge, calc_forces, mat row
//   k_induce_direct        wake roll-up        mutual_ind + ind_vel(bv) (N x (N+69))
//   k_induce_finish        1/2pi, free stream, explicit-Euler convection
//
// Vortex counts live in device memory (Regions); grids are sized from host-side upper bounds
// and surplus blocks exit.  Vortex families are slabs of one SoA allocation with a moving
// head (popfirst! of the Spalart merge = zero the strength and advance the head).
#include "../../include/unsflow.h"
#include "induce_sym.cuh"
#include "sweep.cuh"
#include "nccl_dl.h"
#include "runtime.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace unsflow {

constexpr int kNdMax = 256;    // max ndiv
constexpr int kNaMax = 128;    // max naterm
constexpr int kSolveThreads = 256;
// free vortices from which the pair-tile kernel wins (env UNSFLOW_SYM_THRESHOLD overrides, for tests)
static int64_t sym_threshold() {
    static int64_t v = [] { const char *e = getenv("UNSFLOW_SYM_THRESHOLD"); return e ? (int64_t)atoll(e) : (int64_t)36864; }();
    return v;
}

struct StepState {
    Regions wake;        // [0] tev, [1] lev, [2] extv, [3] bound vortices (absolute slab indices)
    Regions bnd;         // [0] = [0, ndiv): targets of the bound sweep
    long long step;      // steps completed = next row of the kinematics table
    long long mat_row;   // next row of the mat buffer
    int levflag;
    int pad_;
    double kin[6];       // alpha, h, alphadot, hdot, u, udot
    double fs[2];        // curfield.u[1], curfield.w[1]
    double t;
    double a0, a0dot, a0prev;
    double cl, cd, cm;
    double k2[22];       // KinemPar2DOF
    double sp[11];       // TwoDOFPar
};

// per-surface scalars of the multi-surface driver (ldvm(surf::Vector{TwoDSurf}), solvers.jl:270-445)
struct SurfState {
    double c, uref, pvt, lespcrit, vc_new, vc4_new;
    double kin[6];
    double a0, a0dot, a0prev;
    double cl, cd, cm;
    int levflag;
    int pad_;
};
constexpr int kMaxSurf = 4;

struct LveState;
struct LdvmDev {
    StepState *st;
    SurfState *ms;       // [nsurf] (multi-surface driver only)
    LveState *lve;       // LVE driver only
    double *lvea;        // LVE panel arrays [kLveArrays][ndiv]
    int nsurf;
    int ndiv, naterm, driver, solve_mode, del_kind;
    long long del_limit;
    double del_dist, del_tol;
    double c, uref, pvt, lespcrit, dt, vc_new, vc4_new;
    double *theta, *x, *cam, *cam_slope, *bnd_x, *bnd_z, *uind, *wind, *downwash, *aterm, *adot, *aprev;
    const double *costab, *sintab;   // [(naterm+1)][ndiv]
    double *wx, *wz, *wg, *wvc4, *wvc, *wvx, *wvz;   // wake slabs
    const double *kin;
    long long kin_rows;
    double *mat;
    const double *p2u, *p2w;
    int S2;
    long long p2stride;
};

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
    StepState *s = d.st;
    const int tid = threadIdx.x;
    __shared__ double sh_kin[6];
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
    for (int i = tid; i < d.ndiv; i += blockDim.x) {
        d.bnd_x[i] = d.bnd_x[i] + d.dt * ((d.pvt * d.c - d.x[i]) * sa * alphadot - u + d.cam[i] * ca * alphadot);
        d.bnd_z[i] = d.bnd_z[i] + d.dt * (hdot + (d.pvt * d.c - d.x[i]) * ca * alphadot - d.cam[i] * sa * alphadot);
    }
    __syncthreads();
    if (tid == 0 && d.driver != UNSFLOW_DRIVER_LDVMLIN) {
        // place_tev, calcs.jl:375-387
        const long long b = s->wake.begin[0], e = s->wake.end[0];
        const int nd = d.ndiv;
        double xloc, zloc;
        if (e == b) {
            xloc = d.bnd_x[nd - 1] + 0.5 * u * d.dt;
            zloc = d.bnd_z[nd - 1];
        } else {
            xloc = d.bnd_x[nd - 1] + (1. / 3.) * (d.wx[e - 1] - d.bnd_x[nd - 1]);
            zloc = d.bnd_z[nd - 1] + (1. / 3.) * (d.wz[e - 1] - d.bnd_z[nd - 1]);
        }
        d.wx[e] = xloc; d.wz[e] = zloc; d.wg[e] = 0.0; d.wvc4[e] = d.vc4_new; d.wvc[e] = d.vc_new;
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

__device__ void step_solve(const LdvmDev &d) {
    StepState *s = d.st;
    const int tid = threadIdx.x, nd = d.ndiv, na = d.naterm;
    __shared__ double uind[kNdMax], wind[kNdMax], dw[kNdMax], dw1[kNdMax], dw2[kNdMax];
    __shared__ double uT[kNdMax], wT[kNdMax], uL[kNdMax], wL[kNdMax], gam[kNdMax];
    __shared__ double aterm[kNaMax + 1];   // aterm[0] unused slot for a0-style indexing: aterm[n] = A_n
    __shared__ double dth[kNdMax];
    __shared__ double sc[32];              // scalar scratch
    __shared__ double kin[6], fs[2];
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
    for (int i = tid; i < nd; i += blockDim.x) {
        if (lautat) {   // lautat never calls update_indbound: uind/wind persist (solvers.jl:76-92)
            uind[i] = d.uind[i];
            wind[i] = d.wind[i];
        } else {
            double su = 0.0, sw = 0.0;
            int k = 0;
            for (; k + 8 <= d.S2; k += 8) {
                double bu[8], bw[8];
#pragma unroll
                for (int q = 0; q < 8; q++) { bu[q] = d.p2u[(long long)(k + q) * d.p2stride + i]; bw[q] = d.p2w[(long long)(k + q) * d.p2stride + i]; }
#pragma unroll
                for (int q = 0; q < 8; q++) { su += bu[q]; sw += bw[q]; }
            }
            for (; k < d.S2; k++) { su += d.p2u[(long long)k * d.p2stride + i]; sw += d.p2w[(long long)k * d.p2stride + i]; }
            uind[i] = su * kInvTwoPi;
            wind[i] = sw * kInvTwoPi;
        }
    }
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
            const double v = warp_trapz_cos((q & 2) ? dw1 : dw, dth, d.costab + (q & 1) * nd, nd);
            if (lane == 0) sc[q] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi;
            const double a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
            const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.);
            const double Kp = kfac * (a0prev + aprev1 / 2.);
            // Kelvin condition typedefs.jl:249-251:  K0 + g*K1 - Kp + g = 0
            const double g = -(K0 - Kp) / (K1 + 1.0);
            sc[8] = g;
            // state the reference is left in: at the root (exact) or at NLsolve's last
            // finite-difference probe g - eps
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
        for (int n = warp; n <= na; n += nwarps) {
            const double v = warp_trapz_cos(dw, dth, d.costab + n * nd, nd);
            if (lane == 0) aterm[n] = v;
        }
        __syncthreads();
        if (tid == 0) {
            const double a0 = -aterm[0] / urefpi;
            sc[10] = a0;
            sc[11] = (a0 - a0prev) / d.dt;                        // a0dot, calcs.jl:7
            lev_shed = (!lautat && fabs(a0) > d.lespcrit) ? 1 : 0;   // solvers.jl:208
        }
        {
            const int nrate = twodof ? na : 3;
            for (int n = tid + 1; n <= nrate; n += blockDim.x) {
                const double an = 2. * aterm[n] / urefpi;
                d.aterm[n - 1] = an;
                d.adot[n - 1] = (an - d.aprev[n - 1]) / d.dt;     // calcs.jl:8-10 / :20-22
            }
        }
        __syncthreads();
        if (lautat && tid == 0) {   // lautat saves a0prev/aprev BEFORE update_a2toan (solvers.jl:98-101)
            // (values are identical either way: a0..a3 do not change afterwards)
        }
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
                d.wx[le] = xl; d.wz[le] = zl; d.wg[le] = 0.0; d.wvc4[le] = d.vc4_new; d.wvc[le] = d.vc_new;
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
                const double v = warp_trapz_cos(q < 2 ? dw : (q < 4 ? dw1 : dw2), dth, d.costab + (q & 1) * nd, nd);
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
                const double g1 = (r1 * m22 - m12 * r2) / det;
                const double g2 = (m11 * r2 - r1 * m21) / det;
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
            for (int n = warp; n <= na; n += nwarps) {
                const double v = warp_trapz_cos(dw, dth, d.costab + n * nd, nd);
                if (lane == 0) aterm[n] = v;
            }
            __syncthreads();
        }
        if (tid == 0) { sc[10] = -aterm[0] / urefpi; }
        for (int n = tid + 1; n <= na; n += blockDim.x) d.aterm[n - 1] = aterm[n] = 2. * aterm[n] / urefpi;
        __syncthreads();
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
        for (int n = warp; n <= na; n += nwarps) {
            const double v1 = warp_trapz_cos(T1, dth, d.costab + n * nd, nd);
            const double v2 = warp_trapz_cos(T2, dth, d.costab + n * nd, nd);
            if (lane == 0) { c1[n] = v1; c2[n] = v2; }
        }
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
            for (int n = warp; n <= na; n += nwarps) {
                const double v3 = warp_trapz_cos(T3, dth, d.costab + n * nd, nd);
                if (lane == 0) c3[n] = v3;
            }
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
                d.wx[le] = xl; d.wz[le] = zl; d.wg[le] = levstr; d.wvc4[le] = d.vc4_new; d.wvc[le] = d.vc_new;
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
            d.wx[te] = xt; d.wz[te] = zt; d.wg[te] = gt; d.wvc4[te] = d.vc4_new; d.wvc[te] = d.vc_new;
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
        double g = (a0 * (1 + d.costab[nd + ib]));
        const double sth = d.sintab[nd + ib];
#pragma unroll 7
        for (int ia = 1; ia <= na; ia++) g = g + aterm[ia] * d.sintab[ia * nd + ib] * sth;
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
    __shared__ double gts[kMaxSurf], a0s[kMaxSurf];
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

    // ---------------- pass 1: per surface, update_indbound + add_indbound_b + 1x1 Kelvin ----------
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
            uind[i] = u; wind[i] = w;
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
            const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.), Kp = kfac * (m.a0prev + d.aprev[q * na] / 2.);
            const double g = -(K0 - Kp) / (K1 + 1.0);
            sc[8] = g;
            // NLsolve emulation: only the LAST unknown is one probe away from the root
            sc[9] = (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE && q == ns - 1) ? g - eps_fd * fmax(1.0, fabs(g)) : g;
            gts[q] = g;
        }
        __syncthreads();
        const double ga = sc[9];
        for (int i = tid; i < nd; i += blockDim.x) {
            const int t = q * nd + i;
            uind[i] += ga * uT[i]; wind[i] += ga * wT[i];
            dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
            d.uind[t] = uind[i]; d.wind[t] = wind[i]; d.downwash[t] = dw[i];
        }
        __syncthreads();
        for (int n = warp; n <= na; n += nwarps) {
            const double v = warp_trapz_cos(dw, dth, d.costab + n * nd, nd);
            if (lane == 0) aterm[n] = v;
        }
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
    // ---------------- pass 2: Kelvin-Kutta 2x2 for shedding surfaces, bv, forces -------------------
    for (int q = 0; q < ns; q++) {
        const SurfState &m = ms[q];
        const double urefpi = m.uref * kPi, kfac = m.uref * m.c * kPi;
        double sa, ca;
        sincos(m.kin[0], &sa, &ca);
        double gt = gts[q], gl = 0.0;
        for (int i = tid; i < nd; i += blockDim.x) { const int t = q * nd + i; uind[i] = d.uind[t]; wind[i] = d.wind[t]; dw[i] = d.downwash[t]; }
        __syncthreads();
        if (nshed > 0 && shed[q]) {
            const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
            const double xl = d.wx[le - ns + q], zl = d.wz[le - ns + q];
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
                unit_influence(xl, zl, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uL[i], &wL[i]);
                uind[i] -= gt * uT[i]; wind[i] -= gt * wT[i];           // functor subtracts the STORED strength
                dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
                dw1[i] = -uT[i] * sa - wT[i] * ca + d.cam_slope[t] * (uT[i] * ca - wT[i] * sa);
                dw2[i] = -uL[i] * sa - wL[i] * ca + d.cam_slope[t] * (uL[i] * ca - wL[i] * sa);
            }
            __syncthreads();
            for (int k = warp; k < 6; k += nwarps) {
                const double v = warp_trapz_cos(k < 2 ? dw : (k < 4 ? dw1 : dw2), dth, d.costab + (k & 1) * nd, nd);
                if (lane == 0) sc[k] = v;
            }
            __syncthreads();
            if (tid == 0) {
                const double a00 = -sc[0] / urefpi, a10 = 2. * sc[1] / urefpi, a01 = -sc[2] / urefpi, a11 = 2. * sc[3] / urefpi;
                const double a02 = -sc[4] / urefpi, a12 = 2. * sc[5] / urefpi;
                const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.), K2 = kfac * (a02 + a12 / 2.);
                const double Kp = kfac * (m.a0prev + d.aprev[q * na] / 2.);
                const double lesp_cond = (a0s[q] > 0) ? m.lespcrit : -m.lespcrit;
                const double m11 = K1 + 1., m12 = K2 + 1., m21 = a01, m22 = a02, r1 = Kp - K0, r2 = lesp_cond - a00;
                const double det = m11 * m22 - m12 * m21;
                const double g1 = (r1 * m22 - m12 * r2) / det, g2 = (m11 * r2 - r1 * m21) / det;
                sc[8] = g1; sc[14] = g2;
                // last unknown of KelvinKuttaMult = LEV strength of the last shedding surface
                int lastshed = -1;
                for (int j = 0; j < ns; j++) if (shed[j]) lastshed = j;
                sc[15] = (d.solve_mode == UNSFLOW_SOLVE_NLSOLVE && q == lastshed) ? g2 - eps_fd * fmax(1.0, fabs(g2)) : g2;
            }
            __syncthreads();
            gt = sc[8]; gl = sc[14];
            const double gla = sc[15];
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                uind[i] += gt * uT[i] + gla * uL[i]; wind[i] += gt * wT[i] + gla * wL[i];
                dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
                d.uind[t] = uind[i]; d.wind[t] = wind[i]; d.downwash[t] = dw[i];
            }
            __syncthreads();
        }
        if (nshed > 0 && !shed[q]) {
            // KelvinKuttaMult re-solves the TEV of a non-shedding surface too (typedefs.jl:413-417):
            // same 1x1 system on the base left by the first solve (identical root unless that base
            // carries the stale probe offset of the NLsolve emulation)
            const double xt = d.wx[te - ns + q], zt = d.wz[te - ns + q];
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                unit_influence(xt, zt, m.vc4_new, d.bnd_x[t], d.bnd_z[t], &uT[i], &wT[i]);
                uind[i] -= gt * uT[i]; wind[i] -= gt * wT[i];
                dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
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
                const double K0 = kfac * (a00 + a10 / 2.), K1 = kfac * (a01 + a11 / 2.), Kp = kfac * (m.a0prev + d.aprev[q * na] / 2.);
                sc[8] = -(K0 - Kp) / (K1 + 1.0);
            }
            __syncthreads();
            gt = sc[8];
            for (int i = tid; i < nd; i += blockDim.x) {
                const int t = q * nd + i;
                uind[i] += gt * uT[i]; wind[i] += gt * wT[i];
                dw[i] = mdownwash_at(d, m, fs, t, uind[i], wind[i], sa, ca);
                d.uind[t] = uind[i]; d.wind[t] = wind[i]; d.downwash[t] = dw[i];
            }
            __syncthreads();
        }
        // final Fourier coefficients of this surface from its final downwash
        for (int n = warp; n <= na; n += nwarps) {
            const double v = warp_trapz_cos(dw, dth, d.costab + n * nd, nd);
            if (lane == 0) aterm[n] = v;
        }
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

__global__ void __launch_bounds__(kSolveThreads) k_mstep_head(const LdvmDev d) { mstep_head(d); }
__global__ void __launch_bounds__(kSolveThreads) k_mstep_solve(const LdvmDev d) { mstep_solve(d); }


// =============================================================================================
// LVE -- lumped-vortex-element panel solver: src/lowOrder2D/LVE.jl:6-302 with
// the helpers of calcs.jl:652-836.  The wake lives in the INERTIAL frame on the device (the
// reference's IFR_field); its body-frame twin (curfield, refreshed by vor_BFR every step,
// calcs.jl:732-754) is never materialised: the collocation points are mapped into the inertial
// frame for the sweep and the induced velocity is rotated back (a rigid transform preserves the
// Vatistas kernel).  Each iteration: k_lve_head, k_sweep (collocation points x wake),
// k_lve_solve (influence matrix, dense (ndiv x ndiv) Gauss-Jordan with partial pivoting in shared
// memory, LEV ejection logic, panel pressures and loads), then the usual roll-up + finish.
// =============================================================================================
constexpr int kLveArrays = 14;
enum { LV_DS, LV_CSD, LV_NX, LV_NZ, LV_TX, LV_TZ, LV_VX, LV_VZ, LV_CX, LV_CZ, LV_PREV, LV_CHG, LV_CIRC, LV_REL };
struct LveState {
    double gamma_crit, rho, lev_stop, t_start, te_x, te_cam, le_x, le_cam, ifr_u0, ifr_hdot0;
    double alpha, X0, Z0;    // frame of the current iteration
    long long iter;          // iterations completed (LVE.jl:60 loop index i)
};

__device__ __forceinline__ void lve_ifr(double pvt, double sa, double ca, double X0, double Z0, double x, double z, double *X, double *Z) {
    *X = (x - pvt) * ca + z * sa + X0 + pvt;        // IFR, calcs.jl:652-663
    *Z = -(x - pvt) * sa + z * ca + Z0;
}
__device__ __forceinline__ void lve_bfr(double pvt, double sa, double ca, double X0, double Z0, double X, double Z, double *x, double *z) {
    *x = (X - X0 - pvt) * ca - (Z - Z0) * sa + pvt;   // BFR, calcs.jl:665-676
    *z = (X - X0 - pvt) * sa + (Z - Z0) * ca;
}

__device__ void lve_head(const LdvmDev &d) {
    StepState *s = d.st;
    LveState *L = d.lve;
    const int tid = threadIdx.x, nd = d.ndiv, np = nd - 1;
    __shared__ double sh[4];
    if (tid == 0) {
        const long long row = s->step < d.kin_rows ? s->step : d.kin_rows - 1;
        const double *r = d.kin + 9 * row;
        s->t = r[0];
        s->kin[0] = r[1]; s->kin[1] = r[2]; s->kin[2] = r[3]; s->kin[3] = r[4]; s->kin[4] = r[5]; s->kin[5] = r[6];   // update_kinem, LVE.jl:64
        L->alpha = r[1]; L->X0 = r[7]; L->Z0 = r[8];
        sh[0] = r[1]; sh[1] = r[7]; sh[2] = r[8];
    }
    __syncthreads();
    double sa, ca;
    sincos(sh[0], &sa, &ca);
    const double X0 = sh[1], Z0 = sh[2];
    const long long bv0 = s->wake.begin[3];
    for (int j = tid; j < np; j += blockDim.x) {
        // IFR_vor (LVE.jl:67-68) and the collocation points in the inertial frame (sweep targets)
        lve_ifr(d.pvt, sa, ca, X0, Z0, d.lvea[LV_VX * nd + j], d.lvea[LV_VZ * nd + j], &d.wx[bv0 + j], &d.wz[bv0 + j]);
        lve_ifr(d.pvt, sa, ca, X0, Z0, d.lvea[LV_CX * nd + j], d.lvea[LV_CZ * nd + j], &d.bnd_x[j], &d.bnd_z[j]);
    }
    if (tid == 0) {   // place_tev2(IFR_surf, IFR_field, dtstar, t), calcs.jl:688-709
        const long long b = s->wake.begin[0], e = s->wake.end[0];
        double TE_x, TE_z, xloc, zloc;
        lve_ifr(d.pvt, sa, ca, X0, Z0, L->te_x, L->te_cam, &TE_x, &TE_z);
        if (e == b) { xloc = TE_x + 0.5 * L->ifr_u0 * d.dt; zloc = TE_z + 0.5 * L->ifr_hdot0 * d.dt; }
        else { xloc = TE_x + (1. / 3.) * (d.wx[e - 1] - TE_x); zloc = TE_z + (1. / 3.) * (d.wz[e - 1] - TE_z); }
        d.wx[e] = xloc; d.wz[e] = zloc; d.wg[e] = 0.0; d.wvc4[e] = d.vc4_new; d.wvc[e] = d.vc_new; d.wvx[e] = 0.0; d.wvz[e] = 0.0;
        s->wake.end[0] = e + 1;
    }
}

// Gauss-Jordan with partial pivoting on the augmented matrix A[n][n+1] in shared memory
// (stands in for Julia's a\RHS, LVE.jl:93,140); all threads of the CTA take part
__device__ void cta_solve(double *A, int n, int *piv_sh) {
    const int tid = threadIdx.x, ld = n + 1, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    (void)piv_sh;
    for (int c = 0; c < n; c++) {
        if (warp == 0) {   // pivot search + row swap by one warp (no CTA barrier in between)
            double best = -1.0; int bi = c;
            for (int r = c + lane; r < n; r += 32) { const double v = fabs(A[r * ld + c]); if (v > best) { best = v; bi = r; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (bi != c) for (int j = c + lane; j <= n; j += 32) { const double t = A[c * ld + j]; A[c * ld + j] = A[bi * ld + j]; A[bi * ld + j] = t; }
        }
        __syncthreads();
        const double inv = 1.0 / A[c * ld + c];
        // eliminate column c from every other row: lanes own columns right of c, each warp a strided set
        // of rows; the row loop is unrolled so the shared-memory round trips of independent rows overlap
        // (column c itself is only read and row c only read, so there is no hazard inside the step)
        for (int j = c + 1 + lane; j <= n; j += 32) {
            const double pc = A[c * ld + j] * inv;
#pragma unroll 4
            for (int r = warp; r < n; r += nwarps) {
                if (r != c) A[r * ld + j] -= A[r * ld + c] * pc;
            }
        }
        __syncthreads();
    }
    // x_i = A[i][n] / A[i][i]  (left to the caller)
}

__device__ void lve_solve(const LdvmDev &d, double *A /* dynamic smem: nd*(nd+1) */) {
    StepState *s = d.st;
    LveState *L = d.lve;
    const int tid = threadIdx.x, nd = d.ndiv, np = nd - 1, ld = nd + 1;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    __shared__ double uind[kNdMax], wind[kNdMax], relx[kNdMax], relz[kNdMax], circ[kNdMax], a1[kNdMax], prev[kNdMax], sc[16];
    __shared__ int piv_sh, lev_shed;
    double *la = d.lvea;
    const double alpha = s->kin[0], alphadot = s->kin[2], hdot = s->kin[3], u = s->kin[4];
    double sa, ca;
    sincos(alpha, &sa, &ca);
    const double X0 = L->X0, Z0 = L->Z0, t = s->t;
    const long long tb = s->wake.begin[0], te = s->wake.end[0];
    const long long iter = L->iter;
    // ---- induced velocity at the collocation points: inertial-frame sums -> body axes -----------
    for (int j = tid; j < np; j += blockDim.x) {
        double su = 0.0, sw = 0.0;
        for (int k = 0; k < d.S2; k++) { su += d.p2u[(long long)k * d.p2stride + j]; sw += d.p2w[(long long)k * d.p2stride + j]; }
        su *= kInvTwoPi; sw *= kInvTwoPi;
        double ub = su * ca - sw * sa, wb = su * sa + sw * ca;
        if (te - tb <= 1) { ub = d.uind[j]; wb = d.wind[j]; }     // LVE.jl:80: only once curfield holds vortices
        uind[j] = ub; wind[j] = wb;
        d.uind[j] = ub; d.wind[j] = wb;
        const double cx = la[LV_CX * nd + j], cz = la[LV_CZ * nd + j];
        relx[j] = (ca * u - sa * (-hdot)) + (-alphadot * cz);      // LVE.jl:86
        relz[j] = (sa * u + ca * (-hdot)) + alphadot * (cx - d.pvt);
        prev[j] = la[LV_PREV * nd + j];
    }
    __syncthreads();
    if (tid == 0) {
        lve_bfr(d.pvt, sa, ca, X0, Z0, d.wx[te - 1], d.wz[te - 1], &sc[0], &sc[1]);   // x_w, z_w  :74
        double sp = 0.0;
        for (int j = 0; j < np; j++) sp += prev[j];
        sc[2] = sp;
        lev_shed = 0;
    }
    __syncthreads();
    const double x_w = sc[0], z_w = sc[1];
    // ---- influence_coeff (calcs.jl:756-787) + RHS (LVE.jl:78-90) into the augmented matrix -------
    auto build = [&](bool lev, double x_lev, double z_lev, double gcu) {
        for (int e = tid; e < nd * nd; e += blockDim.x) {
            const int i = e / nd, j = e % nd;
            double v;
            if (i == np) v = 1.0;
            else {
                const double t_x = la[LV_CX * nd + i], t_z = la[LV_CZ * nd + i];
                double uu, ww;
                // column j of the (possibly shifted) matrix: unshifted source index
                const int js = lev ? j + 1 : j;                      // mod_influence_coeff shifts left (:818)
                if (lev && j == nd - 1) {
                    const double xd = x_lev - t_x, zd = z_lev - t_z, dsq = xd * xd + zd * zd;
                    const double den = kTwoPi * sqrt(d.vc4_new + dsq * dsq);
                    uu = -zd / den; ww = xd / den;
                } else if (js < np) {
                    const double xd = la[LV_VX * nd + js] - t_x, zd = la[LV_VZ * nd + js] - t_z, dsq = xd * xd + zd * zd;
                    uu = -zd / (kTwoPi * dsq); ww = xd / (kTwoPi * dsq);
                } else {
                    const double xd = x_w - t_x, zd = z_w - t_z, dsq = xd * xd + zd * zd;
                    const double den = kTwoPi * sqrt(d.vc4_new + dsq * dsq);
                    uu = -zd / den; ww = xd / den;
                }
                v = uu * la[LV_NX * nd + i] + ww * la[LV_NZ * nd + i];
            }
            A[i * ld + j] = v;
        }
        for (int i = tid; i < nd; i += blockDim.x) {
            double rhs;
            if (i == np) rhs = sc[2] - (lev ? gcu : 0.0);                                  // :78, :138
            else {
                rhs = -((relx[i] + uind[i]) * la[LV_NX * nd + i] + (relz[i] + wind[i]) * la[LV_NZ * nd + i]);
                if (lev) rhs -= gcu * a1[i];                                               // :136
            }
            A[i * ld + nd] = rhs;
        }
    };
    build(false, 0.0, 0.0, 0.0);
    __syncthreads();
    for (int i = tid; i < nd; i += blockDim.x) a1[i] = A[i * ld + 0];                      // first column (calcs.jl:817)
    __syncthreads();
    cta_solve(A, nd, &piv_sh);
    for (int i = tid; i < nd; i += blockDim.x) circ[i] = A[i * ld + nd] / A[i * ld + i];   // circ = a\RHS :93
    __syncthreads();
    for (int j = tid; j < np; j += blockDim.x) {                                            // :96-100
        double g1 = 0.0, g2 = 0.0;
        for (int k = 0; k < j; k++) { g1 += prev[k]; g2 += circ[k]; }
        la[LV_CHG * nd + j] = (g2 - g1) / d.dt;
    }
    const double xx = 1 - 2 * la[LV_DS * nd] / d.c;
    const double geo = u * d.c * (acos(xx) + sin(acos(xx)));
    if (tid == 0) {
        if (fabs(circ[0]) > L->gamma_crit) {                                                // :104
            double gcu;
            if (s->levflag == 0 && t < L->lev_stop) { gcu = fabs(circ[0]) / circ[0] * L->gamma_crit; L->t_start = t; }   // :105-108
            else {                                                                          // :109-115
                const double m_slp = -d.lespcrit / (L->lev_stop - L->t_start);
                const double c_slp = d.lespcrit - (m_slp * L->t_start);
                gcu = fabs(circ[0]) / circ[0] * ((m_slp * t) + c_slp) / 1.13 * geo;
            }
            // place_lev2(surf, IFR_field, dtstar, t), calcs.jl:712-730
            const long long lb = s->wake.begin[1], le = s->wake.end[1];
            double LE_x, LE_z, xl, zl;
            lve_ifr(d.pvt, sa, ca, X0, Z0, L->le_x, L->le_cam, &LE_x, &LE_z);
            const double vx = u - alphadot * sa * d.pvt * d.c + uind[0];
            const double vz = -alphadot * ca * d.pvt * d.c - hdot + wind[0];
            if (s->levflag == 0 || le == lb) { xl = LE_x + 0.5 * vx * d.dt; zl = LE_z + 0.5 * vz * d.dt; }
            else { xl = LE_x + (1. / 3.) * (d.wx[le - 1] - LE_x); zl = LE_z + (1. / 3.) * (d.wz[le - 1] - LE_z); }
            d.wx[le] = xl; d.wz[le] = zl; d.wg[le] = 0.0; d.wvc4[le] = d.vc4_new; d.wvc[le] = d.vc_new; d.wvx[le] = 0.0; d.wvz[le] = 0.0;
            s->wake.end[1] = le + 1;
            lve_bfr(d.pvt, sa, ca, X0, Z0, xl, zl, &sc[4], &sc[5]);                         // x_lev, z_lev :126
            sc[6] = gcu;
            lev_shed = 1;
        }
    }
    __syncthreads();
    if (lev_shed) {
        const double gcu = sc[6];
        build(true, sc[4], sc[5], gcu);                                                     // mod_influence_coeff + RHS :127-138
        __syncthreads();
        cta_solve(A, nd, &piv_sh);
        __shared__ double c2[kNdMax];
        for (int i = tid; i < nd; i += blockDim.x) c2[i] = A[i * ld + nd] / A[i * ld + i];  // :140
        __syncthreads();
        if (tid == 0) d.wg[s->wake.end[1] - 1] = c2[nd - 1];                                // :142
        for (int i = tid; i < nd; i += blockDim.x) circ[i] = i == 0 ? gcu : c2[i - 1];      // :144-145
        __syncthreads();
    }
    // ---- strengths (:154-162), loads (:203-232) -------------------------------------------------
    const long long bv0 = s->wake.begin[3];
    for (int j = tid; j < np; j += blockDim.x) {
        la[LV_PREV * nd + j] = circ[j];
        la[LV_CIRC * nd + j] = circ[j];
        d.wg[bv0 + j] = circ[j];
    }
    if (tid == 0) {
        la[LV_CIRC * nd + np] = circ[np];
        if (te - tb > 1) d.wg[te - 1] = circ[np];
        s->levflag = lev_shed;
    }
    if (warp == 1 % nwarps) {
        double cn = 0.0, m1 = 0.0, m2 = 0.0;
        for (int j = lane; j < np; j += 32) {
            const double dsj = la[LV_DS * nd + j];
            const double dp = L->rho * (((relx[j] + uind[j]) * la[LV_TX * nd + j] + (relz[j] + wind[j]) * la[LV_TZ * nd + j]) * circ[j] / dsj
                                        + la[LV_CHG * nd + j]);
            cn += dp * dsj;
            m1 += dp * dsj * la[LV_TX * nd + j] * (la[LV_VX * nd + j] - d.pvt);    // cosd(cam_slope) = tau_x
            m2 += dp * dsj * la[LV_TZ * nd + j] * la[LV_VZ * nd + j];               // sind(cam_slope) = tau_z
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xffffffffu, cn, o); m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o);
        }
        if (lane == 0) {
            double *m = d.mat + 9 * s->mat_row;     // 9th column: LEV shed in this iteration (LEVhist, LVE.jl:147-149)
            m[0] = t; m[1] = alpha * 180 / kPi; m[2] = s->kin[1]; m[3] = u;
            m[4] = 0.0; m[5] = 0.0; m[6] = 0.0; m[7] = 0.0; m[8] = lev_shed;
            if (iter > 0) {
                double ndem = .5 * L->rho * u * u * d.c;
                cn /= ndem;
                const double a0 = 1.13 * circ[0] / geo;
                const double cs = 2 * kPi * a0 * a0;
                ndem *= d.c;
                m[4] = a0; m[5] = cn * ca + cs * sa; m[6] = cn * sa - cs * ca; m[7] = -m1 / ndem + m2 / ndem;
                s->a0 = a0; s->cl = m[5]; s->cd = m[6]; s->cm = m[7];
            }
            s->mat_row += 1;
            s->step += 1;
            L->iter = iter + 1;
            s->fs[0] = 0.0; s->fs[1] = 0.0;     // IFR_field.u, w stay 0 (TwoDFlowField(), LVE.jl:16)
        }
    }
}

__global__ void __launch_bounds__(kSolveThreads) k_lve_head(const LdvmDev d) { lve_head(d); }
__global__ void __launch_bounds__(kSolveThreads) k_lve_solve(const LdvmDev d) {
    extern __shared__ __align__(16) double lve_dyn[];
    lve_solve(d, lve_dyn);
}

__global__ void __launch_bounds__(kSolveThreads) k_step_head(const LdvmDev d) { step_head(d); }
__global__ void __launch_bounds__(kSolveThreads) k_step_solve(const LdvmDev d) { step_solve(d); }

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

__global__ void __launch_bounds__(kSolveThreads) k_ensemble(const LdvmDev *members, long long nsteps) {
    __shared__ double tile[4 * kEnsTile];
    __shared__ LdvmDev d;
    if (threadIdx.x == 0) d = members[blockIdx.x];
    __syncthreads();
    double *p2u = const_cast<double *>(d.p2u), *p2w = const_cast<double *>(d.p2w);
    for (long long is = 0; is < nsteps; is++) {
        step_head(d);
        __syncthreads();
        if (d.driver != UNSFLOW_DRIVER_LAUTAT) {
            int S2;
            ens_sweep(d, p2u, p2w, &S2, tile);
            if (threadIdx.x == 0) d.S2 = S2;
            __syncthreads();
        }
        step_solve(d);
        __syncthreads();
        ens_rollup(d, tile);
    }
}

}  // namespace unsflow

using namespace unsflow;

static thread_local float g_batch_ms = 0.f;   // CUDA-event time of the last k_ensemble launch

struct unsflow_ldvm {
    unsflow_opts opts{};
    int ndiv = 0, naterm = 0;
    int nsurf = 1, mat_cols = 8;
    DevBuf<SurfState> ms;
    DevBuf<LveState> lves;    // LVE driver only
    DevBuf<double> lvea;
    int64_t cap_t = 0, cap_l = 0, cap_e = 0, bvpad = 0, tot = 0;
    int64_t off_t = 0, off_l = 0, off_e = 0, off_b = 0;
    int64_t ntev0 = 0, nlev0 = 0, nextv = 0;
    int64_t steps_done = 0;          // since create
    int64_t mat_cap = 0;
    bool uniform = true;
    double vc4u = 1.0;
    // tightened count knowledge (async read-back)
    int64_t known_step = 0, known_ntev_end = 0, known_nlev_end = 0;
    DevBuf<StepState> st;
    DevBuf<double> surf;      // 9*ndiv + 3*naterm
    DevBuf<double> tabs;      // costab, sintab
    DevBuf<double> wake;      // 7 slabs
    DevBuf<double> kin;       // rows x 9
    DevBuf<double> mat;       // mat_cap x 8
    DevBuf<double> p2;        // bound-sweep partials [2][S2max][ndivpad]
    DevBuf<double> p1;        // roll-up partials [2][S1max][tot]
    DevBuf<double> planes;    // symmetric kernel partial planes [ctas][2][tot] (allocated on first use)
    int ctas = 0;
    int S2max = 1, S1max = 1;
    int64_t ndivpad = 0;
    int64_t kin_rows = 0;
    LdvmDev dev{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc = nullptr;
    StepState *pinned = nullptr;
    bool readback_pending = false;
    int64_t readback_step = 0;
    int64_t launches = 0;
    float last_ms = 0;
    // multi-GPU: pair tiles of the roll-up split over ranks, everything else replicated bit-identically
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    DevBuf<double> vsum;      // [2][tot] raw plane sums exchanged by all-reduce
    // CUDA-graph cache for the launch-bound regime
    cudaGraphExec_t gexec = nullptr;
    int64_t gkey_t = -1, gkey_l = -1;
    const double *gkey_kin = nullptr;
};

static int ldvm_sync_counts(unsflow_ldvm *h, StepState *out) {
    UF_CUDA(cudaMemcpyAsync(out, h->st.p, sizeof(StepState), cudaMemcpyDeviceToHost, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" {

int unsflow_ldvm_destroy(unsflow_ldvm *h) {
    if (!h) return 0;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && nccl_api()) nccl_api()->CommDestroy(h->comm);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evc) cudaEventDestroy(h->evc);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

#define UNSFLOW_DRIVER_LVE_INTERNAL 4
static int ldvm_create_impl(unsflow_ldvm **out, int nsurf, const unsflow_surf *sfs, const unsflow_field *fl, const unsflow_opts *op,
                            bool allow_lve = false) {
    UF_CHECK(out != nullptr, "unsflow_ldvm_create: null handle pointer");
    *out = nullptr;
    UF_CHECK(sfs && fl && op, "unsflow_ldvm_create: null argument");
    UF_CHECK(nsurf >= 1 && nsurf <= kMaxSurf, "unsflow_ldvm_create: nsurf=%d outside [1,%d]", nsurf, kMaxSurf);
    const unsflow_surf *sf = &sfs[0];
    for (int q = 1; q < nsurf; q++) {
        UF_CHECK(sfs[q].ndiv == sf->ndiv && sfs[q].naterm == sf->naterm, "unsflow_ldvm_create: all surfaces must share ndiv and naterm");
        UF_CHECK(sfs[q].theta && sfs[q].x && sfs[q].cam && sfs[q].cam_slope && sfs[q].bnd_x && sfs[q].bnd_z && sfs[q].uind && sfs[q].wind && sfs[q].downwash && sfs[q].aterm && sfs[q].adot && sfs[q].aprev, "unsflow_ldvm_create: null surf array");
        UF_CHECK(sfs[q].bv.n == sf->ndiv - 1, "unsflow_ldvm_create: surf.bv must hold ndiv-1 vortices");
        for (int i = 0; i < sf->ndiv; i++) UF_CHECK(sfs[q].theta[i] == sf->theta[i], "unsflow_ldvm_create: all surfaces must share the theta grid");
    }
    UF_CHECK(nsurf == 1 || op->driver == UNSFLOW_DRIVER_LDVM, "unsflow_ldvm_create: several surfaces are supported by the ldvm driver only");
    UF_CHECK(sf->ndiv >= 3 && sf->ndiv <= kNdMax, "unsflow_ldvm_create: ndiv=%d outside [3,%d]", sf->ndiv, kNdMax);
    UF_CHECK(sf->naterm >= 3 && sf->naterm <= kNaMax, "unsflow_ldvm_create: naterm=%d outside [3,%d]", sf->naterm, kNaMax);
    UF_CHECK(op->driver >= 0 && (op->driver <= 3 || (allow_lve && op->driver == UNSFLOW_DRIVER_LVE_INTERNAL)), "unsflow_ldvm_create: bad driver %d", op->driver);
    UF_CHECK(op->solve_mode == 0 || op->solve_mode == 1, "unsflow_ldvm_create: bad solve_mode %d", op->solve_mode);
    UF_CHECK(op->delvort_kind == 0 || op->delvort_kind == 1, "unsflow_ldvm_create: bad delvort_kind %d", op->delvort_kind);
    UF_CHECK(op->max_steps >= 0 && op->dt > 0, "unsflow_ldvm_create: need max_steps >= 0 and dt > 0");
    UF_CHECK(sf->theta && sf->x && sf->cam && sf->cam_slope && sf->bnd_x && sf->bnd_z && sf->uind && sf->wind &&
                 sf->downwash && sf->aterm && sf->adot && sf->aprev,
             "unsflow_ldvm_create: null surf array");
    UF_CHECK(sf->bv.n == sf->ndiv - 1, "unsflow_ldvm_create: surf.bv must hold ndiv-1 vortices");
    if (int rc = require_device()) return rc;
    if (op->device >= 0) UF_CUDA(cudaSetDevice(op->device));
    unsflow_ldvm *h = new (std::nothrow) unsflow_ldvm();
    UF_CHECK(h != nullptr, "out of host memory");
    auto fail = [&](int rc) { unsflow_ldvm_destroy(h); return rc; };
    h->opts = *op;
    const int nd = h->ndiv = sf->ndiv, na = h->naterm = sf->naterm;
    const int ns = h->nsurf = nsurf;
    h->mat_cols = op->driver == UNSFLOW_DRIVER_LVE_INTERNAL ? 9 : (nsurf == 1 ? 8 : 1 + 7 * nsurf);
    h->ntev0 = fl->tev.n; h->nlev0 = fl->lev.n; h->nextv = fl->extv.n;
    // slabs start on pair-tile boundaries (kSymTB is a multiple of the direct kernel's kTS)
    h->cap_t = round_up(h->ntev0 + ns * op->max_steps + ns, kSymTB);
    h->cap_l = round_up(h->nlev0 + ns * op->max_steps + ns, kSymTB);
    h->cap_e = round_up(std::max<int64_t>(h->nextv, 1), kSymTB);
    h->bvpad = round_up(ns * (nd - 1), kSymTB);
    h->off_t = 0; h->off_l = h->cap_t; h->off_e = h->off_l + h->cap_l; h->off_b = h->off_e + h->cap_e;
    h->tot = h->off_b + h->bvpad;
    h->mat_cap = std::max<int64_t>(op->max_steps, 1);
    h->ndivpad = round_up(ns * nd, 8);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(2); }
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1); cudaEventCreate(&h->evc);
    if (cudaMallocHost((void **)&h->pinned, sizeof(StepState)) != cudaSuccess) { set_error("cudaMallocHost failed"); return fail(2); }
    const int64_t tot = h->tot;
    h->S2max = 256;
    h->S1max = (int)std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)2e9 / (16 * tot)));
    if (h->st.alloc(1) || h->ms.alloc(ns) || h->surf.alloc(ns * (9 * nd + 3 * na)) || h->tabs.alloc(2 * (na + 1) * nd) || h->wake.alloc(7 * tot) ||
        h->mat.alloc(h->mat_cols * h->mat_cap) || h->p2.alloc(2 * (int64_t)h->S2max * h->ndivpad) ||
        h->p1.alloc(2 * (int64_t)h->S1max * tot))
        return fail(2);
    if (h->wake.zero(h->stream) || h->mat.zero(h->stream)) return fail(2);

    // ---- host staging -------------------------------------------------------------------------
    // surface arrays laid out [array][surface][ndiv] (+ [array][surface][naterm])
    std::vector<double> hs((size_t)ns * (9 * nd + 3 * na));
    for (int q = 0; q < ns; q++) {
        const unsflow_surf *sq = &sfs[q];
        const double *sarr[9] = {sq->theta, sq->x, sq->cam, sq->cam_slope, sq->bnd_x, sq->bnd_z, sq->uind, sq->wind, sq->downwash};
        for (int k = 0; k < 9; k++) memcpy(&hs[((size_t)k * ns + q) * nd], sarr[k], sizeof(double) * nd);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)0 * ns + q) * na], sq->aterm, sizeof(double) * na);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)1 * ns + q) * na], sq->adot, sizeof(double) * na);
        memcpy(&hs[(size_t)9 * ns * nd + ((size_t)2 * ns + q) * na], sq->aprev, sizeof(double) * na);
    }
    std::vector<double> ht((size_t)(2 * (na + 1) * nd));
    for (int n = 0; n <= na; n++)
        for (int i = 0; i < nd; i++) {
            // same expressions as the reference: cos.(ia*theta) calcs.jl:4, sin(ia*theta) :209
            ht[(size_t)n * nd + i] = n == 0 ? 1.0 : std::cos(n * sf->theta[i]);
            ht[(size_t)(na + 1) * nd + (size_t)n * nd + i] = std::sin(n * sf->theta[i]);
        }
    std::vector<double> hw((size_t)(7 * tot), 0.0);
    double *wx = &hw[0], *wz = &hw[tot], *wg = &hw[2 * tot], *wv4 = &hw[3 * tot], *wvc = &hw[4 * tot], *wvx = &hw[5 * tot], *wvz = &hw[6 * tot];
    const double vc_new = 0.02 * sf->c;
    for (int64_t i = 0; i < tot; i++) { wv4[i] = vc_new * vc_new * vc_new * vc_new; wvc[i] = vc_new; }
    bool uni = true;
    auto put = [&](const unsflow_vorts &v, int64_t off) -> int {
        UF_CHECK(v.n == 0 || (v.x && v.z && v.s && v.vc), "unsflow_ldvm_create: null vortex array");
        for (int64_t i = 0; i < v.n; i++) {
            wx[off + i] = v.x[i]; wz[off + i] = v.z[i]; wg[off + i] = v.s[i];
            wvc[off + i] = v.vc[i]; wv4[off + i] = v.vc[i] * v.vc[i] * v.vc[i] * v.vc[i];
            wvx[off + i] = v.vx ? v.vx[i] : 0.0; wvz[off + i] = v.vz ? v.vz[i] : 0.0;
            uni = uni && v.vc[i] == vc_new;
        }
        return 0;
    };
    if (put(fl->tev, h->off_t) || put(fl->lev, h->off_l) || put(fl->extv, h->off_e)) return fail(1);
    for (int q = 0; q < ns; q++)
        if (put(sfs[q].bv, h->off_b + (int64_t)q * (nd - 1))) return fail(1);
    h->uniform = uni;
    h->vc4u = vc_new * vc_new * vc_new * vc_new;

    StepState s0;
    memset(&s0, 0, sizeof(s0));
    s0.wake.begin[0] = h->off_t; s0.wake.end[0] = h->off_t + h->ntev0;
    s0.wake.begin[1] = h->off_l; s0.wake.end[1] = h->off_l + h->nlev0;
    s0.wake.begin[2] = h->off_e; s0.wake.end[2] = h->off_e + h->nextv;
    s0.wake.begin[3] = h->off_b; s0.wake.end[3] = h->off_b + (int64_t)ns * (nd - 1);
    s0.bnd.begin[0] = 0; s0.bnd.end[0] = (int64_t)ns * nd;
    s0.levflag = sf->levflag;
    for (int k = 0; k < 6; k++) s0.kin[k] = sf->kinem[k];
    s0.fs[0] = fl->u; s0.fs[1] = fl->w;
    s0.a0 = sf->a0; s0.a0dot = sf->a0dot; s0.a0prev = sf->a0prev;
    *h->pinned = s0;
    h->known_step = 0; h->known_ntev_end = h->ntev0; h->known_nlev_end = h->nlev0;

    std::vector<SurfState> hms((size_t)ns);
    for (int q = 0; q < ns; q++) {
        SurfState &m = hms[q];
        memset(&m, 0, sizeof(m));
        const unsflow_surf *sq = &sfs[q];
        m.c = sq->c; m.uref = sq->uref; m.pvt = sq->pvt; m.lespcrit = sq->lespcrit;
        m.vc_new = 0.02 * sq->c; m.vc4_new = m.vc_new * m.vc_new * m.vc_new * m.vc_new;
        for (int k = 0; k < 6; k++) m.kin[k] = sq->kinem[k];
        m.a0 = sq->a0; m.a0dot = sq->a0dot; m.a0prev = sq->a0prev; m.levflag = sq->levflag;
        if (q > 0 && m.vc_new != vc_new) uni = false;
    }
    h->uniform = uni;
    cudaStream_t st = h->stream;
    if (cudaMemcpyAsync(h->ms.p, hms.data(), sizeof(SurfState) * ns, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->st.p, &s0, sizeof(s0), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->surf.p, hs.data(), sizeof(double) * hs.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->tabs.p, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(h->wake.p, hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("unsflow_ldvm_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(2);
    }

    LdvmDev &d = h->dev;
    d.st = h->st.p;
    d.ndiv = nd; d.naterm = na; d.driver = op->driver; d.solve_mode = op->solve_mode; d.del_kind = op->delvort_kind;
    d.del_limit = op->del_limit; d.del_dist = op->del_dist; d.del_tol = op->del_tol;
    d.c = sf->c; d.uref = sf->uref; d.pvt = sf->pvt; d.lespcrit = sf->lespcrit; d.dt = op->dt;
    d.vc_new = vc_new; d.vc4_new = h->vc4u;
    double *sp = h->surf.p;
    const int64_t snd = (int64_t)ns * nd, sna = (int64_t)ns * na;
    d.ms = h->ms.p; d.nsurf = ns;
    d.theta = sp; d.x = sp + snd; d.cam = sp + 2 * snd; d.cam_slope = sp + 3 * snd; d.bnd_x = sp + 4 * snd; d.bnd_z = sp + 5 * snd;
    d.uind = sp + 6 * snd; d.wind = sp + 7 * snd; d.downwash = sp + 8 * snd;
    d.aterm = sp + 9 * snd; d.adot = sp + 9 * snd + sna; d.aprev = sp + 9 * snd + 2 * sna;
    d.costab = h->tabs.p; d.sintab = h->tabs.p + (int64_t)(na + 1) * nd;
    double *wp = h->wake.p;
    d.wx = wp; d.wz = wp + tot; d.wg = wp + 2 * tot; d.wvc4 = wp + 3 * tot; d.wvc = wp + 4 * tot; d.wvx = wp + 5 * tot; d.wvz = wp + 6 * tot;
    d.kin = nullptr; d.kin_rows = 0;
    d.mat = h->mat.p;
    d.p2u = h->p2.p; d.p2w = h->p2.p + (int64_t)h->S2max * h->ndivpad; d.p2stride = h->ndivpad; d.S2 = 1;
    *out = h;
    return 0;
}

int unsflow_ldvm_create(unsflow_ldvm **out, const unsflow_surf *sf, const unsflow_field *fl, const unsflow_opts *op) {
    return ldvm_create_impl(out, 1, sf, fl, op);
}

int unsflow_ldvm_multi_create(unsflow_ldvm **out, int32_t nsurf, const unsflow_surf *surfs, const unsflow_field *fl,
                              const unsflow_opts *op) {
    return ldvm_create_impl(out, nsurf, surfs, fl, op);
}

int unsflow_lve_create(unsflow_ldvm **out, const unsflow_surf *sf, const unsflow_field *ifr_field, const unsflow_opts *op,
                       double rho, double lev_stop, double ifr_xend, double ifr_camend, double ifr_u0, double ifr_hdot0) {
    UF_CHECK(out && sf && ifr_field && op, "unsflow_lve_create: null argument");
    UF_CHECK(sf->ndiv >= 3 && (size_t)sf->ndiv * (sf->ndiv + 1) * sizeof(double) <= 200 * 1024, "unsflow_lve_create: ndiv=%d too large for the in-CTA dense solve (max 159)", sf->ndiv);
    unsflow_opts o = *op;
    o.driver = UNSFLOW_DRIVER_LVE_INTERNAL;
    o.delvort_kind = UNSFLOW_DEL_NONE;
    if (int rc = ldvm_create_impl(out, 1, sf, ifr_field, &o, true)) return rc;
    unsflow_ldvm *h = *out;
    auto fail = [&](int rc) { unsflow_ldvm_destroy(h); *out = nullptr; return rc; };
    const int nd = h->ndiv, np = nd - 1;
    if (h->lves.alloc(1) || h->lvea.alloc((int64_t)kLveArrays * nd)) return fail(2);
    std::vector<double> a((size_t)kLveArrays * nd, 0.0);
    const double d2r = kPi / 180.0;
    for (int j = 0; j < np; j++) {                         // LVE.jl:29-47
        const double dx = sf->x[j] - sf->x[j + 1], dc = sf->cam[j] - sf->cam[j + 1];
        const double ds = std::sqrt(dx * dx + dc * dc);
        const double csd = std::asin((sf->cam[j + 1] - sf->cam[j]) / ds) / d2r;
        const double sd = std::sin(csd * d2r), cd = std::cos(csd * d2r);
        a[LV_DS * nd + j] = ds; a[LV_CSD * nd + j] = csd;
        a[LV_NX * nd + j] = -sd; a[LV_NZ * nd + j] = cd; a[LV_TX * nd + j] = cd; a[LV_TZ * nd + j] = sd;
        a[LV_VX * nd + j] = sf->x[j] + .25 * ds * cd; a[LV_VZ * nd + j] = sf->cam[j] + .25 * ds * sd;
        a[LV_CX * nd + j] = sf->x[j] + .75 * ds * cd; a[LV_CZ * nd + j] = sf->cam[j] + .75 * ds * sd;
    }
    LveState L;
    memset(&L, 0, sizeof(L));
    const double x = 1 - 2 * a[LV_DS * nd] / sf->c;        // LVE.jl:50-52
    L.gamma_crit = sf->lespcrit / 1.13 * (sf->kinem[4] * sf->c * (std::acos(x) + std::sin(std::acos(x))));
    L.rho = rho; L.lev_stop = lev_stop; L.te_x = ifr_xend; L.te_cam = ifr_camend; L.le_x = sf->x[0]; L.le_cam = sf->cam[0];
    L.ifr_u0 = ifr_u0; L.ifr_hdot0 = ifr_hdot0;
    if (cudaMemcpy(h->lves.p, &L, sizeof(L), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->lvea.p, a.data(), sizeof(double) * a.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("unsflow_lve_create: upload failed");
        return fail(2);
    }
    const size_t dyn = sizeof(double) * nd * (nd + 1);
    // static + dynamic shared memory exceed the 48 KB default already at ndiv = 70: always opt in
    if (cudaFuncSetAttribute(k_lve_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
        set_error("unsflow_lve_create: cannot reserve %zu bytes of shared memory", dyn);
        return fail(2);
    }
    h->dev.lve = h->lves.p;
    h->dev.lvea = h->lvea.p;
    return 0;
}

/* LVE extras of the state: circ (ndiv: panel circulations + shed TEV strength), circChange, prevCirc (ndiv-1 used) */
int unsflow_lve_get_circ(unsflow_ldvm *h, double *circ_ndiv, double *gamma_crit) {
    UF_CHECK(h && h->lvea.p && circ_ndiv, "unsflow_lve_get_circ: not an LVE handle");
    UF_CUDA(cudaStreamSynchronize(h->stream));
    UF_CUDA(cudaMemcpy(circ_ndiv, h->lvea.p + (int64_t)LV_CIRC * h->ndiv, sizeof(double) * h->ndiv, cudaMemcpyDeviceToHost));
    if (gamma_crit) {
        LveState L;
        UF_CUDA(cudaMemcpy(&L, h->lves.p, sizeof(L), cudaMemcpyDeviceToHost));
        *gamma_crit = L.gamma_crit;
    }
    return 0;
}

int unsflow_ldvm_set_kinematics(unsflow_ldvm *h, int64_t nrows, const double *table) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_set_kinematics: null handle");
    UF_CHECK(nrows > 0 && table != nullptr, "unsflow_ldvm_set_kinematics: empty table");
    UF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->kin.alloc(9 * nrows * h->nsurf)) return 2;
    UF_CUDA(cudaMemcpyAsync(h->kin.p, table, sizeof(double) * 9 * nrows * h->nsurf, cudaMemcpyHostToDevice, h->stream));
    // rows are consumed from the start of the new table
    StepState *zero_step = h->pinned;
    (void)zero_step;
    long long z = 0;
    UF_CUDA(cudaMemcpyAsync(&h->st.p->step, &z, sizeof(z), cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // kernel params are baked into the graph
    h->kin_rows = nrows;
    h->dev.kin = h->kin.p;
    h->dev.kin_rows = nrows;
    return 0;
}

int unsflow_ldvm_set_parallel(unsflow_ldvm *h, int rank, int nranks, const void *nccl_unique_id) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_set_parallel: null handle");
    UF_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "unsflow_ldvm_set_parallel: bad rank %d / %d", rank, nranks);
    UF_CHECK(h->steps_done == 0 && h->comm == nullptr, "unsflow_ldvm_set_parallel: call once, before the first run");
    if (nranks == 1) return 0;
    UF_CHECK(nccl_unique_id != nullptr, "unsflow_ldvm_set_parallel: nccl_unique_id required");
    UF_CHECK(h->uniform, "unsflow_ldvm_set_parallel: the multi-GPU roll-up needs a uniform core radius");
    const NcclApi *nc = nccl_api();
    if (!nc) return 4;
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    if (nc->CommInitRank(&h->comm, nranks, id, rank) != ncclSuccess) { set_error("ncclCommInitRank failed"); return 4; }
    if (h->vsum.alloc(2 * h->tot) || h->vsum.zero(h->stream)) return 2;
    UF_CUDA(cudaStreamSynchronize(h->stream));
    h->rank = rank; h->nranks = nranks;
    return 0;
}

int unsflow_ldvm_set_2dof(unsflow_ldvm *h, const double *sp11, const double *k22) {
    UF_CHECK(h && sp11 && k22, "unsflow_ldvm_set_2dof: null argument");
    UF_CUDA(cudaMemcpyAsync(h->st.p->sp, sp11, sizeof(double) * 11, cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaMemcpyAsync(h->st.p->k2, k22, sizeof(double) * 22, cudaMemcpyHostToDevice, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int unsflow_ldvm_get_2dof(unsflow_ldvm *h, double *k22) {
    UF_CHECK(h && k22, "unsflow_ldvm_get_2dof: null argument");
    UF_CUDA(cudaMemcpyAsync(k22, h->st.p->k2, sizeof(double) * 22, cudaMemcpyDeviceToHost, h->stream));
    UF_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// upper bounds of the live end offsets at `ahead` steps from now (known counts + steps since)
static void step_bounds(const unsflow_ldvm *h, int64_t ahead, int64_t *ntev_ub, int64_t *nlev_ub) {
    const int64_t since = (h->steps_done - h->known_step + ahead) * h->nsurf;
    *ntev_ub = std::min(h->cap_t, h->known_ntev_end + since);
    *nlev_ub = std::min(h->cap_l, h->known_nlev_end + since);
}

// one time step = 5 launches on the handle's stream (also used under stream capture)
static int enqueue_step(unsflow_ldvm *h, int64_t ntev_ub, int64_t nlev_ub, cudaStream_t st) {
    const int driver = h->opts.driver, smc = sm_count(), order = rsqrt_order();
    const int64_t tot = h->tot, nwake_ub = ntev_ub + nlev_ub + h->nextv;
    const bool multi = h->nsurf > 1;
    const bool lve = driver == UNSFLOW_DRIVER_LVE_INTERNAL;
    if (lve) k_lve_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    else if (multi) k_mstep_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    else k_step_head<<<1, kSolveThreads, 0, st>>>(h->dev);
    h->launches++;
    if (driver != UNSFLOW_DRIVER_LAUTAT) {
        // bound-point sweep (update_indbound, calcs.jl:35-38): targets = bound points
        InducePlan p2 = plan_induce((long long)h->ndiv * h->nsurf, 1, nwake_ub, 3, smc, h->S2max);
        InduceArgs a;
        a.sx = h->dev.wx; a.sz = h->dev.wz; a.sg = h->dev.wg; a.svc4 = h->dev.wvc4;
        a.tx = h->dev.bnd_x; a.tz = h->dev.bnd_z;
        a.sreg = &h->st.p->wake; a.treg = &h->st.p->bnd; a.nsreg = 3; a.ntreg = 1;
        a.vc4_uniform = h->vc4u;
        a.pu = h->p2.p; a.pw = h->p2.p + (int64_t)h->S2max * h->ndivpad; a.tstride = h->ndivpad;
        if (lve || (!multi && sweep_supported(h->ndiv))) {
            // dedicated few-targets kernel (sweep.cuh): one partial plane per block
            const long long tsrc = (nwake_ub + kTS - 1) / kTS + 3;
            p2.S = (int)std::max<long long>(1, std::min<long long>(tsrc, h->S2max));
            if (int rc = launch_sweep(a, lve ? h->ndiv - 1 : h->ndiv, p2.S, h->uniform, st)) return rc;
        } else {
            if (int rc = launch_induce_direct(a, p2, h->uniform, order, st)) return rc;
        }
        h->dev.S2 = p2.S;
        h->launches++;
    }
    if (lve) k_lve_solve<<<1, kSolveThreads, sizeof(double) * h->ndiv * (h->ndiv + 1), st>>>(h->dev);
    else if (multi) k_mstep_solve<<<1, kSolveThreads, 0, st>>>(h->dev);
    else k_step_solve<<<1, kSolveThreads, 0, st>>>(h->dev);
    h->launches++;
    // wake roll-up (calcs.jl:222-294): targets = free vortices, sources = free + bound
    const bool sym = h->uniform && nwake_ub >= sym_threshold();
    InducePlan p1 = plan_induce(nwake_ub, 3, nwake_ub + (long long)h->nsurf * (h->ndiv - 1), 4, smc, h->S1max);
    InduceArgs a;
    a.sx = h->dev.wx; a.sz = h->dev.wz; a.sg = h->dev.wg; a.svc4 = h->dev.wvc4;
    a.tx = h->dev.wx; a.tz = h->dev.wz;
    a.sreg = &h->st.p->wake; a.treg = &h->st.p->wake; a.nsreg = 4; a.ntreg = 3;
    a.vc4_uniform = h->vc4u;
    a.pu = h->p1.p; a.pw = h->p1.p + (int64_t)h->S1max * tot; a.tstride = tot;
    if (sym) {
        if (!h->planes.p && h->planes.alloc((int64_t)sym_ctas_max() * 2 * tot)) return 2;
        const int symR = sym_rows_per_thread(nwake_ub);
        h->ctas = sym_ctas(symR);
        // zero the live column ranges of every plane (TEV, LEV, EXTV slabs)
        const int64_t lo[3] = {h->off_t, h->off_l, h->off_e};
        const int64_t hi[3] = {h->off_t + ntev_ub, h->off_l + nlev_ub, h->off_e + h->nextv};
        for (int r = 0; r < 3; r++) {
            const int64_t b = lo[r], e = std::min(round_up(hi[r], kSymTB), tot);
            if (e > b)
                UF_CUDA(cudaMemset2DAsync(h->planes.p + b, sizeof(double) * tot, 0, sizeof(double) * (e - b), (size_t)h->ctas * 2, st));
        }
        SymArgs sa;
        sa.x = h->dev.wx; sa.z = h->dev.wz; sa.g = h->dev.wg;
        sa.reg = &h->st.p->wake; sa.nreg = 3; sa.bv_reg = 3;
        sa.vc4 = h->vc4u; sa.P = h->planes.p; sa.n = tot; sa.rank = h->rank; sa.nranks = h->nranks;
        if (int rc = launch_induce_sym(sa, symR, order, st)) return rc;
    } else {
        if (int rc = launch_induce_direct(a, p1, h->uniform, order, st)) return rc;
    }
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.pu = a.pu; f.pw = a.pw; f.tstride = tot; f.S = p1.S;
    if (sym) { f.pu = h->planes.p; f.pw = h->planes.p + tot; f.tstride = 2 * tot; f.S = h->ctas; }
    f.treg = &h->st.p->wake; f.ntreg = 3;
    f.vx = h->dev.wvx; f.vz = h->dev.wvz; f.x = h->dev.wx; f.z = h->dev.wz;
    f.fs = h->st.p->fs; f.dt = h->opts.dt;
    if (sym && h->nranks > 1) {
        // every rank holds the partial velocities of ITS share of the pair tiles for all vortices:
        // plain plane sums -> NCCL all-reduce over the live slab ranges -> 1/2pi, free stream, convection.
        // The all-reduce returns identical bits on every rank, so the replicated state stays in lockstep.
        FinishArgs fr = f;
        fr.raw = 1; fr.vx = h->vsum.p; fr.vz = h->vsum.p + tot;
        if (int rc = launch_induce_finish(fr, nwake_ub, st)) return rc;
        const NcclApi *nc = nccl_api();
        if (!nc) return 4;
        const int64_t lo[3] = {h->off_t, h->off_l, h->off_e};
        const int64_t hi[3] = {h->off_t + ntev_ub, h->off_l + nlev_ub, h->off_e + h->nextv};
        if (nc->GroupStart() != ncclSuccess) { set_error("ncclGroupStart failed"); return 4; }
        for (int r = 0; r < 3; r++) {
            const int64_t cnt = std::min(hi[r], tot) - lo[r];
            if (cnt <= 0) continue;
            for (int c = 0; c < 2; c++) {
                double *p = h->vsum.p + c * tot + lo[r];
                if (nc->AllReduce(p, p, (size_t)cnt, ncclDouble, ncclSum, h->comm, st) != ncclSuccess) { set_error("ncclAllReduce failed"); return 4; }
            }
        }
        if (nc->GroupEnd() != ncclSuccess) { set_error("ncclGroupEnd failed"); return 4; }
        FinishArgs f2 = f;
        f2.pu = h->vsum.p; f2.pw = h->vsum.p + tot; f2.tstride = 0; f2.S = 1;
        if (int rc = launch_induce_finish(f2, nwake_ub, st)) return rc;
        h->launches += 3;
        return 0;
    }
    if (int rc = launch_induce_finish(f, nwake_ub, st)) return rc;
    h->launches += 2;
    return 0;
}

static void poll_readback(unsflow_ldvm *h) {
    if (h->readback_pending && cudaEventQuery(h->evc) == cudaSuccess) {
        h->readback_pending = false;
        h->known_step = h->readback_step;
        h->known_ntev_end = h->pinned->wake.end[0] - h->off_t;
        h->known_nlev_end = h->pinned->wake.end[1] - h->off_l;
    }
}

// Launch-bound regime (small wakes): replay a CUDA graph of kGraphChunk steps whose grids are
// sized for a quantised upper bound at the END of the chunk; the graph is re-captured only when
// that quantised bound (the key) changes.
constexpr int64_t kGraphChunk = 32;
constexpr int64_t kGraphQuant = 512;
static bool graphs_enabled() {
    static bool v = [] { const char *e = getenv("UNSFLOW_NO_GRAPH"); return !(e && atoi(e) != 0); }();
    return v;
}

int unsflow_ldvm_run(unsflow_ldvm *h, int64_t nsteps, double *mat_out) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_run: null handle");
    UF_CHECK(nsteps >= 0, "unsflow_ldvm_run: negative nsteps");
    UF_CHECK(nsteps == 0 || mat_out != nullptr, "unsflow_ldvm_run: null mat_out");
    UF_CHECK(h->steps_done + nsteps <= h->opts.max_steps, "unsflow_ldvm_run: %lld steps exceed opts.max_steps=%lld",
             (long long)(h->steps_done + nsteps), (long long)h->opts.max_steps);
    UF_CHECK(h->dev.kin != nullptr, "unsflow_ldvm_run: call unsflow_ldvm_set_kinematics first");
    if (nsteps == 0) return 0;
    cudaStream_t st = h->stream;
    // mat rows of this call start at row 0 of the device buffer
    long long z = 0;
    UF_CUDA(cudaMemcpyAsync(&h->st.p->mat_row, &z, sizeof(z), cudaMemcpyHostToDevice, st));
    h->launches = 0;
    UF_CUDA(cudaEventRecord(h->ev0, st));
    int64_t is = 0;
    while (is < nsteps) {
        poll_readback(h);
        int64_t tev_ub, lev_ub;
        step_bounds(h, kGraphChunk, &tev_ub, &lev_ub);
        const bool small = tev_ub + lev_ub + h->nextv < 16384;
        if (graphs_enabled() && small && nsteps - is >= kGraphChunk) {
            const int64_t qt = std::min(h->cap_t, round_up(tev_ub, kGraphQuant)), ql = std::min(h->cap_l, round_up(lev_ub, kGraphQuant));
            if (!h->gexec || h->gkey_t != qt || h->gkey_l != ql || h->gkey_kin != h->dev.kin) {
                if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
                cudaGraph_t g = nullptr;
                const int64_t l0 = h->launches;
                UF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int64_t k = 0; k < kGraphChunk && rc == 0; k++) rc = enqueue_step(h, qt, ql, st);
                cudaError_t ce = cudaStreamEndCapture(st, &g);
                h->launches = l0;
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                UF_CUDA(ce);
                UF_CUDA(cudaGraphInstantiate(&h->gexec, g, 0));
                cudaGraphDestroy(g);
                h->gkey_t = qt; h->gkey_l = ql; h->gkey_kin = h->dev.kin;
            }
            UF_CUDA(cudaGraphLaunch(h->gexec, st));
            h->launches += kGraphChunk * (h->opts.driver == UNSFLOW_DRIVER_LAUTAT ? 4 : 5);
            h->steps_done += kGraphChunk;
            is += kGraphChunk;
        } else {
            step_bounds(h, 1, &tev_ub, &lev_ub);
            if (int rc = enqueue_step(h, tev_ub, lev_ub, st)) return rc;
            UF_CUDA(cudaGetLastError());
            h->steps_done++;
            is++;
        }
        if (h->nranks > 1) {
            // grid sizes must be identical on all ranks (they fix the grouping of partial sums, hence the
            // bits): no timing-dependent read-back; a synchronous one at fixed step counts instead
            if (h->steps_done - h->readback_step >= 256) {
                StepState sx;
                if (int rc = ldvm_sync_counts(h, &sx)) return rc;
                h->known_step = h->readback_step = h->steps_done;
                h->known_ntev_end = sx.wake.end[0] - h->off_t;
                h->known_nlev_end = sx.wake.end[1] - h->off_l;
            }
        } else if (!h->readback_pending && h->steps_done - h->readback_step >= 64) {
            UF_CUDA(cudaMemcpyAsync(h->pinned, h->st.p, sizeof(StepState), cudaMemcpyDeviceToHost, st));
            UF_CUDA(cudaEventRecord(h->evc, st));
            h->readback_pending = true;
            h->readback_step = h->steps_done;
        }
    }
    UF_CUDA(cudaEventRecord(h->ev1, st));
    const int64_t tot = h->tot;
    (void)tot;
    // mat: device row-major [nsteps][8] -> caller column-major nsteps x 8
    const int mc = h->mat_cols;
    std::vector<double> hm((size_t)(mc * nsteps));
    UF_CUDA(cudaMemcpyAsync(hm.data(), h->mat.p, sizeof(double) * mc * nsteps, cudaMemcpyDeviceToHost, st));
    UF_CUDA(cudaStreamSynchronize(st));
    h->readback_pending = false;
    for (int64_t i = 0; i < nsteps; i++)
        for (int j = 0; j < mc; j++) mat_out[i + j * nsteps] = hm[(size_t)(mc * i + j)];
    UF_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    // exact counts are known now
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    h->known_step = h->steps_done;
    h->known_ntev_end = s.wake.end[0] - h->off_t;
    h->known_nlev_end = s.wake.end[1] - h->off_l;
    return 0;
}

int unsflow_ldvm_counts(unsflow_ldvm *h, int64_t *ntev, int64_t *nlev, int64_t *nextv) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_counts: null handle");
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    if (ntev) *ntev = s.wake.end[0] - s.wake.begin[0];
    if (nlev) *nlev = s.wake.end[1] - s.wake.begin[1];
    if (nextv) *nextv = s.wake.end[2] - s.wake.begin[2];
    return 0;
}

static int ldvm_get_state_impl(unsflow_ldvm *h, int nsurf, unsflow_surf *sfs, unsflow_field *fl) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_get_state: null handle");
    UF_CHECK(sfs == nullptr || nsurf == h->nsurf, "unsflow_ldvm_get_state: handle has %d surface(s), caller passed %d", h->nsurf, nsurf);
    StepState s;
    if (int rc = ldvm_sync_counts(h, &s)) return rc;
    cudaStream_t st = h->stream;
    const int nd = h->ndiv, na = h->naterm, ns = h->nsurf;
    const int64_t tot = h->tot;
    auto get = [&](unsflow_vorts *v, int64_t b, int64_t n) -> int {
        if (!v) return 0;
        UF_CHECK(v->n >= n, "unsflow_ldvm_get_state: vortex array capacity %lld < count %lld", (long long)v->n, (long long)n);
        double *dst[6] = {v->x, v->z, v->s, v->vc, v->vx, v->vz};
        const int slab[6] = {0, 1, 2, 4, 5, 6};
        for (int k = 0; k < 6; k++)
            if (dst[k] && n > 0)
                UF_CUDA(cudaMemcpyAsync(dst[k], h->wake.p + slab[k] * tot + b, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        v->n = n;
        return 0;
    };
    if (sfs) {
        std::vector<double> hs((size_t)ns * (9 * nd + 3 * na));
        std::vector<SurfState> hms((size_t)ns);
        UF_CUDA(cudaMemcpyAsync(hs.data(), h->surf.p, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, st));
        UF_CUDA(cudaMemcpyAsync(hms.data(), h->ms.p, sizeof(SurfState) * ns, cudaMemcpyDeviceToHost, st));
        UF_CUDA(cudaStreamSynchronize(st));
        for (int q = 0; q < ns; q++) {
            unsflow_surf *sf = &sfs[q];
            UF_CHECK(sf->ndiv == nd && sf->naterm == na, "unsflow_ldvm_get_state: surf dimensions differ from the handle's");
            double *sarr[9] = {nullptr, nullptr, nullptr, nullptr, sf->bnd_x, sf->bnd_z, sf->uind, sf->wind, sf->downwash};
            for (int k = 4; k < 9; k++) if (sarr[k]) memcpy(sarr[k], &hs[((size_t)k * ns + q) * nd], sizeof(double) * nd);
            if (sf->aterm) memcpy(sf->aterm, &hs[(size_t)9 * ns * nd + ((size_t)0 * ns + q) * na], sizeof(double) * na);
            if (sf->adot) memcpy(sf->adot, &hs[(size_t)9 * ns * nd + ((size_t)1 * ns + q) * na], sizeof(double) * na);
            if (sf->aprev) memcpy(sf->aprev, &hs[(size_t)9 * ns * nd + ((size_t)2 * ns + q) * na], sizeof(double) * na);
            if (ns == 1) {
                sf->levflag = s.levflag;
                for (int k = 0; k < 6; k++) sf->kinem[k] = s.kin[k];
                sf->a0 = s.a0; sf->a0dot = s.a0dot; sf->a0prev = s.a0prev;
            } else {
                sf->levflag = hms[q].levflag;
                for (int k = 0; k < 6; k++) sf->kinem[k] = hms[q].kin[k];
                sf->a0 = hms[q].a0; sf->a0dot = hms[q].a0dot; sf->a0prev = hms[q].a0prev;
            }
            if (int rc = get(&sf->bv, s.wake.begin[3] + (int64_t)q * (nd - 1), nd - 1)) return rc;
        }
    }
    if (fl) {
        for (int r = 0; r < 3; r++) {
            unsflow_vorts *v = r == 0 ? &fl->tev : (r == 1 ? &fl->lev : &fl->extv);
            if (int rc = get(v, s.wake.begin[r], s.wake.end[r] - s.wake.begin[r])) return rc;
        }
        fl->u = s.fs[0]; fl->w = s.fs[1];
    }
    UF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int unsflow_ldvm_get_state(unsflow_ldvm *h, unsflow_surf *sf, unsflow_field *fl) { return ldvm_get_state_impl(h, 1, sf, fl); }

int unsflow_ldvm_multi_get_state(unsflow_ldvm *h, int32_t nsurf, unsflow_surf *surfs, unsflow_field *fl) {
    return ldvm_get_state_impl(h, nsurf, surfs, fl);
}

int unsflow_ldvm_timing(unsflow_ldvm *h, float *ms_total, int64_t *n_launches) {
    UF_CHECK(h != nullptr, "unsflow_ldvm_timing: null handle");
    if (ms_total) *ms_total = h->last_ms;
    if (n_launches) *n_launches = h->launches;
    return 0;
}

int unsflow_ldvm_batch_run(const unsflow_surf *sf, const unsflow_opts *op, int64_t nmember, int64_t nsteps,
                           const double *tables, const double *lespcrit, double *mat_out, int64_t *ntev_out,
                           int64_t *nlev_out) {
    UF_CHECK(sf && op && tables && mat_out, "unsflow_ldvm_batch_run: null argument");
    UF_CHECK(nmember > 0 && nsteps > 0, "unsflow_ldvm_batch_run: need nmember > 0 and nsteps > 0");
    UF_CHECK(sf->ndiv >= 3 && sf->ndiv <= kNdMax && sf->naterm >= 3 && sf->naterm <= kNaMax, "unsflow_ldvm_batch_run: bad ndiv/naterm");
    UF_CHECK(op->driver == UNSFLOW_DRIVER_LDVM || op->driver == UNSFLOW_DRIVER_LDVMLIN || op->driver == UNSFLOW_DRIVER_LAUTAT,
             "unsflow_ldvm_batch_run: prescribed-motion drivers only (ldvm, ldvmLin, lautat)");
    UF_CHECK(op->dt > 0, "unsflow_ldvm_batch_run: dt must be > 0");
    UF_CHECK(sf->bv.n == sf->ndiv - 1, "unsflow_ldvm_batch_run: surf.bv must hold ndiv-1 vortices");
    if (int rc = require_device()) return rc;
    if (op->device >= 0) UF_CUDA(cudaSetDevice(op->device));
    const int nd = sf->ndiv, na = sf->naterm;
    const int64_t M = nmember;
    // per-member slabs: [TEV | LEV | EXTV(empty) | BV]
    const int64_t cap_t = round_up(nsteps + 1, 8), cap_l = round_up(nsteps + 1, 8), cap_e = 8, bvp = round_up(nd - 1, 8);
    const int64_t off_l = cap_t, off_e = off_l + cap_l, off_b = off_e + cap_e, tot = off_b + bvp;
    const int64_t nsurf = 9 * nd + 3 * na, ndp = round_up(nd, 8), np2 = 2 * 4 * ndp;
    DevBuf<StepState> st;
    DevBuf<LdvmDev> devs;
    DevBuf<double> surf, tabs, wake, kin, mat, p2;
    if (st.alloc(M) || devs.alloc(M) || surf.alloc(M * nsurf) || tabs.alloc(2 * (na + 1) * nd) || wake.alloc(M * 7 * tot) ||
        kin.alloc(M * nsteps * 9) || mat.alloc(M * nsteps * 8) || p2.alloc(M * np2))
        return 2;
    cudaStream_t stream = 0;
    // ---- host staging of ONE member, replicated ------------------------------------------------
    std::vector<double> hs((size_t)nsurf);
    const double *sarr[9] = {sf->theta, sf->x, sf->cam, sf->cam_slope, sf->bnd_x, sf->bnd_z, sf->uind, sf->wind, sf->downwash};
    for (int k = 0; k < 9; k++) memcpy(&hs[(size_t)k * nd], sarr[k], sizeof(double) * nd);
    memcpy(&hs[(size_t)9 * nd], sf->aterm, sizeof(double) * na);
    memcpy(&hs[(size_t)9 * nd + na], sf->adot, sizeof(double) * na);
    memcpy(&hs[(size_t)9 * nd + 2 * na], sf->aprev, sizeof(double) * na);
    std::vector<double> ht((size_t)(2 * (na + 1) * nd));
    for (int n = 0; n <= na; n++)
        for (int i = 0; i < nd; i++) {
            ht[(size_t)n * nd + i] = n == 0 ? 1.0 : std::cos(n * sf->theta[i]);
            ht[(size_t)(na + 1) * nd + (size_t)n * nd + i] = std::sin(n * sf->theta[i]);
        }
    const double vc_new = 0.02 * sf->c, vc4 = vc_new * vc_new * vc_new * vc_new;
    std::vector<double> hw((size_t)(7 * tot), 0.0);
    for (int64_t i = 0; i < tot; i++) { hw[3 * tot + i] = vc4; hw[4 * tot + i] = vc_new; }
    for (int64_t i = 0; i < nd - 1; i++) {
        hw[0 * tot + off_b + i] = sf->bv.x[i]; hw[1 * tot + off_b + i] = sf->bv.z[i]; hw[2 * tot + off_b + i] = sf->bv.s[i];
        hw[4 * tot + off_b + i] = sf->bv.vc[i];
        hw[3 * tot + off_b + i] = sf->bv.vc[i] * sf->bv.vc[i] * sf->bv.vc[i] * sf->bv.vc[i];
    }
    std::vector<StepState> hst((size_t)M);
    std::vector<LdvmDev> hd((size_t)M);
    StepState s0;
    memset(&s0, 0, sizeof(s0));
    s0.wake.begin[0] = 0; s0.wake.end[0] = 0;
    s0.wake.begin[1] = off_l; s0.wake.end[1] = off_l;
    s0.wake.begin[2] = off_e; s0.wake.end[2] = off_e;
    s0.wake.begin[3] = off_b; s0.wake.end[3] = off_b + nd - 1;
    s0.bnd.begin[0] = 0; s0.bnd.end[0] = nd;
    s0.levflag = sf->levflag;
    for (int k = 0; k < 6; k++) s0.kin[k] = sf->kinem[k];
    s0.a0 = sf->a0; s0.a0dot = sf->a0dot; s0.a0prev = sf->a0prev;
    for (int64_t m = 0; m < M; m++) {
        hst[m] = s0;
        LdvmDev &d = hd[m];
        memset(&d, 0, sizeof(d));
        d.st = st.p + m;
        d.ndiv = nd; d.naterm = na; d.driver = op->driver; d.solve_mode = op->solve_mode; d.del_kind = op->delvort_kind;
        d.del_limit = op->del_limit; d.del_dist = op->del_dist; d.del_tol = op->del_tol;
        d.c = sf->c; d.uref = sf->uref; d.pvt = sf->pvt; d.lespcrit = lespcrit ? lespcrit[m] : sf->lespcrit; d.dt = op->dt;
        d.vc_new = vc_new; d.vc4_new = vc4;
        double *sp = surf.p + m * nsurf;
        d.theta = sp; d.x = sp + nd; d.cam = sp + 2 * nd; d.cam_slope = sp + 3 * nd; d.bnd_x = sp + 4 * nd; d.bnd_z = sp + 5 * nd;
        d.uind = sp + 6 * nd; d.wind = sp + 7 * nd; d.downwash = sp + 8 * nd;
        d.aterm = sp + 9 * nd; d.adot = sp + 9 * nd + na; d.aprev = sp + 9 * nd + 2 * na;
        d.costab = tabs.p; d.sintab = tabs.p + (int64_t)(na + 1) * nd;
        double *wp = wake.p + m * 7 * tot;
        d.wx = wp; d.wz = wp + tot; d.wg = wp + 2 * tot; d.wvc4 = wp + 3 * tot; d.wvc = wp + 4 * tot; d.wvx = wp + 5 * tot; d.wvz = wp + 6 * tot;
        d.kin = kin.p + m * nsteps * 9; d.kin_rows = nsteps;
        d.mat = mat.p + m * nsteps * 8;
        d.p2u = p2.p + m * np2; d.p2w = p2.p + m * np2 + 4 * ndp; d.p2stride = ndp; d.S2 = 1;
    }
    UF_CUDA(cudaMemcpyAsync(st.p, hst.data(), sizeof(StepState) * M, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(devs.p, hd.data(), sizeof(LdvmDev) * M, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(tabs.p, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(kin.p, tables, sizeof(double) * M * nsteps * 9, cudaMemcpyHostToDevice, stream));
    // replicate the member image on the device: upload once, then D2D copies with doubling
    UF_CUDA(cudaMemcpyAsync(surf.p, hs.data(), sizeof(double) * nsurf, cudaMemcpyHostToDevice, stream));
    UF_CUDA(cudaMemcpyAsync(wake.p, hw.data(), sizeof(double) * 7 * tot, cudaMemcpyHostToDevice, stream));
    for (int64_t have = 1; have < M; have *= 2) {
        const int64_t cnt = std::min(have, M - have);
        UF_CUDA(cudaMemcpyAsync(surf.p + have * nsurf, surf.p, sizeof(double) * nsurf * cnt, cudaMemcpyDeviceToDevice, stream));
        UF_CUDA(cudaMemcpyAsync(wake.p + have * 7 * tot, wake.p, sizeof(double) * 7 * tot * cnt, cudaMemcpyDeviceToDevice, stream));
    }
    cudaEvent_t e0, e1;
    UF_CUDA(cudaEventCreate(&e0));
    UF_CUDA(cudaEventCreate(&e1));
    UF_CUDA(cudaEventRecord(e0, stream));
    k_ensemble<<<(unsigned)M, kSolveThreads, 0, stream>>>(devs.p, nsteps);
    UF_CUDA(cudaGetLastError());
    UF_CUDA(cudaEventRecord(e1, stream));
    std::vector<double> hm((size_t)(M * nsteps * 8));
    UF_CUDA(cudaMemcpyAsync(hm.data(), mat.p, sizeof(double) * hm.size(), cudaMemcpyDeviceToHost, stream));
    UF_CUDA(cudaMemcpyAsync(hst.data(), st.p, sizeof(StepState) * M, cudaMemcpyDeviceToHost, stream));
    UF_CUDA(cudaStreamSynchronize(stream));
    UF_CUDA(cudaEventElapsedTime(&g_batch_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (int64_t m = 0; m < M; m++) {
        double *mo = mat_out + m * nsteps * 8;
        const double *mi = &hm[(size_t)(m * nsteps * 8)];
        for (int64_t i = 0; i < nsteps; i++)
            for (int j = 0; j < 8; j++) mo[i + j * nsteps] = mi[8 * i + j];
        if (ntev_out) ntev_out[m] = hst[m].wake.end[0] - hst[m].wake.begin[0];
        if (nlev_out) nlev_out[m] = hst[m].wake.end[1] - hst[m].wake.begin[1];
    }
    return 0;
}

int unsflow_ldvm_batch_timing(float *ms_kernel) {
    UF_CHECK(ms_kernel != nullptr, "unsflow_ldvm_batch_timing: null argument");
    *ms_kernel = g_batch_ms;
    return 0;
}

}  // extern "C"
