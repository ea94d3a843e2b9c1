// step_lve.cuh -- LVE panel-solver step (k_lve_head, k_lve_solve bodies)
#pragma once
#include "step_single.cuh"

namespace unsflow {

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

}  // namespace unsflow
