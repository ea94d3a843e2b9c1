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
// k_lve_solve (the influence matrix is a rank-1 / rank-2 update of a constant matrix whose inverse is
// computed once on the host: Sherman-Morrison / Woodbury with 2-4 warp-parallel mat-vecs per step,
// LEV ejection logic, panel pressures and loads), then the usual roll-up + finish.
// =============================================================================================
constexpr int kLveArrays = 15;
enum { LV_DS, LV_CSD, LV_NX, LV_NZ, LV_TX, LV_TZ, LV_VX, LV_VZ, LV_CX, LV_CZ, LV_PREV, LV_CHG, LV_CIRC, LV_REL, LV_A1 };
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

// y = Binv * v for the shared-memory copy of Binv (n x n, row-major): one warp per row, lanes strided
// over the columns, butterfly sum.
__device__ __forceinline__ void cta_matvec(const double *Binv, int n, const double *v, double *y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < n; r += nwarps) {
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) acc = fma(Binv[r * n + j], v[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[r] = acc;
    }
}

__device__ void lve_solve(const LdvmDev &d, double *A /* dynamic smem: nd*nd doubles for Binv */) {
    StepState *s = d.st;
    LveState *L = d.lve;
    const int tid = threadIdx.x, nd = d.ndiv, np = nd - 1;
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    __shared__ double uind[kNdMax], wind[kNdMax], relx[kNdMax], relz[kNdMax], circ[kNdMax], prev[kNdMax], sc[16];
    __shared__ double rhs[kNdMax], wcol[kNdMax], lcol[kNdMax], yv[kNdMax], zw[kNdMax], zl[kNdMax];
    __shared__ int lev_shed;
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
    // ---- a\RHS (LVE.jl:93) without refactorising every step -----------------------------------------
    // The influence matrix of influence_coeff (calcs.jl:756-787) is  A = B + (w - e_np) e_np^T : the
    // bound-vortex columns and the Kelvin row of ones never change (B, with the unit vector e_np as its
    // last column, is inverted ONCE on the host), only the wake column w moves.  Sherman-Morrison:
    //   x = y - z * y_np / (1 + z_np),   y = Binv RHS,   z = Binv w - Binv[:, np].
    double *Binv = A;                                   // dynamic shared memory: nd x nd
    for (int e = tid; e < nd * nd; e += blockDim.x) Binv[e] = d.lve_binv[e];
    auto fill_column = [&](double *col, double xs, double zs) {      // Vatistas-core source (calcs.jl:776-782)
        for (int i = tid; i < nd; i += blockDim.x) {
            if (i == np) { col[i] = 1.0; continue; }
            const double xd = xs - la[LV_CX * nd + i], zd = zs - la[LV_CZ * nd + i], dsq = xd * xd + zd * zd;
            const double den = kTwoPi * sqrt(d.vc4_new + dsq * dsq);
            col[i] = (-zd / den) * la[LV_NX * nd + i] + (xd / den) * la[LV_NZ * nd + i];
        }
    };
    auto fill_rhs = [&](bool lev, double gcu) {
        for (int i = tid; i < nd; i += blockDim.x) {
            double r;
            if (i == np) r = sc[2] - (lev ? gcu : 0.0);                                    // :78, :138
            else {
                r = -((relx[i] + uind[i]) * la[LV_NX * nd + i] + (relz[i] + wind[i]) * la[LV_NZ * nd + i]);   // :89
                if (lev) r -= gcu * la[LV_A1 * nd + i];                                    // :136 (a1 = first column of a)
            }
            rhs[i] = r;
        }
    };
    fill_column(wcol, x_w, z_w);
    fill_rhs(false, 0.0);
    __syncthreads();
    cta_matvec(Binv, nd, rhs, yv);
    cta_matvec(Binv, nd, wcol, zw);
    __syncthreads();
    for (int i = tid; i < nd; i += blockDim.x) zw[i] -= Binv[i * nd + np];
    __syncthreads();
    {
        const double fac = yv[np] / (1.0 + zw[np]);
        for (int i = tid; i < nd; i += blockDim.x) circ[i] = yv[i] - zw[i] * fac;           // circ = a\RHS :93
    }
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
        // mod_influence_coeff (calcs.jl:789-836) drops the first bound column, shifts left and appends the
        // LEV column l: a' = [b_1 .. b_{np-1}, w, l].  Moving l to the front, a'' = [l, b_1 .. b_{np-1}, w]
        //   = B + (l - b_0) e_0^T + (w - e_np) e_np^T  -- a rank-2 update of the same B (Woodbury):
        //   x'' = y - Z (I + V^T Z)^{-1} V^T y,  Z = [Binv l - e_0, z],  V = [e_0, e_np],
        // with x'' = [G_lev, G_1 .. G_{np-1}, G_wake]  (the reference's shifted circ, LVE.jl:142-145).
        fill_column(lcol, sc[4], sc[5]);
        fill_rhs(true, gcu);
        __syncthreads();
        cta_matvec(Binv, nd, rhs, yv);
        cta_matvec(Binv, nd, lcol, zl);
        __syncthreads();
        if (tid == 0) {
            zl[0] -= 1.0;
            const double m00 = 1.0 + zl[0], m01 = zw[0], m10 = zl[np], m11 = 1.0 + zw[np];
            const double det = m00 * m11 - m01 * m10;
            sc[8] = (yv[0] * m11 - m01 * yv[np]) / det;
            sc[9] = (m00 * yv[np] - yv[0] * m10) / det;
        }
        __syncthreads();
        {
            const double c0 = sc[8], c1 = sc[9];
            for (int i = tid; i < nd; i += blockDim.x) rhs[i] = yv[i] - zl[i] * c0 - zw[i] * c1;   // x''
        }
        __syncthreads();
        if (tid == 0) d.wg[s->wake.end[1] - 1] = rhs[0];                                    // LEV strength :142
        for (int i = tid; i < nd; i += blockDim.x) circ[i] = i == 0 ? gcu : rhs[i];         // :144-145
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
