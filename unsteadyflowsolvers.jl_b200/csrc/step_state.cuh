// step_state.cuh -- device-resident state of the LDVM / LVE time steppers (shared by all step kernels)
#pragma once
#include "../../include/unsflow.h"
#include "induce_sym.cuh"
#include "sweep.cuh"
#include "runtime.h"
#include <cstdlib>

namespace unsflow {

constexpr int kNdMax = 256;    // max ndiv
constexpr int kNaMax = 128;    // max naterm
constexpr int kSolveThreads = 256;
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
    const double *lve_binv;   // LVE: inverse of the constant part of the influence matrix [ndiv][ndiv]
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

}  // namespace unsflow
