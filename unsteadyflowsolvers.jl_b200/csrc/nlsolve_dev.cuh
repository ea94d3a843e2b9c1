// nlsolve_dev.cuh -- device-side restatement of NLsolve.jl's default solver as the reference calls it
// (nlsolve(not_in_place(f), x0), solvers.jl:199,216): trust region with Powell's dogleg, factor 1,
// autoscale, ftol = 1e-8 (inf-norm), central finite-difference Jacobian probing +eps then -eps with
// eps_j = cbrt(eps(Float64)) * max(1, |x_j|), Jacobian refreshed after every accepted step.
//
// Used only in UNSFLOW_SOLVE_NLSOLVE mode.  The Kelvin / Kelvin-Kutta residuals are affine in the shed
// strengths except for the sign switch of the LESP condition (typedefs.jl:340-344), so the functor is
// evaluated from its precomputed affine coefficients by ONE thread; what the iteration adds over a closed
// form is the PATH: with the sign switch the 2x2 system has two self-consistent roots (A0 = +LESPcrit and
// A0 = -LESPcrit) and which one NLsolve lands on depends on its iterates.  NLsolve.jl itself is an unpinned
// dependency of the reference; this follows its published algorithm.
#pragma once

namespace unsflow {

template <int N>
struct TrWork {
    double r[N], rn[N], J[N * N], d[N], p[N], rp[N];
};

template <int N>
__device__ inline void tr_solve_lin(const double *J, const double *rhs, double *p) {
    if (N == 1) { p[0] = rhs[0] / J[0]; return; }
    // 2x2 Gaussian elimination with partial pivoting (same operation order as a general LU)
    double a00 = J[0], a01 = J[1], b0 = rhs[0], a10 = J[2], a11 = J[3], b1 = rhs[1];
    if (fabs(a10) > fabs(a00)) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = b0; b0 = b1; b1 = t;
    }
    const double f = a10 / a00;
    a11 -= f * a01;
    b1 -= f * b0;
    p[1] = b1 / a11;
    p[0] = (b0 - a01 * p[1]) / a00;
}

template <int N>
__device__ inline double tr_wnorm(const double *d, const double *v) {
    double s = 0;
    for (int i = 0; i < N; i++) s += (d[i] * v[i]) * (d[i] * v[i]);
    return sqrt(s);
}

template <int N, class F>
__device__ inline void tr_fd_jacobian(F &fn, const double *x, double *J) {
    const double cbrt_eps = 6.0554544523933395e-06;
    double xt[N], fp[N], fm[N];
    for (int j = 0; j < N; j++) {
        const double e = cbrt_eps * fmax(1.0, fabs(x[j]));
        for (int k = 0; k < N; k++) xt[k] = x[k];
        xt[j] = x[j] + e; fn(xt, fp);
        xt[j] = x[j] - e; fn(xt, fm);
        for (int i = 0; i < N; i++) J[i * N + j] = (fp[i] - fm[i]) / (2 * e);
    }
}

template <int N>
__device__ inline void tr_dogleg(const double *r, const double *d, const double *J, double delta, double *p) {
    double pi_[N], rhs[N];
    for (int i = 0; i < N; i++) rhs[i] = -r[i];
    tr_solve_lin<N>(J, rhs, pi_);                       // Gauss-Newton step
    if (tr_wnorm<N>(d, pi_) <= delta) { for (int i = 0; i < N; i++) p[i] = pi_[i]; return; }
    double g[N], Jg[N], pc[N], df[N];
    for (int j = 0; j < N; j++) {
        g[j] = 0;
        for (int i = 0; i < N; i++) g[j] += J[i * N + j] * r[i];
        g[j] /= d[j] * d[j];
    }
    for (int i = 0; i < N; i++) { Jg[i] = 0; for (int j = 0; j < N; j++) Jg[i] += J[i * N + j] * g[j]; }
    const double gn = tr_wnorm<N>(d, g);
    double Jgn2 = 0;
    for (int i = 0; i < N; i++) Jgn2 += Jg[i] * Jg[i];
    for (int i = 0; i < N; i++) pc[i] = -(gn * gn / Jgn2) * g[i];   // Cauchy point
    const double pcn = tr_wnorm<N>(d, pc);
    if (pcn >= delta) { for (int i = 0; i < N; i++) p[i] = pc[i] * (delta / pcn); return; }
    for (int i = 0; i < N; i++) df[i] = pi_[i] - pc[i];
    double b = 0;
    for (int i = 0; i < N; i++) b += 2 * pc[i] * df[i] * d[i] * d[i];
    double a = tr_wnorm<N>(d, df);
    a = a * a;
    const double tau = (-b + sqrt(b * b - 4 * a * (pcn * pcn - delta * delta))) / (2 * a);
    for (int i = 0; i < N; i++) p[i] = pc[i] + tau * df[i];
}

// x: start on entry, soln.zero on exit.  The last functor evaluation (whose side effects the reference keeps)
// is the Jacobian probe at x - eps_N * e_N.
template <int N, class F>
__device__ inline void dev_nlsolve_tr(F &fn, double *x) {
    const double ftol = 1e-8, eta = 1e-4;
    TrWork<N> w;
    fn(x, w.r);
    tr_fd_jacobian<N>(fn, x, w.J);
    double fmaxv = 0;
    for (int i = 0; i < N; i++) fmaxv = fmax(fmaxv, fabs(w.r[i]));
    if (fmaxv <= ftol) return;
    for (int j = 0; j < N; j++) {
        double s = 0;
        for (int i = 0; i < N; i++) s += w.J[i * N + j] * w.J[i * N + j];
        w.d[j] = sqrt(s);
        if (w.d[j] == 0) w.d[j] = 1;
    }
    double delta = tr_wnorm<N>(w.d, x);
    if (delta == 0) delta = 1.0;
    for (int it = 0; it < 1000; it++) {
        tr_dogleg<N>(w.r, w.d, w.J, delta, w.p);
        for (int i = 0; i < N; i++) x[i] += w.p[i];
        fn(x, w.rn);
        double s_r = 0, s_rn = 0, s_rp = 0;
        for (int i = 0; i < N; i++) {
            w.rp[i] = w.r[i];
            for (int j = 0; j < N; j++) w.rp[i] += w.J[i * N + j] * w.p[j];
            s_r += w.r[i] * w.r[i]; s_rn += w.rn[i] * w.rn[i]; s_rp += w.rp[i] * w.rp[i];
        }
        const double rho = (s_r - s_rn) / (s_r - s_rp);
        bool converged = false;
        if (rho > eta) {
            for (int i = 0; i < N; i++) w.r[i] = w.rn[i];
            tr_fd_jacobian<N>(fn, x, w.J);
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int i = 0; i < N; i++) s += w.J[i * N + j] * w.J[i * N + j];
                w.d[j] = fmax(0.1 * w.d[j], sqrt(s));
            }
            fmaxv = 0;
            for (int i = 0; i < N; i++) fmaxv = fmax(fmaxv, fabs(w.r[i]));
            converged = fmaxv <= ftol;
        } else {
            for (int i = 0; i < N; i++) x[i] -= w.p[i];
        }
        if (rho < 0.1) delta = delta / 2;
        else if (rho >= 0.9) delta = 2 * tr_wnorm<N>(w.d, w.p);
        else if (rho >= 0.5) delta = fmax(delta, 2 * tr_wnorm<N>(w.d, w.p));
        if (converged) break;
    }
}

}  // namespace unsflow
