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

// ---- the same iteration for a system whose size is only known at run time (n <= NMAX): the multi-surface driver's
// KelvinConditionMult (n = nsurf, solvers.jl:343) and KelvinKuttaMult (n = nsurf + nshed, :372).  All work arrays live
// in a caller-provided workspace (shared memory, trn_ws_doubles(NMAX) doubles) instead of the thread's stack. ---------
__host__ __device__ constexpr int trn_ws_doubles(int nmax) { return 2 * nmax * nmax + 14 * nmax; }

// Gaussian elimination with partial pivoting, row-by-row elimination, fixed operation order; A = n x (n + 1) scratch
__device__ inline void trn_solve_lin(int n, const double *J, const double *rhs, double *p, double *A) {
    const int w = n + 1;
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[i * w + j] = J[i * n + j]; A[i * w + n] = rhs[i]; }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(A[r * w + c]) > fabs(A[piv * w + c])) piv = r;
        if (piv != c) for (int j = 0; j <= n; j++) { const double t = A[c * w + j]; A[c * w + j] = A[piv * w + j]; A[piv * w + j] = t; }
        for (int r = c + 1; r < n; r++) {
            const double f = A[r * w + c] / A[c * w + c];
            for (int j = c; j <= n; j++) A[r * w + j] -= f * A[c * w + j];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double v = A[i * w + n];
        for (int j = i + 1; j < n; j++) v -= A[i * w + j] * p[j];
        p[i] = v / A[i * w + i];
    }
}

__device__ inline double trn_wnorm(int n, const double *d, const double *v) {
    double s = 0;
    for (int i = 0; i < n; i++) s += (d[i] * v[i]) * (d[i] * v[i]);
    return sqrt(s);
}

// t = 3 n doubles of scratch
template <class F>
__device__ inline void trn_fd_jacobian(F &fn, int n, const double *x, double *J, double *t) {
    const double cbrt_eps = 6.0554544523933395e-06;
    double *xt = t, *fp = t + n, *fm = t + 2 * n;
    for (int j = 0; j < n; j++) {
        const double e = cbrt_eps * fmax(1.0, fabs(x[j]));
        for (int k = 0; k < n; k++) xt[k] = x[k];
        xt[j] = x[j] + e; fn(xt, fp);
        xt[j] = x[j] - e; fn(xt, fm);
        for (int i = 0; i < n; i++) J[i * n + j] = (fp[i] - fm[i]) / (2 * e);
    }
}

// t = 6 n + n (n + 1) doubles of scratch
__device__ inline void trn_dogleg(int n, const double *r, const double *d, const double *J, double delta, double *p, double *t) {
    double *pi_ = t, *rhs = t + n, *g = t + 2 * n, *Jg = t + 3 * n, *pc = t + 4 * n, *df = t + 5 * n, *A = t + 6 * n;
    for (int i = 0; i < n; i++) rhs[i] = -r[i];
    trn_solve_lin(n, J, rhs, pi_, A);
    if (trn_wnorm(n, d, pi_) <= delta) { for (int i = 0; i < n; i++) p[i] = pi_[i]; return; }
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int i = 0; i < n; i++) s += J[i * n + j] * r[i];
        g[j] = s / (d[j] * d[j]);
    }
    for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += J[i * n + j] * g[j]; Jg[i] = s; }
    const double gn = trn_wnorm(n, d, g);
    double Jgn2 = 0;
    for (int i = 0; i < n; i++) Jgn2 += Jg[i] * Jg[i];
    for (int i = 0; i < n; i++) pc[i] = -(gn * gn / Jgn2) * g[i];
    const double pcn = trn_wnorm(n, d, pc);
    if (pcn >= delta) { for (int i = 0; i < n; i++) p[i] = pc[i] * (delta / pcn); return; }
    for (int i = 0; i < n; i++) df[i] = pi_[i] - pc[i];
    double b = 0;
    for (int i = 0; i < n; i++) b += 2 * pc[i] * df[i] * d[i] * d[i];
    double a = trn_wnorm(n, d, df);
    a = a * a;
    const double tau = (-b + sqrt(b * b - 4 * a * (pcn * pcn - delta * delta))) / (2 * a);
    for (int i = 0; i < n; i++) p[i] = pc[i] + tau * df[i];
}

// x: start on entry, soln.zero on exit; ws: trn_ws_doubles(nmax >= n) doubles
template <class F>
__device__ inline void dev_nlsolve_tr_n(F &fn, int n, double *x, double *ws) {
    const double ftol = 1e-8, eta = 1e-4;
    double *r = ws, *rn = ws + n, *d = ws + 2 * n, *p = ws + 3 * n, *rp = ws + 4 * n, *J = ws + 5 * n, *t = J + n * n;
    fn(x, r);
    trn_fd_jacobian(fn, n, x, J, t);
    double fmaxv = 0;
    for (int i = 0; i < n; i++) fmaxv = fmax(fmaxv, fabs(r[i]));
    if (fmaxv <= ftol) return;
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int i = 0; i < n; i++) s += J[i * n + j] * J[i * n + j];
        d[j] = sqrt(s);
        if (d[j] == 0) d[j] = 1;
    }
    double delta = trn_wnorm(n, d, x);
    if (delta == 0) delta = 1.0;
    for (int it = 0; it < 1000; it++) {
        trn_dogleg(n, r, d, J, delta, p, t);
        for (int i = 0; i < n; i++) x[i] += p[i];
        fn(x, rn);
        double s_r = 0, s_rn = 0, s_rp = 0;
        for (int i = 0; i < n; i++) {
            double v = r[i];
            for (int j = 0; j < n; j++) v += J[i * n + j] * p[j];
            rp[i] = v;
            s_r += r[i] * r[i]; s_rn += rn[i] * rn[i]; s_rp += rp[i] * rp[i];
        }
        const double rho = (s_r - s_rn) / (s_r - s_rp);
        bool converged = false;
        if (rho > eta) {
            for (int i = 0; i < n; i++) r[i] = rn[i];
            trn_fd_jacobian(fn, n, x, J, t);
            for (int j = 0; j < n; j++) {
                double s = 0;
                for (int i = 0; i < n; i++) s += J[i * n + j] * J[i * n + j];
                d[j] = fmax(0.1 * d[j], sqrt(s));
            }
            fmaxv = 0;
            for (int i = 0; i < n; i++) fmaxv = fmax(fmaxv, fabs(r[i]));
            converged = fmaxv <= ftol;
        } else {
            for (int i = 0; i < n; i++) x[i] -= p[i];
        }
        if (rho < 0.1) delta = delta / 2;
        else if (rho >= 0.9) delta = 2 * trn_wnorm(n, d, p);
        else if (rho >= 0.5) delta = fmax(delta, 2 * trn_wnorm(n, d, p));
        if (converged) break;
    }
}

}  // namespace unsflow
