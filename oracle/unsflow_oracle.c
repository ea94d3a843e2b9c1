/*
 * unsflow_oracle.c -- CPU restatement of the UNSflow discrete-vortex time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (unsteadyflowsolvers.jl_b200/,
 * libunsflow_cuda.so) may link, load or call this file.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py use it.
 *
 * PARITY UNPINNED: the reference is Julia and no julia binary exists in this image, and
 * the reference ships no golden vectors (its single test, test/ldvmLinTest.jl:33-36, only
 * bounds Cl_end to +-50 %).  This file restates the reference formulas in the reference's
 * own loop and summation order; it is pinned only against (i) that one test bound and
 * (ii) closed-form identities derived from the reference formulas (tests/test_oracle*.py).
 * NLsolve.jl (unpinned third-party dependency, Project.toml:13) is restated from its
 * published trust-region/dogleg algorithm in o_nlsolve_tr(); see DESIGN.md "H1".
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout, e.g. src/lowOrder2D/calcs.jl).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

#define O_PI 3.141592653589793238462643383279502884

/* TwoDVort, src/lowOrder2D/typedefs.jl:11-18 (same 6-double record) */
typedef struct { double x, z, s, vc, vx, vz; } o_vort;

typedef struct { o_vort *v; long n, cap; } o_vec;

static void vec_push(o_vec *a, o_vort q) {
    if (a->n == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 64;
        a->v = (o_vort *)realloc(a->v, (size_t)a->cap * sizeof(o_vort));
    }
    a->v[a->n++] = q;
}
static void vec_popfirst(o_vec *a) {
    memmove(a->v, a->v + 1, (size_t)(a->n - 1) * sizeof(o_vort));
    a->n--;
}

/* ------------------------------------------------------------------------------------------
 * Induction kernels
 * ---------------------------------------------------------------------------------------- */

/* ind_vel(src::Vector{TwoDVort}, t_x, t_z), src/lowOrder2D/calcs.jl:462-477.
 * Same expression shape: the denominator 2*pi*sqrt(vc^4 + d^4) is spelled out twice. */
static void ind_vel_aos(const o_vort *src, long ns, const double *tx, const double *tz, long nt,
                        double *uind, double *wind) {
    for (long itr = 0; itr < nt; itr++) {
        double u = 0.0, w = 0.0;
        for (long isr = 0; isr < ns; isr++) {
            double xdist = src[isr].x - tx[itr];
            double zdist = src[isr].z - tz[itr];
            double distsq = xdist * xdist + zdist * zdist;
            double vc = src[isr].vc;
            u = u - src[isr].s * zdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
            w = w + src[isr].s * xdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
        }
        uind[itr] = u;
        wind[itr] = w;
    }
}

/* mutual_ind(vorts), src/lowOrder2D/calcs.jl:491-510: symmetric i<j loop, accumulates into
 * the existing vx, vz (the caller zeroes them, calcs.jl:229-242). */
static void mutual_ind_aos(o_vort *v, long n) {
    for (long i = 0; i < n; i++) {
        for (long j = i + 1; j < n; j++) {
            double dx = v[i].x - v[j].x;
            double dz = v[i].z - v[j].z;
            double dsq = dx * dx + dz * dz;
            double magitr = 1. / (2 * O_PI * sqrt(v[j].vc * v[j].vc * v[j].vc * v[j].vc + dsq * dsq));
            double magjtr = 1. / (2 * O_PI * sqrt(v[i].vc * v[i].vc * v[i].vc * v[i].vc + dsq * dsq));
            v[j].vx -= dz * v[i].s * magjtr;
            v[j].vz += dx * v[i].s * magjtr;
            v[i].vx += dz * v[j].s * magitr;
            v[i].vz -= dx * v[j].s * magitr;
        }
    }
}

static o_vort *soa_to_aos(long n, const double *x, const double *z, const double *s, const double *vc,
                          const double *vx, const double *vz) {
    o_vort *v = (o_vort *)malloc((size_t)(n > 0 ? n : 1) * sizeof(o_vort));
    for (long i = 0; i < n; i++) {
        v[i].x = x[i]; v[i].z = z[i]; v[i].s = s[i]; v[i].vc = vc[i];
        v[i].vx = vx ? vx[i] : 0.0; v[i].vz = vz ? vz[i] : 0.0;
    }
    return v;
}

/* SoA-boundary wrapper of ind_vel (calcs.jl:462-477) */
void o_ind_vel(long ns, const double *sx, const double *sz, const double *ss, const double *svc,
               long nt, const double *tx, const double *tz, double *u, double *w) {
    o_vort *src = soa_to_aos(ns, sx, sz, ss, svc, NULL, NULL);
    ind_vel_aos(src, ns, tx, tz, nt, u, w);
    free(src);
}

/* SoA-boundary wrapper of mutual_ind (calcs.jl:491-510); vx, vz are accumulated into. */
void o_mutual_ind(long n, const double *x, const double *z, const double *s, const double *vc,
                  double *vx, double *vz) {
    o_vort *v = soa_to_aos(n, x, z, s, vc, vx, vz);
    mutual_ind_aos(v, n);
    for (long i = 0; i < n; i++) { vx[i] = v[i].vx; vz[i] = v[i].vz; }
    free(v);
}

/* "Truth" for the summation-error metric: same per-term formula as
 * calcs.jl:472-473 evaluated in long double with a long-double running sum, plus the
 * sum of |terms| that normalises the error.  Targets [t0,t1) only so large N is affordable.
 * Not a reference function: it is the yardstick the reordered GPU sum is measured with. */
typedef struct {
    long ns; const double *sx, *sz, *ss, *svc; long t0, t1; const double *tx, *tz;
    double *u, *w, *uabs, *wabs; const o_vort *aos; int kind;
} o_job;

static void *job_run(void *arg) {
    o_job *j = (o_job *)arg;
    if (j->kind == 1) {   /* reference-expression slice, calcs.jl:467-475 */
        ind_vel_aos(j->aos, j->ns, j->tx + j->t0, j->tz + j->t0, j->t1 - j->t0, j->u + j->t0, j->w + j->t0);
        return NULL;
    }
    for (long itr = j->t0; itr < j->t1; itr++) {
        long double su = 0, sw = 0, au = 0, aw = 0;
        for (long isr = 0; isr < j->ns; isr++) {
            long double xd = (long double)j->sx[isr] - j->tx[itr];
            long double zd = (long double)j->sz[isr] - j->tz[itr];
            long double d2 = xd * xd + zd * zd;
            long double vc = j->svc[isr];
            long double den = 2 * (long double)O_PI * sqrtl(vc * vc * vc * vc + d2 * d2);
            long double tu = -(long double)j->ss[isr] * zd / den;
            long double tw = (long double)j->ss[isr] * xd / den;
            su += tu; sw += tw; au += fabsl(tu); aw += fabsl(tw);
        }
        j->u[itr] = (double)su; j->w[itr] = (double)sw;
        if (j->uabs) j->uabs[itr] = (double)au;
        if (j->wabs) j->wabs[itr] = (double)aw;
    }
    return NULL;
}

int o_num_threads(void) {
    const char *e = getenv("UNSFLOW_ORACLE_THREADS");
    if (e && atoi(e) > 0) return atoi(e);
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void run_jobs(o_job proto, long nt) {
    int nth = o_num_threads();
    if (nth > nt) nth = (int)(nt > 0 ? nt : 1);
    pthread_t th[nth];
    o_job jobs[nth];
    for (int k = 0; k < nth; k++) {
        jobs[k] = proto;
        jobs[k].t0 = nt * k / nth;
        jobs[k].t1 = nt * (k + 1) / nth;
        pthread_create(&th[k], NULL, job_run, &jobs[k]);
    }
    for (int k = 0; k < nth; k++) pthread_join(th[k], NULL);
}

void o_ind_vel_truth(long ns, const double *sx, const double *sz, const double *ss, const double *svc,
                     long nt, const double *tx, const double *tz,
                     double *u, double *w, double *uabs, double *wabs) {
    o_job j = { ns, sx, sz, ss, svc, 0, 0, tx, tz, u, w, uabs, wabs, NULL, 0 };
    run_jobs(j, nt);
}

/* CPU-baseline variants (bench.py cpu_baseline / --impl reference).
 * o_mutual_ind_time: exactly the reference loop above on AoS records, returns seconds.
 * o_ind_vel_slice_mt: the same per-pair expression as calcs.jl:472-473, target-sliced over
 * pthreads (the reference itself is single-threaded; this is the generous bound). */
double o_now(void);
double o_mutual_ind_time(long n, const double *x, const double *z, const double *s, const double *vc,
                         double *vx, double *vz) {
    o_vort *v = soa_to_aos(n, x, z, s, vc, NULL, NULL);
    double t0 = o_now();
    mutual_ind_aos(v, n);
    double t1 = o_now();
    for (long i = 0; i < n; i++) { vx[i] = v[i].vx; vz[i] = v[i].vz; }
    free(v);
    return t1 - t0;
}

double o_ind_vel_slice_mt(long ns, const double *sx, const double *sz, const double *ss, const double *svc,
                          long nt, const double *tx, const double *tz, double *u, double *w) {
    o_vort *src = soa_to_aos(ns, sx, sz, ss, svc, NULL, NULL);
    double t0 = o_now();
    o_job j = { ns, sx, sz, ss, svc, 0, 0, tx, tz, u, w, NULL, NULL, src, 1 };
    run_jobs(j, nt);
    double t1 = o_now();
    free(src);
    return t1 - t0;
}

#include <time.h>
double o_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------
 * Surface / field state (TwoDSurf typedefs.jl:38-176, TwoDFlowField typedefs.jl:20-36)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int ndiv, naterm;
    double c, uref, pvt, lespcrit;
    int levflag;
    double alpha, h, alphadot, hdot, u, udot;          /* KinemPar, src/kinem.jl:3-10 */
    double a0, a0dot, a0prev;
    double *theta, *x, *cam, *cam_slope, *bnd_x, *bnd_z, *uind, *wind, *downwash; /* ndiv */
    double *aterm, *adot, *aprev;                      /* naterm */
    o_vort *bv;                                        /* ndiv-1 */
} o_surf;

typedef struct {
    double u, w;                                       /* curfield.u[1], curfield.w[1] */
    o_vec tev, lev, extv;
} o_field;

/* TwoDOFPar typedefs.jl:178-190 and KinemPar2DOF typedefs.jl:193-220, flat */
typedef struct {
    double x_alpha, r_alpha, kappa, w_alpha, w_h, w_alphadot, w_hdot,
           cubic_h_1, cubic_h_3, cubic_alpha_1, cubic_alpha_3;
} o_strpar;
typedef struct {
    double alpha, h, alphadot, hdot, u, udot, alphaddot, hddot, alpha_pr, h_pr,
           alphadot_pr, alphadot_pr2, alphadot_pr3, hdot_pr, hdot_pr2, hdot_pr3,
           alphaddot_pr, alphaddot_pr2, alphaddot_pr3, hddot_pr, hddot_pr2, hddot_pr3;
} o_kin2dof;

typedef struct {
    o_surf s;
    o_field f;
    o_strpar sp;
    o_kin2dof k2;
    double t;
    double cl, cd, cm;   /* carried between steps by ldvm2DOF, solvers.jl:692-693,765 */
    long nfev;           /* functor evaluations (diagnostic) */
    int lesp_force;      /* exact-root mode: LESP branch fixed to the sign of the A0 that triggered shedding (0 = per evaluation) */
} o_sim;

/* simpleTrapz, src/utils.jl:105-115 */
static double simple_trapz(const double *y, const double *x, int len) {
    double r = 0.0;
    for (int i = 1; i < len; i++) r += (x[i] - x[i - 1]) * (y[i] + y[i - 1]);
    return r / 2.0;
}

/* update_downwash, src/lowOrder2D/calcs.jl:49-54 (expression order kept) */
static void update_downwash(o_surf *s, double v1, double v2) {
    for (int ib = 0; ib < s->ndiv; ib++) {
        s->downwash[ib] = -(s->u + v1) * sin(s->alpha) - s->uind[ib] * sin(s->alpha)
            + (s->hdot - v2) * cos(s->alpha) - s->wind[ib] * cos(s->alpha)
            - s->alphadot * (s->x[ib] - s->pvt * s->c)
            + s->cam_slope[ib] * (s->uind[ib] * cos(s->alpha) + (s->u + v1) * cos(s->alpha)
                                  + (s->hdot - v2) * sin(s->alpha) - s->wind[ib] * sin(s->alpha));
    }
}

/* helper: simpleTrapz(downwash .* cos.(ia*theta), theta) as in calcs.jl:4,16,59,68 */
static double trapz_dw_cos(const o_surf *s, int ia, double *tmp) {
    for (int ib = 0; ib < s->ndiv; ib++) tmp[ib] = s->downwash[ib] * cos(ia * s->theta[ib]);
    return simple_trapz(tmp, s->theta, s->ndiv);
}

/* update_a0anda1, src/lowOrder2D/calcs.jl:57-63 */
static void update_a0anda1(o_surf *s) {
    double tmp[s->ndiv];
    s->a0 = simple_trapz(s->downwash, s->theta, s->ndiv);
    for (int ib = 0; ib < s->ndiv; ib++) tmp[ib] = s->downwash[ib] * cos(s->theta[ib]);
    s->aterm[0] = simple_trapz(tmp, s->theta, s->ndiv);
    s->a0 = -s->a0 / (s->uref * O_PI);
    s->aterm[0] = 2. * s->aterm[0] / (s->uref * O_PI);
}

/* update_a2toan, src/lowOrder2D/calcs.jl:66-72 */
static void update_a2toan(o_surf *s) {
    double tmp[s->ndiv];
    for (int ia = 2; ia <= s->naterm; ia++) {
        s->aterm[ia - 1] = trapz_dw_cos(s, ia, tmp);
        s->aterm[ia - 1] = 2. * s->aterm[ia - 1] / (s->uref * O_PI);
    }
}

/* update_a2a3adot, src/lowOrder2D/calcs.jl:2-12 */
static void update_a2a3adot(o_surf *s, double dt) {
    double tmp[s->ndiv];
    for (int ia = 2; ia <= 3; ia++) {
        s->aterm[ia - 1] = trapz_dw_cos(s, ia, tmp);
        s->aterm[ia - 1] = 2. * s->aterm[ia - 1] / (s->uref * O_PI);
    }
    s->a0dot = (s->a0 - s->a0prev) / dt;
    for (int ia = 1; ia <= 3; ia++) s->adot[ia - 1] = (s->aterm[ia - 1] - s->aprev[ia - 1]) / dt;
}

/* update_atermdot, src/lowOrder2D/calcs.jl:14-24 */
static void update_atermdot(o_surf *s, double dt) {
    double tmp[s->ndiv];
    for (int ia = 2; ia <= s->naterm; ia++) {
        s->aterm[ia - 1] = trapz_dw_cos(s, ia, tmp);
        s->aterm[ia - 1] = 2. * s->aterm[ia - 1] / (s->uref * O_PI);
    }
    s->a0dot = (s->a0 - s->a0prev) / dt;
    for (int ia = 1; ia <= s->naterm; ia++) s->adot[ia - 1] = (s->aterm[ia - 1] - s->aprev[ia - 1]) / dt;
}

/* update_indbound, src/lowOrder2D/calcs.jl:35-38: sources [tev; lev; extv] in that order */
static void update_indbound(o_surf *s, const o_field *f) {
    long n = f->tev.n + f->lev.n + f->extv.n;
    o_vort *all = (o_vort *)malloc((size_t)(n > 0 ? n : 1) * sizeof(o_vort));
    memcpy(all, f->tev.v, (size_t)f->tev.n * sizeof(o_vort));
    memcpy(all + f->tev.n, f->lev.v, (size_t)f->lev.n * sizeof(o_vort));
    memcpy(all + f->tev.n + f->lev.n, f->extv.v, (size_t)f->extv.n * sizeof(o_vort));
    ind_vel_aos(all, n, s->bnd_x, s->bnd_z, s->ndiv, s->uind, s->wind);
    free(all);
}

/* update_bv, src/lowOrder2D/calcs.jl:204-219 */
static void update_bv(o_surf *s) {
    double gamma[s->ndiv];
    for (int ib = 0; ib < s->ndiv; ib++) {
        gamma[ib] = (s->a0 * (1 + cos(s->theta[ib])));
        for (int ia = 1; ia <= s->naterm; ia++)
            gamma[ib] = gamma[ib] + s->aterm[ia - 1] * sin(ia * s->theta[ib]) * sin(s->theta[ib]);
        gamma[ib] = gamma[ib] * s->uref * s->c;
    }
    for (int ib = 1; ib < s->ndiv; ib++) {
        s->bv[ib - 1].s = (gamma[ib] + gamma[ib - 1]) * (s->theta[1] - s->theta[0]) / 2.;
        s->bv[ib - 1].x = (s->bnd_x[ib] + s->bnd_x[ib - 1]) / 2.;
        s->bv[ib - 1].z = (s->bnd_z[ib] + s->bnd_z[ib - 1]) / 2.;
    }
}

/* update_boundpos, src/lowOrder2D/calcs.jl:453-459 */
static void update_boundpos(o_surf *s, double dt) {
    for (int i = 0; i < s->ndiv; i++) {
        s->bnd_x[i] = s->bnd_x[i] + dt * ((s->pvt * s->c - s->x[i]) * sin(s->alpha) * s->alphadot - s->u
                                          + s->cam[i] * cos(s->alpha) * s->alphadot);
        s->bnd_z[i] = s->bnd_z[i] + dt * (s->hdot + (s->pvt * s->c - s->x[i]) * cos(s->alpha) * s->alphadot
                                          - s->cam[i] * sin(s->alpha) * s->alphadot);
    }
}

/* place_tev, src/lowOrder2D/calcs.jl:375-387 */
static void place_tev(o_surf *s, o_field *f, double dt) {
    long ntev = f->tev.n;
    double xloc, zloc;
    int nd = s->ndiv;
    if (ntev == 0) {
        xloc = s->bnd_x[nd - 1] + 0.5 * s->u * dt;
        zloc = s->bnd_z[nd - 1];
    } else {
        xloc = s->bnd_x[nd - 1] + (1. / 3.) * (f->tev.v[ntev - 1].x - s->bnd_x[nd - 1]);
        zloc = s->bnd_z[nd - 1] + (1. / 3.) * (f->tev.v[ntev - 1].z - s->bnd_z[nd - 1]);
    }
    o_vort q = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
    vec_push(&f->tev, q);
}

/* LE velocity + location shared by place_lev (calcs.jl:409-426) and ldvmLin (solvers.jl:561-570) */
static void lev_location(const o_surf *s, const o_field *f, double dt, double *xloc, double *zloc) {
    long nlev = f->lev.n;
    double le_vel_x = s->u - s->alphadot * sin(s->alpha) * s->pvt * s->c + s->uind[0];
    double le_vel_z = -s->alphadot * cos(s->alpha) * s->pvt * s->c - s->hdot + s->wind[0];
    if (s->levflag == 0) {
        *xloc = s->bnd_x[0] + 0.5 * le_vel_x * dt;
        *zloc = s->bnd_z[0] + 0.5 * le_vel_z * dt;
    } else {
        *xloc = s->bnd_x[0] + (1. / 3.) * (f->lev.v[nlev - 1].x - s->bnd_x[0]);
        *zloc = s->bnd_z[0] + (1. / 3.) * (f->lev.v[nlev - 1].z - s->bnd_z[0]);
    }
}

/* place_lev, src/lowOrder2D/calcs.jl:409-426 */
static void place_lev(o_surf *s, o_field *f, double dt) {
    double xloc, zloc;
    lev_location(s, f, dt, &xloc, &zloc);
    o_vort q = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
    vec_push(&f->lev, q);
}

/* wakeroll(surf::TwoDSurf, curfield, dt), src/lowOrder2D/calcs.jl:222-294 */
static void wakeroll(o_surf *s, o_field *f, double dt) {
    long ntev = f->tev.n, nlev = f->lev.n, nextv = f->extv.n, n = ntev + nlev + nextv;
    o_vort *all = (o_vort *)malloc((size_t)(n > 0 ? n : 1) * sizeof(o_vort));
    memcpy(all, f->tev.v, (size_t)ntev * sizeof(o_vort));
    memcpy(all + ntev, f->lev.v, (size_t)nlev * sizeof(o_vort));
    memcpy(all + ntev + nlev, f->extv.v, (size_t)nextv * sizeof(o_vort));
    for (long i = 0; i < n; i++) { all[i].vx = 0; all[i].vz = 0; }      /* :229-242 */
    mutual_ind_aos(all, n);                                               /* :245 */
    double *tx = (double *)malloc((size_t)(n > 0 ? n : 1) * 4 * sizeof(double));
    double *tz = tx + n, *ut = tz + n, *wt = ut + n;
    for (long i = 0; i < n; i++) { tx[i] = all[i].x; tz[i] = all[i].z; }
    ind_vel_aos(s->bv, s->ndiv - 1, tx, tz, n, ut, wt);                   /* :250 */
    for (long i = 0; i < n; i++) { all[i].vx += ut[i]; all[i].vz += wt[i]; }  /* :252-263 */
    for (long i = 0; i < n; i++) { all[i].vx += f->u; all[i].vz += f->w; }    /* :266-277 */
    for (long i = 0; i < n; i++) { all[i].x += dt * all[i].vx; all[i].z += dt * all[i].vz; } /* :280-291 */
    memcpy(f->tev.v, all, (size_t)ntev * sizeof(o_vort));
    memcpy(f->lev.v, all + ntev, (size_t)nlev * sizeof(o_vort));
    memcpy(f->extv.v, all + ntev + nlev, (size_t)nextv * sizeof(o_vort));
    free(tx);
    free(all);
}

/* one family of controlVortCount's Spalart merge, src/lowOrder2D/calcs.jl:550-572 (TEV) and
 * :577-598 (LEV).  Julia's tev[i], i=2:20 indexes out of bounds if fewer than 20 exist; the
 * reference would throw BoundsError there, we stop at the end of the vector instead. */
static void spalart_family(o_vec *a, double D0, double V0, double lx, double lz) {
    o_vort *v = a->v;
    double gamma_j = v[0].s;
    double d_j = sqrt((v[0].x - lx) * (v[0].x - lx) + (v[0].z - lz) * (v[0].z - lz));
    double z_j = sqrt(v[0].x * v[0].x + v[0].z * v[0].z);
    for (long i = 1; i < 20 && i < a->n; i++) {
        double gamma_k = v[i].s;
        double d_k = sqrt((v[i].x - lx) * (v[i].x - lx) + (v[i].z - lz) * (v[i].z - lz));
        double z_k = sqrt(v[i].x * v[i].x + v[i].z * v[i].z);
        double fact = fabs(gamma_j * gamma_k) * fabs(z_j - z_k)
                      / (fabs(gamma_j + gamma_k) * pow(D0 + d_j, 1.5) * pow(D0 + d_k, 1.5));
        if (fact < V0) {
            v[i].x = (fabs(gamma_j) * v[0].x + fabs(gamma_k) * v[i].x) / (fabs(gamma_j + gamma_k));
            v[i].z = (fabs(gamma_j) * v[0].z + fabs(gamma_k) * v[i].z) / (fabs(gamma_j + gamma_k));
            v[i].s += v[0].s;
            vec_popfirst(a);
            break;
        }
    }
}

/* controlVortCount, src/lowOrder2D/calcs.jl:541-602; delvort kind 0 = delNone, 1 = delSpalart
 * (src/delVort.jl:1-9).  The LEV block is nested inside the TEV-count test (:574). */
static void control_vort_count(int kind, long limit, double dist, double tol,
                               double lx, double lz, o_field *f) {
    if (kind != 1) return;
    if (f->tev.n > limit) {
        spalart_family(&f->tev, dist, tol, lx, lz);
        if (f->lev.n > limit) spalart_family(&f->lev, dist, tol, lx, lz);
    }
}

/* calc_forces, src/lowOrder2D/postprocess.jl:1-32 */
static void calc_forces(const o_surf *s, double v1, double v2, double *cl, double *cd, double *cm, double *cn_out) {
    double cnc = 2 * O_PI * ((s->u + v1) * cos(s->alpha) / s->uref + (s->hdot - v2) * sin(s->alpha) / s->uref)
                 * (s->a0 + s->aterm[0] / 2.);
    double cnnc = 2 * O_PI * (3 * s->c * s->a0dot / (4 * s->uref) + s->c * s->adot[0] / (4 * s->uref)
                              + s->c * s->adot[1] / (8 * s->uref));
    double cs = 2 * O_PI * s->a0 * s->a0;
    double nonl = 0, nonl_m = 0;
    for (int ib = 0; ib < s->ndiv - 1; ib++) {
        nonl = nonl + (s->uind[ib] * cos(s->alpha) - s->wind[ib] * sin(s->alpha)) * s->bv[ib].s;
        nonl_m = nonl_m + (s->uind[ib] * cos(s->alpha) - s->wind[ib] * sin(s->alpha)) * s->x[ib] * s->bv[ib].s;
    }
    nonl = nonl * 2. / (s->uref * s->uref * s->c);
    nonl_m = nonl_m * 2. / (s->uref * s->uref * s->c * s->c);
    double cn = cnc + cnnc + nonl;
    *cl = cn * cos(s->alpha) + cs * sin(s->alpha);
    *cd = cn * sin(s->alpha) - cs * cos(s->alpha);
    *cm = cn * s->pvt
          - 2 * O_PI * (((s->u + v1) * cos(s->alpha) / s->uref + (s->hdot - v2) * sin(s->alpha) / s->uref)
                            * (s->a0 / 4. + s->aterm[0] / 4. - s->aterm[1] / 8.)
                        + (s->c / s->uref) * (7. * s->a0dot / 16. + 3. * s->adot[0] / 16. + s->adot[1] / 16.
                                              - s->adot[2] / 64.))
          - nonl_m;
    if (cn_out) *cn_out = cn;
}

/* update_kinem2DOF, src/lowOrder2D/calcs.jl:604-650 */
static void update_kinem2dof(o_surf *s, const o_strpar *p, o_kin2dof *k, double dt, double cl, double cm) {
    k->alpha_pr = k->alpha;
    k->h_pr = k->h;
    k->alphadot_pr3 = k->alphadot_pr2; k->alphadot_pr2 = k->alphadot_pr; k->alphadot_pr = k->alphadot;
    k->hdot_pr3 = k->hdot_pr2; k->hdot_pr2 = k->hdot_pr; k->hdot_pr = k->hdot;
    k->alphaddot_pr3 = k->alphaddot_pr2; k->alphaddot_pr2 = k->alphaddot_pr;
    k->hddot_pr3 = k->hddot_pr2; k->hddot_pr2 = k->hddot_pr;

    double m11 = 2. / s->c;
    double m12 = -p->x_alpha * cos(k->alpha);
    double m21 = -2. * p->x_alpha * cos(k->alpha) / s->c;
    double m22 = p->r_alpha * p->r_alpha;
    double R1 = 4 * p->kappa * s->uref * s->uref * cl / (O_PI * s->c * s->c)
                - 2 * p->w_h * p->w_h * (p->cubic_h_1 * k->h + p->cubic_h_3 * (k->h * k->h * k->h)) / s->c
                - p->x_alpha * sin(k->alpha) * k->alphadot * k->alphadot;
    double R2 = 8 * p->kappa * s->uref * s->uref * cm / (O_PI * s->c * s->c)
                - p->w_alpha * p->w_alpha * p->r_alpha * p->r_alpha
                      * (p->cubic_alpha_1 * k->alpha + p->cubic_alpha_3 * (k->alpha * k->alpha * k->alpha));
    k->hddot_pr = (1 / (m11 * m22 - m21 * m12)) * (m22 * R1 - m12 * R2);
    k->alphaddot_pr = (1 / (m11 * m22 - m21 * m12)) * (-m21 * R1 + m11 * R2);

    k->alphadot = k->alphadot_pr + (dt / 12.) * (23 * k->alphaddot_pr - 16 * k->alphaddot_pr2 + 5 * k->alphaddot_pr3);
    k->hdot = k->hdot_pr + (dt / 12) * (23 * k->hddot_pr - 16 * k->hddot_pr2 + 5 * k->hddot_pr3);
    k->alpha = k->alpha_pr + (dt / 12) * (23 * k->alphadot_pr - 16 * k->alphadot_pr2 + 5 * k->alphadot_pr3);
    k->h = k->h_pr + (dt / 12) * (23 * k->hdot_pr - 16 * k->hdot_pr2 + 5 * k->hdot_pr3);

    s->alphadot = k->alphadot; s->hdot = k->hdot; s->alpha = k->alpha; s->h = k->h;
}

/* ------------------------------------------------------------------------------------------
 * Kelvin / Kelvin-Kutta functors (mutating, exactly like the reference)
 * ---------------------------------------------------------------------------------------- */

/* (kelv::KelvinCondition)(tev_iter), src/lowOrder2D/typedefs.jl:228-254 */
static double kelvin_functor(o_sim *m, double tev_iter) {
    o_surf *s = &m->s; o_field *f = &m->f;
    int nd = s->ndiv;
    double uprev[nd], wprev[nd], unow[nd], wnow[nd];
    o_vort *last = &f->tev.v[f->tev.n - 1];
    ind_vel_aos(last, 1, s->bnd_x, s->bnd_z, nd, uprev, wprev);          /* :233 */
    last->s = tev_iter;                                                   /* :236 */
    ind_vel_aos(last, 1, s->bnd_x, s->bnd_z, nd, unow, wnow);            /* :238 */
    for (int i = 0; i < nd; i++) {                                        /* :240-241 */
        s->uind[i] = s->uind[i] - uprev[i] + unow[i];
        s->wind[i] = s->wind[i] - wprev[i] + wnow[i];
    }
    update_downwash(s, f->u, f->w);                                       /* :244 */
    update_a0anda1(s);                                                    /* :247 */
    m->nfev++;
    return s->uref * s->c * O_PI * (s->a0 + s->aterm[0] / 2.)
           - s->uref * s->c * O_PI * (s->a0prev + s->aprev[0] / 2.) + last->s;   /* :249-251 */
}

/* (kelv::KelvinKutta)(v_iter), src/lowOrder2D/typedefs.jl:313-348 */
static void kelvinkutta_functor(o_sim *m, const double v_iter[2], double val[2]) {
    o_surf *s = &m->s; o_field *f = &m->f;
    int nd = s->ndiv;
    double uprev[nd], wprev[nd], unow[nd], wnow[nd];
    o_vort two[2];
    o_vort *lt = &f->tev.v[f->tev.n - 1], *ll = &f->lev.v[f->lev.n - 1];
    two[0] = *lt; two[1] = *ll;
    ind_vel_aos(two, 2, s->bnd_x, s->bnd_z, nd, uprev, wprev);           /* :319 */
    lt->s = v_iter[0];                                                    /* :322-323 */
    ll->s = v_iter[1];
    two[0] = *lt; two[1] = *ll;
    ind_vel_aos(two, 2, s->bnd_x, s->bnd_z, nd, unow, wnow);             /* :325 */
    for (int i = 0; i < nd; i++) {                                        /* :327-328 */
        s->uind[i] = s->uind[i] - uprev[i] + unow[i];
        s->wind[i] = s->wind[i] - wprev[i] + wnow[i];
    }
    update_downwash(s, f->u, f->w);                                       /* :331 */
    update_a0anda1(s);                                                    /* :334 */
    m->nfev++;
    val[0] = s->uref * s->c * O_PI * (s->a0 + s->aterm[0] / 2.)
             - s->uref * s->c * O_PI * (s->a0prev + s->aprev[0] / 2.) + lt->s + ll->s;  /* :336-338 */
    double lesp_cond = (s->a0 > 0) ? s->lespcrit : -s->lespcrit;          /* :340-344 */
    if (m->lesp_force) lesp_cond = m->lesp_force * s->lespcrit;          /* exact-root mode, see root_solve */
    val[1] = s->a0 - lesp_cond;                                           /* :345 */
}

/* ------------------------------------------------------------------------------------------
 * Root solve at the nlsolve call sites (solvers.jl:91,199,216,718,737).
 *
 * NLsolve.jl is NOT in /root/reference (unpinned dependency, Project.toml:13).  Two modes:
 *
 *  mode 0 "exact":  the residuals are affine in the unknown strengths, so we
 *      build the exact Jacobian by unit differences of the functor, take a Newton step,
 *      repeat until the step no longer changes the iterate, and finish with ONE functor
 *      evaluation at the root so the functor's side effects (uind, wind, downwash, a0, a1)
 *      correspond to the returned strengths.  The Kelvin-Kutta system has TWO self-consistent roots
 *      (A0 = +LESPcrit and A0 = -LESPcrit, typedefs.jl:340-344 switches the target on the sign of A0 at
 *      every evaluation); "exact" means the root on the branch of the A0 that triggered the shedding
 *      (the drivers pin the sign through lesp_force for the duration of the solve).  Which root NLsolve
 *      reaches depends on its iterates: that is mode 1.
 *
 *  mode 1 "nlsolve": restatement of NLsolve 4.x defaults as published -- method
 *      :trust_region (Powell dogleg, Nocedal & Wright alg. 4.1/11.5), factor = 1.0,
 *      autoscale = true, ftol = 1e-8 (inf-norm), xtol = 0, central finite-difference
 *      Jacobian with eps_j = cbrt(eps(Float64))*max(1,|x_j|) probing +eps then -eps,
 *      Jacobian re-evaluated after every accepted step BEFORE the convergence test.  The
 *      consequence that matters (H1): the LAST functor evaluation is the Jacobian probe at
 *      x - eps*e_last, and the reference drivers then use that stale surf state after only
 *      overwriting the strengths with soln.zero (solvers.jl:200,217).
 * ---------------------------------------------------------------------------------------- */
#define O_NMAX 8
typedef void (*o_fun)(void *, const double *, double *);

static void fun1(void *m, const double *x, double *r) { r[0] = kelvin_functor((o_sim *)m, x[0]); }
static void fun2(void *m, const double *x, double *r) { kelvinkutta_functor((o_sim *)m, x, r); }

static void solve_lin(int n, const double *J /*row-major*/, const double *rhs, double *p) {
    /* Gaussian elimination with partial pivoting, n <= O_NMAX */
    double A[O_NMAX][O_NMAX + 1];
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[i][j] = J[i * n + j]; A[i][n] = rhs[i]; }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j <= n; j++) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        for (int r = c + 1; r < n; r++) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j <= n; j++) A[r][j] -= f * A[c][j];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double v = A[i][n];
        for (int j = i + 1; j < n; j++) v -= A[i][j] * p[j];
        p[i] = v / A[i][i];
    }
}

static void solve_exact(void *m, o_fun fn, int n, double *x) {
    /* probe step 1e-3: small enough not to trip the sign switch of the LESP condition
     * (typedefs.jl:340), large enough that the difference quotient of an affine function is
     * exact to ~1e-13; the loop then polishes the root to rounding. */
    double r0[O_NMAX], r1[O_NMAX], J[O_NMAX * O_NMAX], p[O_NMAX], xt[O_NMAX];
    for (int it = 0; it < 12; it++) {
        fn(m, x, r0);
        for (int j = 0; j < n; j++) {
            const double hstep = 1e-3 * fmax(1.0, fabs(x[j]));
            for (int k = 0; k < n; k++) xt[k] = x[k];
            xt[j] = x[j] + hstep;
            fn(m, xt, r1);
            for (int i = 0; i < n; i++) J[i * n + j] = (r1[i] - r0[i]) / hstep;
        }
        double rhs[O_NMAX];
        for (int i = 0; i < n; i++) rhs[i] = -r0[i];
        solve_lin(n, J, rhs, p);
        double chg = 0;
        for (int k = 0; k < n; k++) { chg += fabs(p[k]) / (fabs(x[k]) + 1e-300); x[k] += p[k]; }
        if (chg < 1e-15) break;
    }
    fn(m, x, r0);   /* leave the functor's side effects at the root */
}

static void fd_jacobian(void *m, o_fun fn, int n, const double *x, double *J) {
    /* FiniteDiff central differences: column j uses x+eps*e_j first, then x-eps*e_j */
    const double cbrt_eps = cbrt(2.220446049250313e-16);
    double xt[O_NMAX], fp[O_NMAX], fm[O_NMAX];
    for (int j = 0; j < n; j++) {
        double e = cbrt_eps * fmax(1.0, fabs(x[j]));
        for (int k = 0; k < n; k++) xt[k] = x[k];
        xt[j] = x[j] + e; fn(m, xt, fp);
        xt[j] = x[j] - e; fn(m, xt, fm);
        for (int i = 0; i < n; i++) J[i * n + j] = (fp[i] - fm[i]) / (2 * e);
    }
}

static double wnorm(int n, const double *d, const double *v) {
    double s = 0;
    for (int i = 0; i < n; i++) s += (d[i] * v[i]) * (d[i] * v[i]);
    return sqrt(s);
}

/* Powell dogleg step inside NLsolve's trust_region (scaled by d) */
static void dogleg(int n, const double *r, const double *d, const double *J, double delta, double *p) {
    double pi_[O_NMAX], rhs[O_NMAX];
    for (int i = 0; i < n; i++) rhs[i] = -r[i];
    solve_lin(n, J, rhs, pi_);                      /* Gauss-Newton step */
    if (wnorm(n, d, pi_) <= delta) { for (int i = 0; i < n; i++) p[i] = pi_[i]; return; }
    double g[O_NMAX] = { 0 };                       /* g = J' r  scaled by 1/d^2 */
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) g[j] += J[i * n + j] * r[i];
        g[j] /= d[j] * d[j];
    }
    double Jg[O_NMAX] = { 0 };
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Jg[i] += J[i * n + j] * g[j];
    double gn = wnorm(n, d, g);
    double Jgn2 = 0; for (int i = 0; i < n; i++) Jgn2 += Jg[i] * Jg[i];
    double pc[O_NMAX];
    for (int i = 0; i < n; i++) pc[i] = -(gn * gn / Jgn2) * g[i];   /* Cauchy point */
    double pcn = wnorm(n, d, pc);
    if (pcn >= delta) { for (int i = 0; i < n; i++) p[i] = pc[i] * (delta / pcn); return; }
    double df[O_NMAX]; for (int i = 0; i < n; i++) df[i] = pi_[i] - pc[i];
    double b = 0; for (int i = 0; i < n; i++) b += 2 * pc[i] * df[i] * d[i] * d[i];
    double a = wnorm(n, d, df); a = a * a;
    double tau = (-b + sqrt(b * b - 4 * a * (pcn * pcn - delta * delta))) / (2 * a);
    for (int i = 0; i < n; i++) p[i] = pc[i] + tau * df[i];
}

static void o_nlsolve_tr(void *m, o_fun fn, int n, double *x) {
    const double ftol = 1e-8, eta = 1e-4;
    double r[O_NMAX], rn[O_NMAX], J[O_NMAX * O_NMAX], d[O_NMAX], p[O_NMAX], rp[O_NMAX];
    fn(m, x, r);                       /* value_jacobian!!: value first, then FD probes */
    fd_jacobian(m, fn, n, x, J);
    double fmaxv = 0; for (int i = 0; i < n; i++) fmaxv = fmax(fmaxv, fabs(r[i]));
    if (fmaxv <= ftol) return;
    for (int j = 0; j < n; j++) {
        double s = 0; for (int i = 0; i < n; i++) s += J[i * n + j] * J[i * n + j];
        d[j] = sqrt(s); if (d[j] == 0) d[j] = 1;
    }
    double delta = 1.0 * wnorm(n, d, x);
    if (delta == 0) delta = 1.0;
    for (int it = 0; it < 1000; it++) {
        dogleg(n, r, d, J, delta, p);
        for (int i = 0; i < n; i++) x[i] += p[i];
        fn(m, x, rn);
        double s_r = 0, s_rn = 0, s_rp = 0;
        for (int i = 0; i < n; i++) {
            rp[i] = r[i]; for (int j = 0; j < n; j++) rp[i] += J[i * n + j] * p[j];
            s_r += r[i] * r[i]; s_rn += rn[i] * rn[i]; s_rp += rp[i] * rp[i];
        }
        double rho = (s_r - s_rn) / (s_r - s_rp);
        int converged = 0;
        if (rho > eta) {
            for (int i = 0; i < n; i++) r[i] = rn[i];
            fd_jacobian(m, fn, n, x, J);          /* <- last functor call is x - eps*e_n */
            for (int j = 0; j < n; j++) {
                double s = 0; for (int i = 0; i < n; i++) s += J[i * n + j] * J[i * n + j];
                d[j] = fmax(0.1 * d[j], sqrt(s));
            }
            fmaxv = 0; for (int i = 0; i < n; i++) fmaxv = fmax(fmaxv, fabs(r[i]));
            converged = (fmaxv <= ftol);
        } else {
            for (int i = 0; i < n; i++) x[i] -= p[i];
        }
        if (rho < 0.1) delta = delta / 2;
        else if (rho >= 0.9) delta = 2 * wnorm(n, d, p);
        else if (rho >= 0.5) delta = fmax(delta, 2 * wnorm(n, d, p));
        if (converged) break;
    }
}

static void root_solve(void *m, int mode, o_fun fn, int n, double *x) {
    if (mode == 1) o_nlsolve_tr(m, fn, n, x);
    else solve_exact(m, fn, n, x);
}

/* ------------------------------------------------------------------------------------------
 * Drivers.  kin = nsteps x 9 row-major table [t, alpha, h, alphadot, hdot, u, udot, U, W]
 * produced on the host by update_kinem (calcs.jl:97-198) / update_externalvel (:75-92) with
 * the running sum t = t + dt (solvers.jl:180).  mat = nsteps x 8 row-major
 * [t, alpha, h, u, a0, cl, cd, cm] (solvers.jl:255).
 * driver: 0 ldvm (solvers.jl:144-268), 1 ldvm2DOF (:657-788), 2 ldvmLin (:448-655),
 *         3 lautat (:42-142)
 * ---------------------------------------------------------------------------------------- */
static void set_kinem_from_table(o_sim *m, const double *row) {
    m->t = row[0];
    m->s.alpha = row[1]; m->s.h = row[2]; m->s.alphadot = row[3]; m->s.hdot = row[4];
    m->s.u = row[5]; m->s.udot = row[6];
    m->f.u = row[7]; m->f.w = row[8];
}

static void step_ldvm(o_sim *m, int two_dof, double dt, const double *row, int mode,
                      int dkind, long dlimit, double ddist, double dtol, double *matrow) {
    o_surf *s = &m->s; o_field *f = &m->f;
    if (two_dof) {
        m->t = row[0]; f->u = row[7]; f->w = row[8];                    /* solvers.jl:699-702 */
        update_kinem2dof(s, &m->sp, &m->k2, dt, m->cl, m->cm);           /* :705 */
    } else {
        set_kinem_from_table(m, row);                                    /* :180-186 */
    }
    update_boundpos(s, dt);                                              /* :189 */
    place_tev(s, f, dt);                                                 /* :192 */
    update_indbound(s, f);                                               /* :195 */
    double x1[2] = { -0.01, 0 };
    root_solve(m, mode, fun1, 1, x1);                                    /* :198-199 */
    f->tev.v[f->tev.n - 1].s = x1[0];                                    /* :200 */
    if (two_dof) update_atermdot(s, dt); else update_a2a3adot(s, dt);    /* :203 / :723 */
    double lesp = s->a0;
    if (fabs(lesp) > s->lespcrit) {                                      /* :208 */
        place_lev(s, f, dt);                                             /* :212 */
        double x2[2] = { -0.01, 0.01 };
        m->lesp_force = (mode == 0) ? (lesp > 0 ? 1 : -1) : 0;
        root_solve(m, mode, fun2, 2, x2);                                /* :215-216 */
        m->lesp_force = 0;
        f->tev.v[f->tev.n - 1].s = x2[0];                                /* :217 */
        f->lev.v[f->lev.n - 1].s = x2[1];
        s->levflag = 1;                                                  /* :220 */
    } else {
        s->levflag = 0;                                                  /* :222 */
    }
    update_a2toan(s);                                                    /* :226 */
    s->a0prev = s->a0;                                                   /* :229-232 / :750-753 */
    int nprev = two_dof ? s->naterm : 3;
    for (int ia = 0; ia < nprev; ia++) s->aprev[ia] = s->aterm[ia];
    update_bv(s);                                                        /* :235 */
    int mid = (int)nearbyint(s->ndiv / 2.0) - 1;                         /* Int(round(ndiv/2)) :238 */
    control_vort_count(dkind, dlimit, ddist, dtol, s->bnd_x[mid], s->bnd_z[mid], f);
    wakeroll(s, f, dt);                                                  /* :241 */
    double cl, cd, cm, cn;
    calc_forces(s, f->u, f->w, &cl, &cd, &cm, &cn);                      /* :244 */
    m->cl = cl; m->cd = cd; m->cm = cm;
    matrow[0] = m->t; matrow[1] = s->alpha; matrow[2] = s->h; matrow[3] = s->u;
    matrow[4] = s->a0; matrow[5] = cl; matrow[6] = cd; matrow[7] = cm;   /* :255 */
}

/* lautat step, solvers.jl:76-131 (no update_indbound, no LEV; aprev saved BEFORE update_a2toan) */
static void step_lautat(o_sim *m, double dt, const double *row, int mode,
                        int dkind, long dlimit, double ddist, double dtol, double *matrow) {
    o_surf *s = &m->s; o_field *f = &m->f;
    double fu = f->u, fw = f->w;
    set_kinem_from_table(m, row);                                        /* :78-81 */
    f->u = fu; f->w = fw;              /* lautat never calls update_externalvel */
    update_boundpos(s, dt);                                              /* :84 */
    place_tev(s, f, dt);                                                 /* :87 */
    double x1[2] = { -0.01, 0 };
    root_solve(m, mode, fun1, 1, x1);                                    /* :90-91 */
    f->tev.v[f->tev.n - 1].s = x1[0];                                    /* :92 */
    update_a2a3adot(s, dt);                                              /* :95 */
    s->a0prev = s->a0;                                                   /* :98-101 */
    for (int ia = 0; ia < 3; ia++) s->aprev[ia] = s->aterm[ia];
    update_a2toan(s);                                                    /* :104 */
    update_bv(s);                                                        /* :107 */
    int mid = (int)nearbyint(s->ndiv / 2.0) - 1;
    control_vort_count(dkind, dlimit, ddist, dtol, s->bnd_x[mid], s->bnd_z[mid], f);  /* :110 */
    wakeroll(s, f, dt);                                                  /* :113-115 */
    double cl, cd, cm, cn;
    calc_forces(s, f->u, f->w, &cl, &cd, &cm, &cn);                      /* :118 */
    m->cl = cl; m->cd = cd; m->cm = cm;
    matrow[0] = m->t; matrow[1] = s->alpha; matrow[2] = s->h; matrow[3] = s->u;
    matrow[4] = s->a0; matrow[5] = cl; matrow[6] = cd; matrow[7] = cm;   /* :129 */
}

/* ldvmLin step, solvers.jl:488-644 (with the `0ml` typo at :451 read as 0) */
static void step_ldvmlin(o_sim *m, double dt, const double *row,
                         int dkind, long dlimit, double ddist, double dtol, double *matrow) {
    o_surf *s = &m->s; o_field *f = &m->f;
    int nd = s->ndiv, na = s->naterm;
    double vcore = 0.02 * s->c;                                          /* :481 */
    double T1[nd], T2[nd], T3[nd], tmp[nd];
    set_kinem_from_table(m, row);                                        /* :490-496 */
    update_boundpos(s, dt);                                              /* :499 */
    update_indbound(s, f);                                               /* :502 */
    update_downwash(s, f->u, f->w);                                      /* :505 */
    for (int i = 0; i < nd; i++) T1[i] = s->downwash[i];                 /* :510 */
    for (int i = 0; i < nd; i++) tmp[i] = T1[i] * (cos(s->theta[i]) - 1.);
    double I1 = s->c * simple_trapz(tmp, s->theta, nd);                  /* :511 */
    double J1 = -simple_trapz(T1, s->theta, nd) / (s->uref * O_PI);      /* :512 */
    long ntev = f->tev.n;
    double xloc_tev, zloc_tev;
    if (ntev == 0) {                                                     /* :517-523 */
        xloc_tev = s->bnd_x[nd - 1] + 0.5 * s->u * dt;
        zloc_tev = s->bnd_z[nd - 1];
    } else {
        xloc_tev = s->bnd_x[nd - 1] + (1. / 3.) * (f->tev.v[ntev - 1].x - s->bnd_x[nd - 1]);
        zloc_tev = s->bnd_z[nd - 1] + (1. / 3.) * (f->tev.v[ntev - 1].z - s->bnd_z[nd - 1]);
    }
    for (int ib = 0; ib < nd; ib++) {                                    /* :525-530 */
        double xdist = s->bnd_x[ib] - xloc_tev, zdist = s->bnd_z[ib] - zloc_tev;
        double distsq = xdist * xdist + zdist * zdist;
        T2[ib] = (s->cam_slope[ib] * zdist + xdist) / (2 * O_PI * sqrt(distsq * distsq + vcore * vcore * vcore * vcore));
    }
    double sig_prev = -s->uref * s->c * O_PI * (s->a0prev + s->aprev[0] / 2.);  /* :533 */
    for (int i = 0; i < nd; i++) tmp[i] = T2[i] * (cos(s->theta[i]) - 1.);
    double I2 = simple_trapz(tmp, s->theta, nd);                         /* :535 */
    double J2 = -simple_trapz(T2, s->theta, nd) / (O_PI * s->uref);      /* :536 */
    double tevstr = -(I1 + sig_prev) / (1 + I2);                         /* :538 */
    s->a0 = J1 + J2 * tevstr;                                            /* :541 */
    double c1[na], c2[na], c3[na];
    for (int ia = 1; ia <= na; ia++) {                                   /* :542-544 */
        for (int i = 0; i < nd; i++) tmp[i] = T1[i] * cos(ia * s->theta[i]);
        c1[ia - 1] = simple_trapz(tmp, s->theta, nd);
        for (int i = 0; i < nd; i++) tmp[i] = T2[i] * cos(ia * s->theta[i]);
        c2[ia - 1] = simple_trapz(tmp, s->theta, nd);
        s->aterm[ia - 1] = 2. * (c1[ia - 1] + tevstr * c2[ia - 1]) / (O_PI * s->uref);
    }
    s->a0dot = (s->a0 - s->a0prev) / dt;                                 /* :547-550 */
    for (int ia = 0; ia < na; ia++) s->adot[ia] = (s->aterm[ia] - s->aprev[ia]) / dt;
    if (fabs(s->a0) > s->lespcrit) {                                     /* :553 */
        double lesp_cond = (s->a0 >= 0.) ? s->lespcrit : -s->lespcrit;   /* :554-558 */
        double xloc_lev, zloc_lev;
        lev_location(s, f, dt, &xloc_lev, &zloc_lev);                    /* :561-570 */
        for (int ib = 0; ib < nd; ib++) {                                /* :572-577 */
            double xdist = s->bnd_x[ib] - xloc_lev, zdist = s->bnd_z[ib] - zloc_lev;
            double distsq = xdist * xdist + zdist * zdist;
            T3[ib] = (s->cam_slope[ib] * zdist + xdist) / (2 * O_PI * sqrt(distsq * distsq + vcore * vcore * vcore * vcore));
        }
        for (int i = 0; i < nd; i++) tmp[i] = T3[i] * (cos(s->theta[i]) - 1.);
        double I3 = simple_trapz(tmp, s->theta, nd);                     /* :578 */
        double J3 = -simple_trapz(T3, s->theta, nd) / (O_PI * s->uref);  /* :579 */
        double det = J3 * (I2 + 1.) - J2 * (I3 + 1.);                    /* :581 */
        tevstr = (-J3 * (I1 + sig_prev) + (I3 + 1) * (J1 - lesp_cond)) / det;   /* :583 */
        double levstr = (J2 * (I1 + sig_prev) - (I2 + 1) * (J1 - lesp_cond)) / det;  /* :584 */
        s->a0 = J1 + J2 * tevstr + J3 * levstr;                          /* :587 */
        for (int ia = 1; ia <= na; ia++) {                               /* :588-592, :597-601 */
            for (int i = 0; i < nd; i++) tmp[i] = T3[i] * cos(ia * s->theta[i]);
            c3[ia - 1] = simple_trapz(tmp, s->theta, nd);
            s->aterm[ia - 1] = 2. * (c1[ia - 1] + tevstr * c2[ia - 1] + levstr * c3[ia - 1]) / (O_PI * s->uref);
        }
        o_vort qt = { xloc_tev, zloc_tev, tevstr, vcore, 0., 0. };       /* :594-595 */
        o_vort ql = { xloc_lev, zloc_lev, levstr, vcore, 0., 0. };
        vec_push(&f->tev, qt);
        vec_push(&f->lev, ql);
        s->levflag = 1;                                                  /* :603 */
    } else {
        o_vort qt = { xloc_tev, zloc_tev, tevstr, vcore, 0., 0. };       /* :605 */
        vec_push(&f->tev, qt);
        s->levflag = 0;                                                  /* :612 */
    }
    s->a0prev = s->a0;                                                   /* :616-619 */
    for (int ia = 0; ia < 3; ia++) s->aprev[ia] = s->aterm[ia];
    update_bv(s);                                                        /* :622 */
    int mid = (int)nearbyint(s->ndiv / 2.0) - 1;
    control_vort_count(dkind, dlimit, ddist, dtol, s->bnd_x[mid], s->bnd_z[mid], f);  /* :625 */
    wakeroll(s, f, dt);                                                  /* :628 */
    double cl, cd, cm, cn;
    calc_forces(s, f->u, f->w, &cl, &cd, &cm, &cn);                      /* :631 */
    m->cl = cl; m->cd = cd; m->cm = cm;
    matrow[0] = m->t; matrow[1] = s->alpha; matrow[2] = s->h; matrow[3] = s->u;
    matrow[4] = s->a0; matrow[5] = cl; matrow[6] = cd; matrow[7] = cm;   /* :642 */
}

/* ------------------------------------------------------------------------------------------
 * Flat ctypes-facing API
 *
 * surf pack (doubles): [c, uref, pvt, lespcrit, levflag, alpha, h, alphadot, hdot, u, udot,
 *                       a0, a0dot, a0prev]  (14)
 *   then ndiv each: theta, x, cam, cam_slope, bnd_x, bnd_z, uind, wind, downwash
 *   then naterm each: aterm, adot, aprev
 *   then (ndiv-1) each: bv.x, bv.z, bv.s, bv.vc, bv.vx, bv.vz
 * ---------------------------------------------------------------------------------------- */
#define O_SURF_HDR 14
long o_surf_pack_len(int ndiv, int naterm) { return O_SURF_HDR + 9L * ndiv + 3L * naterm + 6L * (ndiv - 1); }

void *o_sim_create(int ndiv, int naterm) {
    o_sim *m = (o_sim *)calloc(1, sizeof(o_sim));
    o_surf *s = &m->s;
    s->ndiv = ndiv; s->naterm = naterm;
    double **arr[] = { &s->theta, &s->x, &s->cam, &s->cam_slope, &s->bnd_x, &s->bnd_z, &s->uind, &s->wind, &s->downwash };
    for (int k = 0; k < 9; k++) *arr[k] = (double *)calloc((size_t)ndiv, sizeof(double));
    s->aterm = (double *)calloc((size_t)naterm, sizeof(double));
    s->adot = (double *)calloc((size_t)naterm, sizeof(double));
    s->aprev = (double *)calloc((size_t)naterm, sizeof(double));
    s->bv = (o_vort *)calloc((size_t)(ndiv - 1), sizeof(o_vort));
    return m;
}

void o_sim_destroy(void *h) {
    o_sim *m = (o_sim *)h;
    o_surf *s = &m->s;
    free(s->theta); free(s->x); free(s->cam); free(s->cam_slope); free(s->bnd_x); free(s->bnd_z);
    free(s->uind); free(s->wind); free(s->downwash); free(s->aterm); free(s->adot); free(s->aprev); free(s->bv);
    free(m->f.tev.v); free(m->f.lev.v); free(m->f.extv.v);
    free(m);
}

void o_sim_set_surf(void *h, const double *p) {
    o_sim *m = (o_sim *)h; o_surf *s = &m->s;
    int nd = s->ndiv, na = s->naterm;
    s->c = p[0]; s->uref = p[1]; s->pvt = p[2]; s->lespcrit = p[3]; s->levflag = (int)p[4];
    s->alpha = p[5]; s->h = p[6]; s->alphadot = p[7]; s->hdot = p[8]; s->u = p[9]; s->udot = p[10];
    s->a0 = p[11]; s->a0dot = p[12]; s->a0prev = p[13];
    const double *q = p + O_SURF_HDR;
    double *arr[] = { s->theta, s->x, s->cam, s->cam_slope, s->bnd_x, s->bnd_z, s->uind, s->wind, s->downwash };
    for (int k = 0; k < 9; k++) { memcpy(arr[k], q, (size_t)nd * sizeof(double)); q += nd; }
    memcpy(s->aterm, q, (size_t)na * sizeof(double)); q += na;
    memcpy(s->adot, q, (size_t)na * sizeof(double)); q += na;
    memcpy(s->aprev, q, (size_t)na * sizeof(double)); q += na;
    int nb = nd - 1;
    for (int i = 0; i < nb; i++) {
        s->bv[i].x = q[i]; s->bv[i].z = q[nb + i]; s->bv[i].s = q[2 * nb + i];
        s->bv[i].vc = q[3 * nb + i]; s->bv[i].vx = q[4 * nb + i]; s->bv[i].vz = q[5 * nb + i];
    }
}

void o_sim_get_surf(void *h, double *p) {
    o_sim *m = (o_sim *)h; o_surf *s = &m->s;
    int nd = s->ndiv, na = s->naterm;
    p[0] = s->c; p[1] = s->uref; p[2] = s->pvt; p[3] = s->lespcrit; p[4] = s->levflag;
    p[5] = s->alpha; p[6] = s->h; p[7] = s->alphadot; p[8] = s->hdot; p[9] = s->u; p[10] = s->udot;
    p[11] = s->a0; p[12] = s->a0dot; p[13] = s->a0prev;
    double *q = p + O_SURF_HDR;
    double *arr[] = { s->theta, s->x, s->cam, s->cam_slope, s->bnd_x, s->bnd_z, s->uind, s->wind, s->downwash };
    for (int k = 0; k < 9; k++) { memcpy(q, arr[k], (size_t)nd * sizeof(double)); q += nd; }
    memcpy(q, s->aterm, (size_t)na * sizeof(double)); q += na;
    memcpy(q, s->adot, (size_t)na * sizeof(double)); q += na;
    memcpy(q, s->aprev, (size_t)na * sizeof(double)); q += na;
    int nb = nd - 1;
    for (int i = 0; i < nb; i++) {
        q[i] = s->bv[i].x; q[nb + i] = s->bv[i].z; q[2 * nb + i] = s->bv[i].s;
        q[3 * nb + i] = s->bv[i].vc; q[4 * nb + i] = s->bv[i].vx; q[5 * nb + i] = s->bv[i].vz;
    }
}

static o_vec *family(o_sim *m, int fam) { return fam == 0 ? &m->f.tev : fam == 1 ? &m->f.lev : &m->f.extv; }

/* fam: 0 tev, 1 lev, 2 extv; v = 6 x n row-major [x; z; s; vc; vx; vz] */
void o_sim_set_vorts(void *h, int fam, long n, const double *v) {
    o_vec *a = family((o_sim *)h, fam);
    a->n = 0;
    for (long i = 0; i < n; i++) {
        o_vort q = { v[i], v[n + i], v[2 * n + i], v[3 * n + i], v[4 * n + i], v[5 * n + i] };
        vec_push(a, q);
    }
}
long o_sim_count(void *h, int fam) { return family((o_sim *)h, fam)->n; }
void o_sim_get_vorts(void *h, int fam, double *v) {
    o_vec *a = family((o_sim *)h, fam);
    long n = a->n;
    for (long i = 0; i < n; i++) {
        v[i] = a->v[i].x; v[n + i] = a->v[i].z; v[2 * n + i] = a->v[i].s;
        v[3 * n + i] = a->v[i].vc; v[4 * n + i] = a->v[i].vx; v[5 * n + i] = a->v[i].vz;
    }
}
void o_sim_set_field_uw(void *h, double u, double w) { ((o_sim *)h)->f.u = u; ((o_sim *)h)->f.w = w; }
void o_sim_set_2dof(void *h, const double *strpar11, const double *kin22) {
    o_sim *m = (o_sim *)h;
    memcpy(&m->sp, strpar11, sizeof(o_strpar));
    memcpy(&m->k2, kin22, sizeof(o_kin2dof));
}
void o_sim_get_2dof(void *h, double *kin22) { memcpy(kin22, &((o_sim *)h)->k2, sizeof(o_kin2dof)); }
void o_sim_set_forces(void *h, double cl, double cd, double cm) {
    o_sim *m = (o_sim *)h; m->cl = cl; m->cd = cd; m->cm = cm;
}
long o_sim_nfev(void *h) { return ((o_sim *)h)->nfev; }

/* calc_delcp(surf, vels) -> (p_com, p_in, p_out), src/lowOrder2D/postprocess.jl:184-222, and
 * calc_edgeVel(surf, vels) -> (q_com_u, q_com_l), :224-244: the five columns (after x) of the `delcp_edgevel`
 * step file (postprocess.jl:100-110).  out = [p_com | p_in | p_out | q_com_u | q_com_l], ndiv each.
 * rho = surf.rho (the LE radius parameter, typedefs.jl:68) is not part of the oracle's surf state and is passed in.
 * Note the reference's own asymmetry, restated literally: calc_delcp ignores `vels`, calc_edgeVel uses it. */
void o_sim_delcp_edgevel(void *h, double rho, double velx, double velz, double *out) {
    const o_surf *s = &((o_sim *)h)->s;
    const int nd = s->ndiv;
    double *p_com = out, *p_in = out + nd, *p_out = out + 2 * nd, *q_u = out + 3 * nd, *q_l = out + 4 * nd;
    for (int k = 0; k < 5 * nd; k++) out[k] = 0.0;
    for (int i = 0; i < nd; i++) {                                                      /* :193-197 */
        double udash = s->u * cos(s->alpha) + s->hdot * sin(s->alpha) + s->uind[i] * cos(s->alpha) - s->wind[i] * sin(s->alpha);
        p_in[i] = 8 * s->uref * sqrt(s->c) * s->a0 * sqrt(s->x[i]) * udash / (rho * s->c + 2 * s->x[i])
                  + 4 * s->uref * sqrt(s->c) * s->a0dot * sqrt(s->x[i]);
    }
    for (int i = 1; i < nd; i++) {                                                      /* :198-219 */
        double udash = s->u * cos(s->alpha) + s->hdot * sin(s->alpha) + s->uind[i] * cos(s->alpha) - s->wind[i] * sin(s->alpha);
        double th = s->theta[i];
        double gam = s->a0 * (1.0 / tan(th / 2)), gamint = s->a0dot * (th + sin(th)), gammod = 0.0;
        for (int n = 1; n <= s->naterm; n++) {
            gam += s->aterm[n - 1] * sin(n * th);
            gammod += s->aterm[n - 1] * sin(n * th);
            if (n == 1) gamint += s->adot[0] * (th / 2 - sin(2 * th) / 4);
            else gamint += s->adot[n - 1] / 2 * (sin((n - 1) * th) / (n - 1) - sin((n + 1) * th) / (n + 1));
        }
        gam = gam * s->uref;
        gammod = gammod * s->uref;
        gamint = gamint * s->uref * s->c;
        p_out[i] = 2 * udash * gam + gamint;
        p_com[i] = 2 * udash * s->uref * s->a0 * (cos(th / 2) - 1) / sin(th / 2) + 2 * udash * gammod + gamint
                   + 8 * s->uref * udash * s->c * s->a0 * sin(th / 2) / (rho * s->c + 2 * s->c * (sin(th / 2) * sin(th / 2)));
    }
    q_u[0] = s->uref * s->a0 / sqrt(0.5 * rho);                                          /* :230-231 */
    q_l[0] = s->uref * s->a0 / sqrt(0.5 * rho);
    for (int i = 1; i < nd; i++) {                                                      /* :233-241 */
        double udash = (s->u + velx) * cos(s->alpha) + (s->hdot - velz) * sin(s->alpha) + s->uind[i] * cos(s->alpha) - s->wind[i] * sin(s->alpha);
        double th = s->theta[i], gammod = 0.0;
        for (int n = 1; n <= s->naterm; n++) gammod += s->aterm[n - 1] * sin(n * th);
        gammod = gammod * s->uref;
        q_u[i] = s->uref * s->a0 * (cos(th / 2) - 1) / sin(th / 2) + s->uref * gammod
                 + (sqrt(s->c) * s->uref * s->a0 + sqrt(s->x[i]) * udash) / sqrt(s->x[i] + rho * s->c / 2);
        q_l[i] = -s->uref * s->a0 * (cos(th / 2) - 1) / sin(th / 2) - s->uref * gammod
                 + (-sqrt(s->c) * s->uref * s->a0 + sqrt(s->x[i]) * udash) / sqrt(s->x[i] + rho * s->c / 2);
    }
}

/* run nsteps of the chosen driver; returns 0 */
int o_sim_run(void *h, int driver, long nsteps, double dt, const double *kin /*nsteps x 9*/,
              int solver_mode, int dkind, long dlimit, double ddist, double dtol,
              double *mat /*nsteps x 8 row-major*/) {
    o_sim *m = (o_sim *)h;
    for (long is = 0; is < nsteps; is++) {
        const double *row = kin + 9 * is;
        double *mr = mat + 8 * is;
        switch (driver) {
            case 0: step_ldvm(m, 0, dt, row, solver_mode, dkind, dlimit, ddist, dtol, mr); break;
            case 1: step_ldvm(m, 1, dt, row, solver_mode, dkind, dlimit, ddist, dtol, mr); break;
            case 2: step_ldvmlin(m, dt, row, dkind, dlimit, ddist, dtol, mr); break;
            case 3: step_lautat(m, dt, row, solver_mode, dkind, dlimit, ddist, dtol, mr); break;
            default: return 1;
        }
    }
    return 0;
}

/* stand-alone pieces for unit parity tests */
void o_sim_wakeroll(void *h, double dt) { o_sim *m = (o_sim *)h; wakeroll(&m->s, &m->f, dt); }
void o_sim_update_indbound(void *h) { o_sim *m = (o_sim *)h; update_indbound(&m->s, &m->f); }
void o_sim_control_vort_count(void *h, int dkind, long dlimit, double ddist, double dtol) {
    o_sim *m = (o_sim *)h; o_surf *s = &m->s;
    int mid = (int)nearbyint(s->ndiv / 2.0) - 1;
    control_vort_count(dkind, dlimit, ddist, dtol, s->bnd_x[mid], s->bnd_z[mid], &m->f);
}

/* ------------------------------------------------------------------------------------------
 * Kinematics callables, src/kinem.jl (value and analytic d/dt; the reference obtains the
 * derivative with ForwardDiff.derivative, calcs.jl:102 etc., which evaluates the same
 * analytic chain rule).  kind: 0 ConstDef (:108-114), 1 SinDef (:142-151), 2 CosDef
 * (:154-163), 3 EldUpDef (:18-29), 4 EldUptstartDef (:32-44), 5 LinearDef (:117-132),
 * 6 StepGustDef (:165-177).   p = parameters in struct order.
 * ---------------------------------------------------------------------------------------- */
void o_kinem_eval(int kind, const double *p, double t, double *val, double *der) {
    switch (kind) {
        case 0: *val = p[0]; *der = 0; break;
        case 1: *val = p[0] + p[1] * sin(2 * p[2] * t + p[3]); *der = p[1] * cos(2 * p[2] * t + p[3]) * 2 * p[2]; break;
        case 2: *val = p[0] + p[1] * cos(2 * p[2] * t + p[3]); *der = -p[1] * sin(2 * p[2] * t + p[3]) * 2 * p[2]; break;
        case 3: case 4: {
            double amp = p[0], K = p[1], a = p[2];
            double sm = O_PI * O_PI * K / (2 * amp * (1 - a));
            double t1 = (kind == 3) ? 1. : p[3];
            double t2 = t1 + (amp / (2 * K));
            *val = ((K / sm) * log(cosh(sm * (t - t1)) / cosh(sm * (t - t2)))) + (amp / 2);
            *der = K * (tanh(sm * (t - t1)) - tanh(sm * (t - t2)));
            break;
        }
        case 5: {
            double ts = p[0], vs = p[1], ve = p[2], len = p[3];
            if (t < ts) { *val = vs; *der = 0; }
            else if (t > ts + len) { *val = ve; *der = 0; }
            else { *val = vs + (ve - vs) / len * (t - ts); *der = (ve - vs) / len; }
            break;
        }
        case 6: *val = (t >= p[1] && t <= p[1] + p[2]) ? p[0] : 0.; *der = 0; break;
        default: *val = 0; *der = 0;
    }
}

/* ==========================================================================================
 * Multi-surface LDVM: ldvm(surf::Vector{TwoDSurf}, ...), src/lowOrder2D/solvers.jl:270-445
 * (SURVEY 8f rank 3).  Restated literally, including the way the Mult functors only ever feed
 * a surface's OWN newly shed vortices back into its induced velocities (typedefs.jl:279-287).
 * ======================================================================================== */
typedef struct {
    int nsurf;
    o_surf *s;
    o_field f;
    double t;
    long nfev;
    int shed[O_NMAX];   /* shed_ind (0-based surface indices) */
    int nshed;
    int lesp_force[O_NMAX];   /* exact-root mode: per-surface LESP branch fixed to the triggering A0's sign (0 = per evaluation) */
} o_msim;

/* place_tev(surf::Vector{TwoDSurf}, field, dt), calcs.jl:389-406 */
static void place_tev_multi(o_msim *m, double dt) {
    const int ns = m->nsurf;
    const long ntev = m->f.tev.n;
    for (int i = 0; i < ns; i++) {
        o_surf *s = &m->s[i];
        const int nd = s->ndiv;
        double xloc, zloc;
        if (ntev < ns) {
            xloc = s->bnd_x[nd - 1] + 0.5 * s->u * dt;
            zloc = s->bnd_z[nd - 1];
        } else {
            const o_vort *q = &m->f.tev.v[ntev - ns + i];
            xloc = s->bnd_x[nd - 1] + (1. / 3.) * (q->x - s->bnd_x[nd - 1]);
            zloc = s->bnd_z[nd - 1] + (1. / 3.) * (q->z - s->bnd_z[nd - 1]);
        }
        o_vort v = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
        vec_push(&m->f.tev, v);
    }
}

static int in_shed(const o_msim *m, int i) {
    for (int k = 0; k < m->nshed; k++) if (m->shed[k] == i) return 1;
    return 0;
}

/* place_lev(surf::Vector{TwoDSurf}, field, dt, shed_ind), calcs.jl:428-450 */
static void place_lev_multi(o_msim *m, double dt) {
    const int ns = m->nsurf;
    const long nlev = m->f.lev.n;
    for (int i = 0; i < ns; i++) {
        o_surf *s = &m->s[i];
        if (in_shed(m, i)) {
            double xloc, zloc;
            if (s->levflag == 0) {
                double le_vel_x = s->u - s->alphadot * sin(s->alpha) * s->pvt * s->c + s->uind[0];
                double le_vel_z = -s->alphadot * cos(s->alpha) * s->pvt * s->c - s->hdot + s->wind[0];
                xloc = s->bnd_x[0] + 0.5 * le_vel_x * dt;
                zloc = s->bnd_z[0] + 0.5 * le_vel_z * dt;
            } else {
                const o_vort *q = &m->f.lev.v[nlev - ns + i];
                xloc = s->bnd_x[0] + (1. / 3.) * (q->x - s->bnd_x[0]);
                zloc = s->bnd_z[0] + (1. / 3.) * (q->z - s->bnd_z[0]);
            }
            o_vort v = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
            vec_push(&m->f.lev, v);
        } else {
            o_vort v = { 0., 0., 0., 0., 0., 0. };      /* placeholder, calcs.jl:446 */
            vec_push(&m->f.lev, v);
        }
    }
}

/* [tev_list; (lev_list;) bv_list] of the Mult functors for surface i, typedefs.jl:270-279, :364-375 */
static long mult_sources(const o_msim *m, int i, int with_lev, o_vort *out) {
    const int ns = m->nsurf;
    long k = 0;
    for (int j = 0; j < ns; j++) out[k++] = m->f.tev.v[m->f.tev.n - ns + j];
    if (with_lev)
        for (int q = 0; q < m->nshed; q++) out[k++] = m->f.lev.v[m->f.lev.n - ns + m->shed[q]];
    for (int j = 0; j < ns; j++)
        if (j != i)
            for (int b = 0; b < m->s[j].ndiv - 1; b++) out[k++] = m->s[j].bv[b];
    return k;
}

static double kelvin_val(const o_surf *s) {
    return s->uref * s->c * O_PI * (s->a0 + s->aterm[0] / 2.) - s->uref * s->c * O_PI * (s->a0prev + s->aprev[0] / 2.);
}

/* shared body of KelvinConditionMult (typedefs.jl:261-305) and KelvinKuttaMult (:356-420) */
static void mult_functor(o_msim *m, const double *v_iter, double *val, int with_lev) {
    const int ns = m->nsurf;
    long cap = 2 * ns;
    for (int j = 0; j < ns; j++) cap += m->s[j].ndiv;
    o_vort *src = (o_vort *)malloc((size_t)cap * sizeof(o_vort));
    int levcount = 0;
    for (int i = 0; i < ns; i++) {
        o_surf *s = &m->s[i];
        const int nd = s->ndiv;
        double uprev[nd], wprev[nd], unow[nd], wnow[nd];
        long n = mult_sources(m, i, with_lev, src);
        ind_vel_aos(src, n, s->bnd_x, s->bnd_z, nd, uprev, wprev);
        m->f.tev.v[m->f.tev.n - ns + i].s = v_iter[i];
        if (with_lev && in_shed(m, i)) {
            levcount += 1;
            m->f.lev.v[m->f.lev.n - ns + i].s = v_iter[ns + levcount - 1];
        }
        n = mult_sources(m, i, with_lev, src);
        ind_vel_aos(src, n, s->bnd_x, s->bnd_z, nd, unow, wnow);
        for (int b = 0; b < nd; b++) {
            s->uind[b] = s->uind[b] - uprev[b] + unow[b];
            s->wind[b] = s->wind[b] - wprev[b] + wnow[b];
        }
        update_downwash(s, m->f.u, m->f.w);
    }
    levcount = 0;
    for (int i = 0; i < ns; i++) {
        o_surf *s = &m->s[i];
        update_a0anda1(s);
        update_a2toan(s);
        update_bv(s);
        const double tevs = m->f.tev.v[m->f.tev.n - ns + i].s;
        if (with_lev && in_shed(m, i)) {
            val[i] = kelvin_val(s) + tevs + m->f.lev.v[m->f.lev.n - ns + i].s;
            levcount += 1;
            double lesp_cond = (s->a0 > 0) ? s->lespcrit : -s->lespcrit;
            if (m->lesp_force[i]) lesp_cond = m->lesp_force[i] * s->lespcrit;
            val[levcount - 1 + ns] = s->a0 - lesp_cond;
        } else {
            val[i] = kelvin_val(s) + tevs;
        }
    }
    m->nfev++;
    free(src);
}
static void mfun1(void *m, const double *x, double *r) { mult_functor((o_msim *)m, x, r, 0); }
static void mfun2(void *m, const double *x, double *r) { mult_functor((o_msim *)m, x, r, 1); }

/* wakeroll(surf::Vector{TwoDSurf}, curfield, dt), calcs.jl:296-372 */
static void wakeroll_multi(o_msim *m, double dt) {
    o_field *f = &m->f;
    long ntev = f->tev.n, nlev = f->lev.n, nextv = f->extv.n, n = ntev + nlev + nextv;
    o_vort *all = (o_vort *)malloc((size_t)(n > 0 ? n : 1) * sizeof(o_vort));
    memcpy(all, f->tev.v, (size_t)ntev * sizeof(o_vort));
    memcpy(all + ntev, f->lev.v, (size_t)nlev * sizeof(o_vort));
    memcpy(all + ntev + nlev, f->extv.v, (size_t)nextv * sizeof(o_vort));
    for (long i = 0; i < n; i++) { all[i].vx = 0; all[i].vz = 0; }
    mutual_ind_aos(all, n);
    double *tx = (double *)malloc((size_t)(n > 0 ? n : 1) * 6 * sizeof(double));
    double *tz = tx + n, *ut = tz + n, *wt = ut + n, *us = wt + n, *ws = us + n;
    for (long i = 0; i < n; i++) { tx[i] = all[i].x; tz[i] = all[i].z; us[i] = 0; ws[i] = 0; }
    for (int is = 0; is < m->nsurf; is++) {                              /* :325-329 */
        ind_vel_aos(m->s[is].bv, m->s[is].ndiv - 1, tx, tz, n, ut, wt);
        for (long i = 0; i < n; i++) { us[i] += ut[i]; ws[i] += wt[i]; }
    }
    for (long i = 0; i < n; i++) { all[i].vx += us[i]; all[i].vz += ws[i]; }
    for (long i = 0; i < n; i++) { all[i].vx += f->u; all[i].vz += f->w; }
    for (long i = 0; i < n; i++) { all[i].x += dt * all[i].vx; all[i].z += dt * all[i].vz; }
    memcpy(f->tev.v, all, (size_t)ntev * sizeof(o_vort));
    memcpy(f->lev.v, all + ntev, (size_t)nlev * sizeof(o_vort));
    memcpy(f->extv.v, all + ntev + nlev, (size_t)nextv * sizeof(o_vort));
    free(tx);
    free(all);
}

/* one step of ldvm(surf::Vector{TwoDSurf}), solvers.jl:311-434.
 * kin: nsurf rows of 9 [t, alpha, h, alphadot, hdot, u, udot, U, W]; matrow: 1 + 7*nsurf */
static void step_ldvm_multi(o_msim *m, double dt, const double *kin, int mode,
                            int dkind, long dlimit, double ddist, double dtol, double *matrow) {
    const int ns = m->nsurf;
    o_field *f = &m->f;
    m->t = kin[0]; f->u = kin[7]; f->w = kin[8];                         /* :313-316 */
    for (int is = 0; is < ns; is++) {                                    /* :318-324 */
        o_surf *s = &m->s[is];
        const double *r = kin + 9 * is;
        s->alpha = r[1]; s->h = r[2]; s->alphadot = r[3]; s->hdot = r[4]; s->u = r[5]; s->udot = r[6];
        update_boundpos(s, dt);
    }
    place_tev_multi(m, dt);                                              /* :327 */
    for (int is = 0; is < ns; is++) {                                    /* :329-338 */
        o_surf *s = &m->s[is];
        update_indbound(s, f);
        for (int js = 0; js < ns; js++) {
            if (is == js) continue;
            const int nd = s->ndiv;
            double u[nd], w[nd];
            ind_vel_aos(m->s[js].bv, m->s[js].ndiv - 1, s->bnd_x, s->bnd_z, nd, u, w);   /* add_indbound_b, calcs.jl:40-46 */
            for (int b = 0; b < nd; b++) { s->uind[b] += u[b]; s->wind[b] += w[b]; }
        }
    }
    double x[O_NMAX];
    for (int i = 0; i < ns; i++) x[i] = -0.01;                           /* :343 */
    root_solve(m, mode, mfun1, ns, x);
    m->nshed = 0;
    for (int i = 0; i < ns; i++) {                                       /* :345-362 */
        f->tev.v[f->tev.n - ns + i].s = x[i];
        update_a2a3adot(&m->s[i], dt);
        if (fabs(m->s[i].a0) > m->s[i].lespcrit) m->shed[m->nshed++] = i;
    }
    if (m->nshed > 0) {                                                  /* :365-390 */
        place_lev_multi(m, dt);
        for (int i = 0; i < ns + m->nshed; i++) x[i] = -0.01;            /* :372 */
        for (int i = 0; i < ns; i++) m->lesp_force[i] = (mode == 0) ? (m->s[i].a0 > 0 ? 1 : -1) : 0;
        root_solve(m, mode, mfun2, ns + m->nshed, x);
        for (int i = 0; i < ns; i++) m->lesp_force[i] = 0;
        /* strengths: the functor has already stored the iterate in the vortices; the driver's
         * own assignment block (:374-390) re-stores soln.zero for the index patterns it handles */
        int lc = 0;
        for (int i = 0; i < ns; i++) {
            f->tev.v[f->tev.n - ns + i].s = x[i];
            if (in_shed(m, i)) { f->lev.v[f->lev.n - ns + i].s = x[ns + lc]; lc++; m->s[i].levflag = 1; }
            else m->s[i].levflag = 0;
        }
    } else {
        for (int i = 0; i < ns; i++) m->s[i].levflag = 0;                /* :392-394 */
    }
    for (int i = 0; i < ns; i++) {                                       /* :397-403 */
        o_surf *s = &m->s[i];
        s->a0prev = s->a0;
        for (int ia = 0; ia < 3; ia++) s->aprev[ia] = s->aterm[ia];
    }
    double mx = 0, mz = 0;                                               /* :406-408 */
    for (int i = 0; i < ns; i++) {
        const int mid = (int)nearbyint(m->s[i].ndiv / 2.0) - 1;
        mx += m->s[i].bnd_x[mid]; mz += m->s[i].bnd_z[mid];
    }
    control_vort_count(dkind, dlimit, ddist, dtol, mx / ns, mz / ns, f);
    wakeroll_multi(m, dt);                                               /* :411 */
    matrow[0] = m->t;
    for (int i = 0; i < ns; i++) {                                       /* :414-432 */
        double cl, cd, cm, cn;
        calc_forces(&m->s[i], f->u, f->w, &cl, &cd, &cm, &cn);
        double *r = matrow + 1 + 7 * i;
        r[0] = m->s[i].alpha; r[1] = m->s[i].h; r[2] = m->s[i].u; r[3] = m->s[i].a0; r[4] = cl; r[5] = cd; r[6] = cm;
    }
}

void *o_msim_create(int nsurf, int ndiv, int naterm) {
    o_msim *m = (o_msim *)calloc(1, sizeof(o_msim));
    m->nsurf = nsurf;
    m->s = (o_surf *)calloc((size_t)nsurf, sizeof(o_surf));
    for (int i = 0; i < nsurf; i++) {
        o_sim *tmp = (o_sim *)o_sim_create(ndiv, naterm);
        m->s[i] = tmp->s;
        free(tmp);
    }
    return m;
}
void o_msim_destroy(void *h) {
    o_msim *m = (o_msim *)h;
    for (int i = 0; i < m->nsurf; i++) {
        o_surf *s = &m->s[i];
        free(s->theta); free(s->x); free(s->cam); free(s->cam_slope); free(s->bnd_x); free(s->bnd_z);
        free(s->uind); free(s->wind); free(s->downwash); free(s->aterm); free(s->adot); free(s->aprev); free(s->bv);
    }
    free(m->s); free(m->f.tev.v); free(m->f.lev.v); free(m->f.extv.v); free(m);
}
/* reuse the single-surface pack/unpack through a temporary o_sim view of surface i */
void o_msim_set_surf(void *h, int i, const double *p) {
    o_msim *m = (o_msim *)h; o_sim v; memset(&v, 0, sizeof(v)); v.s = m->s[i];
    o_sim_set_surf(&v, p); m->s[i] = v.s;
}
void o_msim_get_surf(void *h, int i, double *p) {
    o_msim *m = (o_msim *)h; o_sim v; memset(&v, 0, sizeof(v)); v.s = m->s[i];
    o_sim_get_surf(&v, p);
}
static o_vec *mfamily(o_msim *m, int fam) { return fam == 0 ? &m->f.tev : fam == 1 ? &m->f.lev : &m->f.extv; }
void o_msim_set_vorts(void *h, int fam, long n, const double *v) {
    o_vec *a = mfamily((o_msim *)h, fam);
    a->n = 0;
    for (long i = 0; i < n; i++) { o_vort q = { v[i], v[n + i], v[2 * n + i], v[3 * n + i], v[4 * n + i], v[5 * n + i] }; vec_push(a, q); }
}
long o_msim_count(void *h, int fam) { return mfamily((o_msim *)h, fam)->n; }
void o_msim_get_vorts(void *h, int fam, double *v) {
    o_vec *a = mfamily((o_msim *)h, fam);
    long n = a->n;
    for (long i = 0; i < n; i++) {
        v[i] = a->v[i].x; v[n + i] = a->v[i].z; v[2 * n + i] = a->v[i].s;
        v[3 * n + i] = a->v[i].vc; v[4 * n + i] = a->v[i].vx; v[5 * n + i] = a->v[i].vz;
    }
}
void o_msim_set_field_uw(void *h, double u, double w) { ((o_msim *)h)->f.u = u; ((o_msim *)h)->f.w = w; }
long o_msim_nfev(void *h) { return ((o_msim *)h)->nfev; }
/* kin: nsteps x nsurf x 9 row-major; mat: nsteps x (1 + 7 nsurf) row-major */
int o_msim_run(void *h, long nsteps, double dt, const double *kin, int solver_mode,
               int dkind, long dlimit, double ddist, double dtol, double *mat) {
    o_msim *m = (o_msim *)h;
    const int ns = m->nsurf, nc = 1 + 7 * ns;
    if (ns > O_NMAX / 2) return 1;
    for (long is = 0; is < nsteps; is++)
        step_ldvm_multi(m, dt, kin + (size_t)is * ns * 9, solver_mode, dkind, dlimit, ddist, dtol, mat + (size_t)is * nc);
    return 0;
}

/* ==========================================================================================
 * LVE -- lumped-vortex-element panel solver, src/lowOrder2D/LVE.jl:6-302 with its helpers
 * src/lowOrder2D/calcs.jl:652-836.  Restated without the plotting /
 * animation / global-variable scaffolding.  Two wakes are kept like the reference: IFR_field
 * (inertial frame, convected by wakeroll) and curfield (body frame, positions refreshed by
 * vor_BFR for the NEXT time level).
 * kin table per iteration (nsteps+1 rows x 9): [t, alpha, h, alphadot, hdot, u, udot, X0, Z0]
 * with alpha = kindef.alpha(t) (= kinem.alpha), X0 = -kindef.u(t)*t, Z0 = kindef.h(t)
 * (calcs.jl:654-656), and one extra row for t + dtstar of the last iteration (vor_BFR).
 * ======================================================================================== */
typedef struct {
    o_surf s;            /* body-frame surf (bv positions = vor_loc) */
    o_vort *ifr_bv;      /* IFR_surf.bv */
    o_field ifr, cur;    /* IFR_field, curfield */
    int nd;
    double *ds, *cam_slope_deg, *nx, *nz, *taux, *tauz, *vorx, *vorz, *colx, *colz;
    double *prevCirc, *circChange, *circ, *RHS, *relx, *relz, *dp;
    double gamma_crit, rho, lev_stop, t_start, te_x, te_cam, le_x, le_cam, ifr_u0, ifr_hdot0;
} o_lve;

/* IFR / BFR, calcs.jl:652-676 (alpha, X0, Z0 supplied from the table) */
static void lve_ifr(double pvt, double al, double X0, double Z0, double x, double z, double *X, double *Z) {
    *X = (x - pvt) * cos(al) + z * sin(al) + X0 + pvt;
    *Z = -(x - pvt) * sin(al) + z * cos(al) + Z0;
}
static void lve_bfr(double pvt, double al, double X0, double Z0, double X, double Z, double *x, double *z) {
    *x = (X - X0 - pvt) * cos(al) - (Z - Z0) * sin(al) + pvt;
    *z = (X - X0 - pvt) * sin(al) + (Z - Z0) * cos(al);
}

/* dense solve a\RHS (Julia: LU with partial pivoting), n <= 256 */
static void dense_solve(int n, const double *a /*row-major n x n*/, const double *b, double *x) {
    double *A = (double *)malloc((size_t)n * (n + 1) * sizeof(double));
    for (int i = 0; i < n; i++) { memcpy(A + (size_t)i * (n + 1), a + (size_t)i * n, (size_t)n * sizeof(double)); A[(size_t)i * (n + 1) + n] = b[i]; }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(A[(size_t)r * (n + 1) + c]) > fabs(A[(size_t)piv * (n + 1) + c])) piv = r;
        if (piv != c) for (int j = 0; j <= n; j++) { double t = A[(size_t)c * (n + 1) + j]; A[(size_t)c * (n + 1) + j] = A[(size_t)piv * (n + 1) + j]; A[(size_t)piv * (n + 1) + j] = t; }
        for (int r = c + 1; r < n; r++) {
            const double f = A[(size_t)r * (n + 1) + c] / A[(size_t)c * (n + 1) + c];
            for (int j = c; j <= n; j++) A[(size_t)r * (n + 1) + j] -= f * A[(size_t)c * (n + 1) + j];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double v = A[(size_t)i * (n + 1) + n];
        for (int j = i + 1; j < n; j++) v -= A[(size_t)i * (n + 1) + j] * x[j];
        x[i] = v / A[(size_t)i * (n + 1) + i];
    }
    free(A);
}

/* influence_coeff, calcs.jl:756-787 (+ the LEV modification of mod_influence_coeff :789-836) */
static void lve_influence(const o_lve *L, double x_w, double z_w, double *a) {
    const int nd = L->nd;
    const double vc = L->s.bv[0].vc;
    memset(a, 0, (size_t)nd * nd * sizeof(double));
    for (int j = 0; j < nd; j++) a[(size_t)(nd - 1) * nd + j] = 1.;
    for (int i = 0; i < nd - 1; i++)
        for (int j = 0; j < nd; j++) {
            const double t_x = L->colx[i], t_z = L->colz[i];
            double u, w;
            if (j < nd - 1) {
                const double xdist = L->s.bv[j].x - t_x, zdist = L->s.bv[j].z - t_z;
                const double distsq = xdist * xdist + zdist * zdist;
                u = -zdist / (2 * O_PI * distsq);
                w = xdist / (2 * O_PI * distsq);
            } else {
                const double xdist = x_w - t_x, zdist = z_w - t_z;
                const double distsq = xdist * xdist + zdist * zdist;
                u = -zdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
                w = xdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
            }
            a[(size_t)i * nd + j] = u * L->nx[i] + w * L->nz[i];
        }
}

void *o_lve_create(int ndiv, int naterm, const double *surfpack, double rho, double lev_stop,
                   double ifr_xend, double ifr_camend, double ifr_u0, double ifr_hdot0) {
    o_lve *L = (o_lve *)calloc(1, sizeof(o_lve));
    o_sim *tmp = (o_sim *)o_sim_create(ndiv, naterm);
    o_sim_set_surf(tmp, surfpack);
    L->s = tmp->s;
    free(tmp);
    const int nd = L->nd = ndiv, np = ndiv - 1;
    o_surf *s = &L->s;
    L->rho = rho; L->lev_stop = lev_stop; L->t_start = 0;
    L->te_x = ifr_xend; L->te_cam = ifr_camend; L->ifr_u0 = ifr_u0; L->ifr_hdot0 = ifr_hdot0;
    L->le_x = s->x[0]; L->le_cam = s->cam[0];
    double **arr[] = { &L->ds, &L->cam_slope_deg, &L->nx, &L->nz, &L->taux, &L->tauz, &L->vorx, &L->vorz, &L->colx, &L->colz,
                       &L->prevCirc, &L->circChange, &L->circ, &L->RHS, &L->relx, &L->relz, &L->dp };
    for (size_t k = 0; k < sizeof(arr) / sizeof(arr[0]); k++) *arr[k] = (double *)calloc((size_t)nd, sizeof(double));
    L->ifr_bv = (o_vort *)calloc((size_t)np, sizeof(o_vort));
    const double d2r = O_PI / 180.0;
    for (int j = 0; j < np; j++) {                                   /* LVE.jl:29-47 */
        const double dx = s->x[j] - s->x[j + 1], dc = s->cam[j] - s->cam[j + 1];
        L->ds[j] = sqrt(dx * dx + dc * dc);
        L->cam_slope_deg[j] = asin((s->cam[j + 1] - s->cam[j]) / L->ds[j]) / d2r;     /* asind */
        const double sd = sin(L->cam_slope_deg[j] * d2r), cd = cos(L->cam_slope_deg[j] * d2r);   /* sind / cosd */
        L->nx[j] = -sd; L->nz[j] = cd; L->taux[j] = cd; L->tauz[j] = sd;
        L->vorx[j] = s->x[j] + .25 * L->ds[j] * cd; L->vorz[j] = s->cam[j] + .25 * L->ds[j] * sd;
        L->colx[j] = s->x[j] + .75 * L->ds[j] * cd; L->colz[j] = s->cam[j] + .75 * L->ds[j] * sd;
        s->bv[j].x = L->vorx[j]; s->bv[j].z = L->vorz[j];             /* refresh_vortex :55 */
        L->ifr_bv[j] = s->bv[j];
    }
    const double x = 1 - 2 * L->ds[0] / s->c;                         /* :50-52 */
    L->gamma_crit = s->lespcrit / 1.13 * (s->u * s->c * (acos(x) + sin(acos(x))));
    return L;
}

void o_lve_destroy(void *h) {
    o_lve *L = (o_lve *)h;
    o_surf *s = &L->s;
    free(s->theta); free(s->x); free(s->cam); free(s->cam_slope); free(s->bnd_x); free(s->bnd_z);
    free(s->uind); free(s->wind); free(s->downwash); free(s->aterm); free(s->adot); free(s->aprev); free(s->bv);
    free(L->ds); free(L->cam_slope_deg); free(L->nx); free(L->nz); free(L->taux); free(L->tauz); free(L->vorx); free(L->vorz);
    free(L->colx); free(L->colz); free(L->prevCirc); free(L->circChange); free(L->circ); free(L->RHS); free(L->relx); free(L->relz);
    free(L->dp); free(L->ifr_bv);
    free(L->ifr.tev.v); free(L->ifr.lev.v); free(L->ifr.extv.v); free(L->cur.tev.v); free(L->cur.lev.v); free(L->cur.extv.v);
    free(L);
}

/* iterations i = 0..nsteps of LVE.jl:60-298; kin has nsteps+2 rows (last = t+dtstar of the final
 * iteration); mat = nsteps x 8 row-major [t, alpha(deg), h, u, a0, cl, cd, cm] */
int o_lve_run(void *h, long nsteps, double dtstar, const double *kin, double *mat) {
    o_lve *L = (o_lve *)h;
    o_surf *s = &L->s;
    const int nd = L->nd, np = nd - 1;
    double *a = (double *)malloc((size_t)nd * nd * sizeof(double));
    double *a1 = (double *)malloc((size_t)nd * sizeof(double));
    const double d2r = O_PI / 180.0;
    for (long i = 0; i <= nsteps; i++) {
        const double *r = kin + 9 * i, *rn = kin + 9 * (i + 1);
        const double t = r[0], al = r[1], X0 = r[7], Z0 = r[8];
        s->alpha = r[1]; s->h = r[2]; s->alphadot = r[3]; s->hdot = r[4]; s->u = r[5]; s->udot = r[6];   /* update_kinem :64 */
        for (int j = 0; j < np; j++)                                                                   /* :67-68 */
            lve_ifr(s->pvt, al, X0, Z0, L->vorx[j], L->vorz[j], &L->ifr_bv[j].x, &L->ifr_bv[j].z);
        {   /* place_tev2(IFR_surf, IFR_field, dtstar, t), calcs.jl:688-709 */
            double TE_x, TE_z, xloc, zloc;
            lve_ifr(s->pvt, al, X0, Z0, L->te_x, L->te_cam, &TE_x, &TE_z);
            if (L->ifr.tev.n == 0) { xloc = TE_x + 0.5 * L->ifr_u0 * dtstar; zloc = TE_z + 0.5 * L->ifr_hdot0 * dtstar; }
            else {
                const o_vort *q = &L->ifr.tev.v[L->ifr.tev.n - 1];
                xloc = TE_x + (1. / 3.) * (q->x - TE_x); zloc = TE_z + (1. / 3.) * (q->z - TE_z);
            }
            o_vort v = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
            vec_push(&L->ifr.tev, v);
        }
        double x_w, z_w;
        { const o_vort *q = &L->ifr.tev.v[L->ifr.tev.n - 1]; lve_bfr(s->pvt, al, X0, Z0, q->x, q->z, &x_w, &z_w); }   /* :74 */
        lve_influence(L, x_w, z_w, a);                                                                 /* :75 */
        double sp = 0; for (int j = 0; j < np; j++) sp += L->prevCirc[j];
        L->RHS[nd - 1] = sp;                                                                           /* :78 */
        if (L->cur.tev.n > 0) {                                                                        /* :80-82 */
            long n = L->cur.tev.n + L->cur.lev.n;
            o_vort *all = (o_vort *)malloc((size_t)n * sizeof(o_vort));
            memcpy(all, L->cur.tev.v, (size_t)L->cur.tev.n * sizeof(o_vort));
            memcpy(all + L->cur.tev.n, L->cur.lev.v, (size_t)L->cur.lev.n * sizeof(o_vort));
            ind_vel_aos(all, n, L->colx, L->colz, np, s->uind, s->wind);
            free(all);
        }
        for (int j = 0; j < np; j++) {                                                                 /* :83-90 */
            const double ux = cos(al) * s->u - sin(al) * (-s->hdot), uz = sin(al) * s->u + cos(al) * (-s->hdot);
            L->relx[j] = ux + (-s->alphadot * L->colz[j]);
            L->relz[j] = uz + s->alphadot * (L->colx[j] - s->pvt);
            L->RHS[j] = -((L->relx[j] + s->uind[j]) * L->nx[j] + (L->relz[j] + s->wind[j]) * L->nz[j]);
        }
        dense_solve(nd, a, L->RHS, L->circ);                                                           /* :93 */
        for (int j = 0; j < np; j++) {                                                                 /* :96-100 */
            double g1 = 0, g2 = 0;
            for (int k = 0; k < j; k++) { g1 += L->prevCirc[k]; g2 += L->circ[k]; }
            L->circChange[j] = (g2 - g1) / dtstar;
        }
        const double xx = 1 - 2 * L->ds[0] / s->c;
        if (fabs(L->circ[0]) > L->gamma_crit) {                                                        /* :104 */
            double gamma_crit_use;
            if (s->levflag == 0 && t < L->lev_stop) {                                                  /* :105-108 */
                gamma_crit_use = fabs(L->circ[0]) / L->circ[0] * L->gamma_crit;
                L->t_start = t;
            } else {                                                                                   /* :109-115 */
                const double m_slp = -s->lespcrit / (L->lev_stop - L->t_start);
                const double c_slp = s->lespcrit - (m_slp * L->t_start);
                const double LESP_crit_use = (m_slp * t) + c_slp;
                gamma_crit_use = fabs(L->circ[0]) / L->circ[0] * LESP_crit_use / 1.13 * (s->u * s->c * (acos(xx) + sin(acos(xx))));
            }
            {   /* place_lev2(surf, IFR_field, dtstar, t), calcs.jl:712-730 */
                double LE_x, LE_z, xloc, zloc;
                lve_ifr(s->pvt, al, X0, Z0, L->le_x, L->le_cam, &LE_x, &LE_z);
                const double le_vel_x = s->u - s->alphadot * sin(s->alpha) * s->pvt * s->c + s->uind[0];
                const double le_vel_z = -s->alphadot * cos(s->alpha) * s->pvt * s->c - s->hdot + s->wind[0];
                if (s->levflag == 0) { xloc = LE_x + 0.5 * le_vel_x * dtstar; zloc = LE_z + 0.5 * le_vel_z * dtstar; }
                else {
                    const o_vort *q = &L->ifr.lev.v[L->ifr.lev.n - 1];
                    xloc = LE_x + (1. / 3.) * (q->x - LE_x); zloc = LE_z + (1. / 3.) * (q->z - LE_z);
                }
                o_vort v = { xloc, zloc, 0., 0.02 * s->c, 0., 0. };
                vec_push(&L->ifr.lev, v);
            }
            s->levflag = 1;                                                                            /* :123 */
            double x_lev, z_lev;
            { const o_vort *q = &L->ifr.lev.v[L->ifr.lev.n - 1]; lve_bfr(s->pvt, al, X0, Z0, q->x, q->z, &x_lev, &z_lev); }
            /* mod_influence_coeff, calcs.jl:789-836 */
            lve_influence(L, x_w, z_w, a);
            for (int ii = 0; ii < nd; ii++) a1[ii] = a[(size_t)ii * nd];
            for (int ii = 0; ii < nd; ii++) for (int j = 0; j < nd - 1; j++) a[(size_t)ii * nd + j] = a[(size_t)ii * nd + j + 1];
            {
                const double vc = s->bv[0].vc;
                for (int ii = 0; ii < np; ii++) {
                    const double xdist = x_lev - L->colx[ii], zdist = z_lev - L->colz[ii];
                    const double distsq = xdist * xdist + zdist * zdist;
                    const double u = -zdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
                    const double w = xdist / (2 * O_PI * sqrt(vc * vc * vc * vc + distsq * distsq));
                    a[(size_t)ii * nd + nd - 1] = u * L->nx[ii] + w * L->nz[ii];
                }
            }
            for (int j = 0; j < np; j++)                                                               /* :134-137 */
                L->RHS[j] = -((L->relx[j] + s->uind[j]) * L->nx[j] + (L->relz[j] + s->wind[j]) * L->nz[j]) - gamma_crit_use * a1[j];
            L->RHS[nd - 1] -= gamma_crit_use;                                                          /* :138 */
            dense_solve(nd, a, L->RHS, L->circ);                                                       /* :140 */
            L->ifr.lev.v[L->ifr.lev.n - 1].s = L->circ[nd - 1];                                        /* :142 */
            for (int j = nd - 1; j >= 1; j--) L->circ[j] = L->circ[j - 1];                             /* :144 */
            L->circ[0] = gamma_crit_use;                                                               /* :145 */
        } else {
            s->levflag = 0;                                                                            /* :151 */
        }
        for (int j = 0; j < np; j++) { L->prevCirc[j] = L->circ[j]; s->bv[j].s = L->circ[j]; L->ifr_bv[j].s = L->circ[j]; }   /* :154-159 */
        if (L->ifr.tev.n > 1) L->ifr.tev.v[L->ifr.tev.n - 1].s = L->circ[nd - 1];                      /* :160-162 */
        if (i > 0) {                                                                                   /* :203-232 */
            for (int j = 0; j < np; j++)
                L->dp[j] = L->rho * (((L->relx[j] + s->uind[j]) * L->taux[j] + (L->relz[j] + s->wind[j]) * L->tauz[j]) * L->circ[j] / L->ds[j]
                                     + L->circChange[j]);
            double ndem = .5 * L->rho * s->u * s->u * s->c;
            double cn = 0; for (int j = 0; j < np; j++) cn += L->dp[j] * L->ds[j];
            cn /= ndem;
            s->a0 = 1.13 * L->circ[0] / (s->u * s->c * (acos(xx) + sin(acos(xx))));
            const double cs = 2 * O_PI * s->a0 * s->a0;
            const double cl = cn * cos(al) + cs * sin(al), cd = cn * sin(al) - cs * cos(al);
            ndem = ndem * s->c;
            double m1 = 0, m2 = 0;
            for (int j = 0; j < np; j++) {
                m1 += L->dp[j] * L->ds[j] * cos(L->cam_slope_deg[j] * d2r) * (L->vorx[j] - s->pvt);
                m2 += L->dp[j] * L->ds[j] * sin(L->cam_slope_deg[j] * d2r) * L->vorz[j];
            }
            const double cm = -m1 / ndem + m2 / ndem;
            double *mr = mat + 8 * (i - 1);
            mr[0] = t; mr[1] = s->alpha * 180 / O_PI; mr[2] = s->h; mr[3] = s->u; mr[4] = s->a0; mr[5] = cl; mr[6] = cd; mr[7] = cm;
        }
        {   /* wakeroll(IFR_surf, IFR_field, dtstar) :235 -- reuse the single-surface restatement */
            o_surf tmp = *s;
            tmp.bv = L->ifr_bv;
            wakeroll(&tmp, &L->ifr, dtstar);
        }
        {   /* :238-242 */
            o_vort v = { 0, 0, L->ifr.tev.v[L->ifr.tev.n - 1].s, 0.02 * s->c, 0., 0. };
            vec_push(&L->cur.tev, v);
            if (L->ifr.lev.n > 0 && s->levflag == 1) {
                o_vort w = { 0, 0, L->ifr.lev.v[L->ifr.lev.n - 1].s, 0.02 * s->c, 0., 0. };
                vec_push(&L->cur.lev, w);
            }
            /* vor_BFR(surf, t+dtstar, IFR_field, curfield), calcs.jl:732-754 */
            for (long k = 0; k < L->ifr.tev.n; k++)
                lve_bfr(s->pvt, rn[1], rn[7], rn[8], L->ifr.tev.v[k].x, L->ifr.tev.v[k].z, &L->cur.tev.v[k].x, &L->cur.tev.v[k].z);
            for (long k = 0; k < L->ifr.lev.n; k++)
                lve_bfr(s->pvt, rn[1], rn[7], rn[8], L->ifr.lev.v[k].x, L->ifr.lev.v[k].z, &L->cur.lev.v[k].x, &L->cur.lev.v[k].z);
        }
    }
    free(a); free(a1);
    return 0;
}

long o_lve_count(void *h, int which) {   /* 0 ifr.tev, 1 ifr.lev, 2 cur.tev, 3 cur.lev */
    o_lve *L = (o_lve *)h;
    return which == 0 ? L->ifr.tev.n : which == 1 ? L->ifr.lev.n : which == 2 ? L->cur.tev.n : L->cur.lev.n;
}
void o_lve_get_vorts(void *h, int which, double *v) {
    o_lve *L = (o_lve *)h;
    o_vec *a = which == 0 ? &L->ifr.tev : which == 1 ? &L->ifr.lev : which == 2 ? &L->cur.tev : &L->cur.lev;
    long n = a->n;
    for (long i = 0; i < n; i++) {
        v[i] = a->v[i].x; v[n + i] = a->v[i].z; v[2 * n + i] = a->v[i].s;
        v[3 * n + i] = a->v[i].vc; v[4 * n + i] = a->v[i].vx; v[5 * n + i] = a->v[i].vz;
    }
}
void o_lve_get_circ(void *h, double *circ_nd, double *gamma_crit) {
    o_lve *L = (o_lve *)h;
    memcpy(circ_nd, L->circ, (size_t)L->nd * sizeof(double));
    *gamma_crit = L->gamma_crit;
}
