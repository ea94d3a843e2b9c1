/*
 * unsflow.h -- C ABI of libunsflow_cuda: the B200 (sm_100a, FP64) discrete-vortex time-step
 * hot path of UnsteadyFlowSolvers.jl (UNSflow), callable from Julia with `ccall`.
 *
 * The reference has NO FFI for this path (pure Julia, no ccall anywhere); the boundary is the
 * set of Julia method signatures listed below.  Each entry point cites the reference method
 * (file:line under the reference checkout) whose body it replaces.  The Julia-side binding is
 * julia/UnsflowCUDA.jl; tests drive the same symbols through ctypes.
 *
 * Conventions
 *   - plain C, extern "C"; arrays are contiguous `double*`; counts are int64_t.
 *   - every function returns int: 0 = ok, non-zero = error; unsflow_last_error() returns a
 *     thread-local message.  No C++ exception crosses the ABI.  The Julia shim turns a
 *     non-zero status into `error(msg)` (the reference throws, e.g. solvers.jl:157).
 *   - "host" entry points take HOST pointers, copy in, run on the GPU, copy out and return
 *     after synchronising (Julia arrays are only GC-rooted for the duration of the ccall).
 *   - "dev"/handle entry points keep state resident in HBM between calls.
 *   - there is no CPU fallback: without a CUDA device every compute entry returns an error.
 *   - one handle = one owner thread at a time; distinct handles may be used concurrently.
 */
#ifndef UNSFLOW_H
#define UNSFLOW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNSFLOW_VERSION 100

/* ---- status / device ------------------------------------------------------------------ */
const char *unsflow_last_error(void);
int unsflow_version(void);
int unsflow_device_count(int *count);
int unsflow_set_device(int device);
/* SM count, SM clock (kHz), L2 bytes, total HBM bytes of the current device */
int unsflow_device_info(int *sm_count, int *sm_clock_khz, int64_t *l2_bytes, int64_t *hbm_bytes);

/* Accuracy of 1/sqrt(vc^4 + d^4) in the induction kernels (process-wide, affects later launches):
 * FULL (default): one MUFU.RSQ64H seed + a cubic step, 2.7e-16 relative error per pair (measured) --
 *   the same accuracy class as the reference's IEEE sqrt and divide;
 * FAST: seed + a bias-centred Newton step, <= 6.5e-13 per pair (measured), two FP64 instructions less
 *   per pair (+9 % throughput); still inside the 1e-12 parity tolerance but no longer round-off exact. */
#define UNSFLOW_PRECISION_FULL 0
#define UNSFLOW_PRECISION_FAST 1
int unsflow_set_precision(int mode);

/* ---- stateless kernels on HOST buffers ------------------------------------------------ */

/* ind_vel(src::Vector{TwoDVort}, t_x, t_z) -> (uind, wind)   src/lowOrder2D/calcs.jl:462-477
 * (and the one-vortex method :480-488 with ns = 1).  s = strength, vc = core radius. */
int unsflow_ind_vel(int64_t ns, const double *sx, const double *sz, const double *ss, const double *svc,
                    int64_t nt, const double *tx, const double *tz, double *u_out, double *w_out);

/* mutual_ind(vorts) -> vorts   src/lowOrder2D/calcs.jl:491-510.  ACCUMULATES into vx, vz
 * exactly like the reference (the caller zeroes them, calcs.jl:229-242). */
int unsflow_mutual_ind(int64_t n, const double *x, const double *z, const double *s, const double *vc,
                       double *vx_inout, double *vz_inout);

/* wakeroll(surf::TwoDSurf, curfield, dt) -> curfield   src/lowOrder2D/calcs.jl:222-294.
 * Free vortices are passed concatenated in the reference's order [tev; lev; extv] (:245);
 * bv_* are surf.bv (ndiv-1 records); (u_inf, w_inf) = curfield.u[1], curfield.w[1].
 * On return vx, vz hold the convection velocity and x, z are advanced by dt (explicit Euler). */
int unsflow_wakeroll(int64_t n, double *x_inout, double *z_inout, const double *s, const double *vc,
                     double *vx_out, double *vz_out,
                     int64_t nbv, const double *bv_x, const double *bv_z, const double *bv_s, const double *bv_vc,
                     double u_inf, double w_inf, double dt);

/* update_indbound(surf, curfield) -> surf   src/lowOrder2D/calcs.jl:35-38
 * uind, wind (ndiv) induced at the bound points by all free vortices [tev; lev; extv]. */
int unsflow_update_indbound(int64_t n, const double *x, const double *z, const double *s, const double *vc,
                            int64_t ndiv, const double *bnd_x, const double *bnd_z,
                            double *uind_out, double *wind_out);

/* The four entry points above keep their device scratch (staging slabs, partial planes) per calling
 * thread between calls; this frees it. */
int unsflow_release_scratch(void);

/* ---- device-resident wake (config "synthetic large-N rollup"; multi-GPU target slices) -- */
typedef struct unsflow_wake unsflow_wake;

/* kernel variants for unsflow_wake_create */
#define UNSFLOW_KERNEL_AUTO      0   /* symmetric pair tiles when vc is uniform, else direct */
#define UNSFLOW_KERNEL_DIRECT    1   /* one-sided target x source tiles (13 FP64 instr / interaction) */
#define UNSFLOW_KERNEL_SYMMETRIC 2   /* pair tiles, each kernel evaluation feeds both vortices */

/* Create a device-resident wake of n free vortices + nbv bound vortices.
 * rank/nranks: this process owns one GPU (the current device) and one slice of the pair work;
 * nranks = 1 for a single GPU.  nccl_unique_id: 128 bytes from unsflow_nccl_unique_id() of
 * rank 0, identical on all ranks (ignored when nranks == 1). */
int unsflow_nccl_unique_id(void *id128);
/* The share of the step owned by `rank` of `nranks` (pure arithmetic, no GPU needed):
 * DIRECT: target index range [begin,end) of the n free vortices; SYMMETRIC: range of work
 * units, a unit being 1/32 of a pair tile of the row-major upper-triangle list of 2048-wide tiles. */
int unsflow_partition(int64_t n, int nranks, int rank, int variant, int64_t *begin, int64_t *end);
int unsflow_wake_create(unsflow_wake **out, int64_t n, int64_t nbv, int rank, int nranks,
                        const void *nccl_unique_id, int kernel_variant);
/* upload / download full state (host buffers, each n or nbv doubles; NULL = skip) */
int unsflow_wake_upload(unsflow_wake *w, const double *x, const double *z, const double *s, const double *vc,
                        const double *bv_x, const double *bv_z, const double *bv_s, const double *bv_vc);
int unsflow_wake_download(unsflow_wake *w, double *x, double *z, double *vx, double *vz);
/* nsteps of wakeroll (calcs.jl:222-294) entirely on device; with nranks > 1 each step ends
 * with the NCCL exchange so that every rank again holds all positions.  Asynchronous on the
 * handle's stream; unsflow_wake_sync() waits. */
int unsflow_wake_step(unsflow_wake *w, int64_t nsteps, double u_inf, double w_inf, double dt);
int unsflow_wake_sync(unsflow_wake *w);
/* CUDA-event time (ms) of the last unsflow_wake_step call: whole call, and the induction
 * kernel(s) alone (sum over the call); valid after unsflow_wake_sync */
int unsflow_wake_timing(unsflow_wake *w, float *ms_total, float *ms_induce, int64_t *n_launches);
int unsflow_wake_destroy(unsflow_wake *w);

/* ---- device-resident LDVM stepper ------------------------------------------------------- */

/* SoA view of a Vector{TwoDVort} (typedefs.jl:11-18); n records, each pointer n doubles */
typedef struct unsflow_vorts {
    int64_t n;
    double *x, *z, *s, *vc, *vx, *vz;
} unsflow_vorts;

/* TwoDSurf (typedefs.jl:38-176): pointers alias the Julia Vector{Float64} fields in place.
 * Used read-only by create and written by get_state. */
typedef struct unsflow_surf {
    int32_t ndiv, naterm;                 /* :43-44 */
    double c, uref, pvt;                  /* :39-42 */
    double lespcrit;                      /* lespcrit[1] :63 */
    int32_t levflag;                      /* levflag[1] :64 */
    int32_t _pad;
    double kinem[6];                      /* KinemPar alpha,h,alphadot,hdot,u,udot  kinem.jl:3-10 */
    double a0, a0dot, a0prev;             /* a0[1], a0dot[1], a0prev[1] :56-60 */
    double *theta, *x, *cam, *cam_slope;  /* ndiv :46-49 */
    double *bnd_x, *bnd_z;                /* ndiv :51-52 */
    double *uind, *wind, *downwash;       /* ndiv :53-55 */
    double *aterm, *adot, *aprev;         /* naterm :57-61 */
    unsflow_vorts bv;                     /* ndiv-1 :62 */
} unsflow_surf;

/* TwoDFlowField (typedefs.jl:20-36) */
typedef struct unsflow_field {
    double u, w;                          /* u[1], w[1] */
    unsflow_vorts tev, lev, extv;
} unsflow_field;

#define UNSFLOW_DRIVER_LDVM     0   /* ldvm      solvers.jl:144-268 */
#define UNSFLOW_DRIVER_LDVM2DOF 1   /* ldvm2DOF  solvers.jl:657-788 */
#define UNSFLOW_DRIVER_LDVMLIN  2   /* ldvmLin   solvers.jl:448-655 */
#define UNSFLOW_DRIVER_LAUTAT   3   /* lautat    solvers.jl:42-142  */

#define UNSFLOW_SOLVE_EXACT   0     /* state left at the root of the (affine) Kelvin/Kutta residual */
#define UNSFLOW_SOLVE_NLSOLVE 1     /* NLsolve trust-region iteration restated on the device; state left at its last finite-difference probe */

#define UNSFLOW_DEL_NONE    0       /* delNone     delVort.jl:3 */
#define UNSFLOW_DEL_SPALART 1       /* delSpalart  delVort.jl:5-9 */

typedef struct unsflow_opts {
    int32_t driver;         /* UNSFLOW_DRIVER_* */
    int32_t solve_mode;     /* UNSFLOW_SOLVE_* */
    int32_t delvort_kind;   /* UNSFLOW_DEL_* */
    int32_t device;         /* CUDA device ordinal, -1 = current */
    int64_t del_limit;      /* delSpalart.limit */
    double del_dist;        /* delSpalart.dist */
    double del_tol;         /* delSpalart.tol */
    double dt;              /* dtstar*c/uref  solvers.jl:161 */
    int64_t max_steps;      /* capacity: total steps this handle will ever run */
} unsflow_opts;

typedef struct unsflow_ldvm unsflow_ldvm;

/* Upload surf + curfield and allocate the wake for max_steps more steps. */
int unsflow_ldvm_create(unsflow_ldvm **out, const unsflow_surf *surf, const unsflow_field *field,
                        const unsflow_opts *opts);
/* Host-tabulated prescribed motion for the next steps, row-major nrows x 9:
 * [t, alpha, h, alphadot, hdot, u, udot, U_ext, W_ext] -- the values update_kinem
 * (calcs.jl:97-198) and update_externalvel (:75-92) would set at the running-sum times
 * t = t + dt (solvers.jl:180).  Rows are consumed one per step.  For ldvm2DOF only the
 * columns t, U_ext, W_ext are read (alpha/h come from the structural solve on device). */
int unsflow_ldvm_set_kinematics(unsflow_ldvm *h, int64_t nrows, const double *table);
/* Multi-GPU stepping (one process per GPU, call once before the first run with the same state on
 * every rank): the pair tiles of the wake roll-up are split over ranks and the partial velocities
 * are summed with an NCCL all-reduce each step; the O(ndiv) kernels and the bound sweep are
 * replicated bit-identically, so every rank returns the same mat and state.  nccl_unique_id:
 * 128 bytes from unsflow_nccl_unique_id() of rank 0.  Needs a uniform core radius. */
int unsflow_ldvm_set_parallel(unsflow_ldvm *h, int rank, int nranks, const void *nccl_unique_id);
/* rank / nranks of the handle and whether the step exchanges through NVLink peer memory (1: the ranks' buffers are
 * mapped into each other with CUDA IPC and one fused kernel sums the partial velocities, convects and writes every
 * rank's slabs; 0: NCCL all-reduce -- taken by all ranks when the mapping could not be set up on one of them) */
int unsflow_ldvm_parallel_info(unsflow_ldvm *h, int *rank, int *nranks, int *peer_exchange);
/* TwoDOFPar (typedefs.jl:178-190, 11 doubles in field order) and KinemPar2DOF
 * (typedefs.jl:193-220, 22 doubles in field order) for ldvm2DOF */
int unsflow_ldvm_set_2dof(unsflow_ldvm *h, const double *strpar11, const double *kinem22);
int unsflow_ldvm_get_2dof(unsflow_ldvm *h, double *kinem22);
/* Run nsteps time steps on device (for istep = 1:nsteps, solvers.jl:178-257 / :697-778) and
 * return their `mat` rows [t, alpha, h, u, a0, cl, cd, cm] (solvers.jl:255) in mat_out,
 * COLUMN-major nsteps x 8 (a Julia Matrix{Float64}(undef, nsteps, 8)). */
int unsflow_ldvm_run(unsflow_ldvm *h, int64_t nsteps, double *mat_out);
/* current vortex counts (to size the arrays for get_state) */
int unsflow_ldvm_counts(unsflow_ldvm *h, int64_t *ntev, int64_t *nlev, int64_t *nextv);
/* Download the full state into caller-owned arrays (pointers may be NULL to skip);
 * surf->bv.n / field->*.n must hold the array capacities and are set to the counts. */
int unsflow_ldvm_get_state(unsflow_ldvm *h, unsflow_surf *surf, unsflow_field *field);
/* Multi-surface ldvm(surf::Vector{TwoDSurf}, curfield, ...)  solvers.jl:270-445 (driver must be
 * UNSFLOW_DRIVER_LDVM; 1 <= nsurf <= 4; all surfaces share ndiv, naterm and the theta grid).
 * `surfs` is an array of nsurf unsflow_surf.  The kinematics table of
 * unsflow_ldvm_set_kinematics then has nrows x nsurf x 9 entries (row-major; t, U_ext, W_ext are
 * read from surface 0's row) and unsflow_ldvm_run returns nsteps x (1 + 7 nsurf) columns
 * [t, (alpha, h, u, a0, cl, cd, cm) per surface] (solvers.jl:427-432), column-major. */
int unsflow_ldvm_multi_create(unsflow_ldvm **out, int32_t nsurf, const unsflow_surf *surfs,
                              const unsflow_field *field, const unsflow_opts *opts);
int unsflow_ldvm_multi_get_state(unsflow_ldvm *h, int32_t nsurf, unsflow_surf *surfs, unsflow_field *field);
/* LVE(surf, curfield, nsteps, dtstar, ...)  src/lowOrder2D/LVE.jl:6-302 (lumped-vortex-element panel
 * solver).  `surf` supplies the panel geometry (x, cam), pvt, lespcrit, initial kinem; ifr_field is
 * the (normally empty) inertial-frame wake; ifr_xend, ifr_camend = x[end], cam[end] and ifr_u0,
 * ifr_hdot0 = kinem.u, kinem.hdot of the reference's auxiliary IFR_surf (LVE.jl:17); rho = surf.rho;
 * lev_stop = LEV_Stop; opts.dt = dtstar (LVE does not rescale it), opts.max_steps = nsteps + 1.
 * unsflow_ldvm_set_kinematics takes (nsteps + 2) x 9 rows [t, alpha, h, alphadot, hdot, u, udot, X0, Z0]
 * (X0 = -kindef.u(t) t, Z0 = kindef.h(t), calcs.jl:654-656); one unsflow_ldvm_run step = one loop
 * iteration i = 0..nsteps, each returning a row of NINE columns [t, alpha(deg), h, u, a0, cl, cd, cm, lev]
 * (lev = 1 when an LEV was ejected in that iteration; zeros in the load columns for i = 0, where the
 * reference records nothing, LVE.jl:203).  The wake returned by
 * unsflow_ldvm_get_state is IFR_field; bv holds the inertial-frame bound vortices (IFR_surf.bv). */
int unsflow_lve_create(unsflow_ldvm **out, const unsflow_surf *surf, const unsflow_field *ifr_field,
                       const unsflow_opts *opts, double rho, double lev_stop, double ifr_xend,
                       double ifr_camend, double ifr_u0, double ifr_hdot0);
int unsflow_lve_get_circ(unsflow_ldvm *h, double *circ_ndiv, double *gamma_crit);
/* CUDA-event milliseconds of the last unsflow_ldvm_run and number of kernel launches in it */
int unsflow_ldvm_timing(unsflow_ldvm *h, float *ms_total, int64_t *n_launches);
int unsflow_ldvm_destroy(unsflow_ldvm *h);

/* ---- batched ensembles (independent members, no communication) -------------------------- */
/* Runs `nmember` independent LDVM simulations (one thread-block cluster per member, member
 * state in shared memory) that share geometry (surf) but differ in kinematics table and
 * lespcrit.  tables: nmember x nsteps x 9 row-major; lespcrit: nmember;
 * mat_out: nmember blocks of COLUMN-major nsteps x 8.  Members start from `surf` with an empty
 * wake.  To spread an ensemble over several GPUs/processes each rank passes its own shard of
 * tables / lespcrit / mat_out (no communication).  Prescribed-motion drivers only. */
int unsflow_ldvm_batch_run(const unsflow_surf *surf, const unsflow_opts *opts, int64_t nmember,
                           int64_t nsteps, const double *tables, const double *lespcrit,
                           double *mat_out, int64_t *ntev_out, int64_t *nlev_out);
/* Same for ldvm2DOF members (solvers.jl:657-788; e.g. a flutter-boundary sweep over structural
 * parameters): strpar11 = nmember x 11 (TwoDOFPar, typedefs.jl:178-190), kinem22 = nmember x 22 (initial
 * KinemPar2DOF, typedefs.jl:193-220), both in field order; kinem22_out (may be NULL) receives every member's final
 * KinemPar2DOF.  opts->driver must be UNSFLOW_DRIVER_LDVM2DOF; only the columns t, U_ext, W_ext of
 * `tables` are read.  delSpalart merging applies per member when opts asks for it. */
int unsflow_ldvm2dof_batch_run(const unsflow_surf *surf, const unsflow_opts *opts, int64_t nmember,
                               int64_t nsteps, const double *tables, const double *lespcrit,
                               const double *strpar11, const double *kinem22,
                               double *mat_out, double *kinem22_out, int64_t *ntev_out, int64_t *nlev_out);
/* CUDA-event milliseconds of the ensemble kernel of the last unsflow_ldvm_batch_run on this thread */
int unsflow_ldvm_batch_timing(float *ms_kernel);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* DFMA-chain microbenchmark: measured FP64 (non-tensor) peak of the current device, TFLOP/s
 * (2 flop per DFMA).  This is the roofline denominator for the induction kernels. */
int unsflow_fp64_peak(double *tflops, double *sm_mhz_est);
/* bound-point sweep (update_indbound, calcs.jl:35-38) micro-benchmark on synthetic device-resident
 * sources: best-of-reps milliseconds of the dedicated sweep kernel and of the generic kernel */
int unsflow_bench_sweep(int64_t n, int ndiv, int reps, float *ms_sweep, float *ms_generic);
/* relative error statistics of the rsqrt seed / refined value over a log-spaced q sweep */
int unsflow_rsqrt_probe(int64_t nsamples, double qmin, double qmax, double *max_rel_seed,
                        double *max_rel_newton, double *max_rel_cubic);

/* maximum relative error of ONE refinement variant over the same sweep, against a double-double truth:
 * order 2 = bias-centred Newton (FAST), 3 = Halley in FP64, 4/5/6 = Halley with the e^2 term in FP32 (measured
 * slower: kept as experiments), 7 = Halley in FP64 on a seed whose low word is donated by a dying register */
int unsflow_rsqrt_probe_order(int order, int64_t nsamples, double qmin, double qmax, double *max_rel);

#ifdef __cplusplus
}
#endif
#endif /* UNSFLOW_H */
