// api.cu -- C-ABI plumbing, stateless host-buffer entry points and measurement helpers.
#include "../../include/unsflow.h"
#include "induce_sym.cuh"
#include "sweep.cuh"
#include "runtime.h"
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

namespace unsflow {

thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("libunsflow_cuda: no usable CUDA device (%s); there is no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        (void)cudaGetLastError();
        return 3;
    }
    return 0;
}

static int g_order_override = 0;   // 0 = environment/default, 2 or 3 = set through the API

int rsqrt_order() {
    if (g_order_override) return g_order_override;
    static int order = [] {
        const char *e = getenv("UNSFLOW_RSQRT_ORDER");
        int o = e ? atoi(e) : 3;
        return (o == 2) ? 2 : 3;
    }();
    return order;
}

int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

// ---------------------------------------------------------------------------------------------
// One-shot induction on host buffers: stage -> k_induce_direct -> k_induce_finish -> copy out.
// Sources may be two groups (wake + bound vortices), each padded to a TS-aligned slab.
// ---------------------------------------------------------------------------------------------
struct HostGroup {
    int64_t n;
    const double *x, *z, *s, *vc;
};

static bool uniform_vc(const HostGroup *g, int ng, double *vc_out) {
    bool have = false;
    double v = 0;
    for (int k = 0; k < ng; k++)
        for (int64_t i = 0; i < g[k].n; i++) {
            if (!have) { v = g[k].vc[i]; have = true; }
            else if (g[k].vc[i] != v) return false;
        }
    *vc_out = have ? v : 1.0;
    return true;
}

struct Scratch {
    DevBuf<double> src, t, part, out;
    DevBuf<Regions> reg;
};
static thread_local Scratch g_scratch;

// targets == nullptr -> targets are the first source group (mutual induction)
static int induce_host(const HostGroup *g, int ng, int64_t nt, const double *tx, const double *tz,
                       double *u_inout, double *w_inout, int accumulate, double fs_u, double fs_w, double dt,
                       double *x_out, double *z_out) {
    if (int rc = require_device()) return rc;
    cudaStream_t st = 0;
    int64_t off[4] = {0, 0, 0, 0}, tot = 0;
    for (int k = 0; k < ng; k++) { off[k] = tot; tot += round_up(std::max<int64_t>(g[k].n, 1), kSymTB); }
    const bool self_targets = (tx == nullptr);
    if (self_targets) nt = g[0].n;
    double vcu = 1.0;
    const bool uni = uniform_vc(g, ng, &vcu);

    // per-thread scratch reused across calls: the drop-in entry points (ind_vel, mutual_ind, wakeroll,
    // update_indbound) are called once per time step by host-side solvers, and cudaMalloc/cudaFree of
    // the partial planes would otherwise cost as much as a mid-size kernel (freed by
    // unsflow_release_scratch or at thread exit)
    DevBuf<double> &d_src = g_scratch.src, &d_t = g_scratch.t, &d_part = g_scratch.part, &d_out = g_scratch.out;
    DevBuf<Regions> &d_reg = g_scratch.reg;
    const int narr = 4;
    if (d_src.ensure(narr * tot)) return 2;
    // host staging: [x | z | g | vc4], padding g = 0, vc4 = 1
    std::vector<double> h((size_t)(narr * tot), 0.0);
    for (int64_t i = 0; i < tot; i++) h[3 * tot + i] = 1.0;
    for (int k = 0; k < ng; k++) {
        if (g[k].n <= 0) continue;
        memcpy(&h[0 * tot + off[k]], g[k].x, sizeof(double) * g[k].n);
        memcpy(&h[1 * tot + off[k]], g[k].z, sizeof(double) * g[k].n);
        memcpy(&h[2 * tot + off[k]], g[k].s, sizeof(double) * g[k].n);
        for (int64_t i = 0; i < g[k].n; i++) {
            const double v = g[k].vc[i];
            h[3 * tot + off[k] + i] = v * v * v * v;
        }
    }
    UF_CUDA(cudaMemcpyAsync(d_src.p, h.data(), sizeof(double) * narr * tot, cudaMemcpyHostToDevice, st));

    const int64_t ntp = std::max<int64_t>(nt, 1);
    const double *d_tx, *d_tz;
    if (self_targets) {
        d_tx = d_src.p;
        d_tz = d_src.p + tot;
    } else {
        if (d_t.ensure(2 * ntp)) return 2;
        UF_CUDA(cudaMemcpyAsync(d_t.p, tx, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
        UF_CUDA(cudaMemcpyAsync(d_t.p + ntp, tz, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
        d_tx = d_t.p;
        d_tz = d_t.p + ntp;
    }

    Regions hreg[2];
    memset(hreg, 0, sizeof(hreg));
    for (int k = 0; k < ng; k++) { hreg[0].begin[k] = off[k]; hreg[0].end[k] = off[k] + g[k].n; }
    hreg[1].begin[0] = 0;
    hreg[1].end[0] = nt;
    if (d_reg.ensure(2)) return 2;
    UF_CUDA(cudaMemcpyAsync(d_reg.p, hreg, sizeof(hreg), cudaMemcpyHostToDevice, st));

    int64_t ns = 0;
    for (int k = 0; k < ng; k++) ns += g[k].n;
    InducePlan plan = plan_induce(nt, 1, ns, ng, sm_count(), 64);
    // large uniform-core mutual induction: symmetric pair-tile kernel (induce_sym.cuh)
    const bool sym = self_targets && uni && g[0].n >= sym_threshold();
    const int symR = sym_rows_per_thread(g[0].n);
    const int ctas = sym ? sym_ctas(symR) : 0;
    if (sym) { plan.S = ctas; }
    // few targets against many sources (update_indbound): dedicated sweep kernel
    const bool sweep = !sym && !self_targets && nt <= 256 && ns >= 2048 && sweep_supported((int)nt);
    if (sweep) plan.S = sweep_splits(ns, ng, 1024);
    const int64_t npart = sym ? sym_plane_doubles(ctas, tot) : 2 * (int64_t)plan.S * ntp;
    if (d_part.ensure(npart)) return 2;
    if (d_out.ensure(4 * ntp)) return 2;  // vx, vz, (x, z for convection)

    InduceArgs a;
    a.sx = d_src.p; a.sz = d_src.p + tot; a.sg = d_src.p + 2 * tot; a.svc4 = d_src.p + 3 * tot;
    a.tx = d_tx; a.tz = d_tz;
    a.sreg = d_reg.p; a.treg = d_reg.p + 1;
    a.nsreg = ng; a.ntreg = 1;
    a.vc4_uniform = vcu * vcu * vcu * vcu;
    a.pu = d_part.p; a.pw = d_part.p + (int64_t)plan.S * ntp;
    a.tstride = ntp;
    if (sym) {
        SymArgs sa;
        sa.x = a.sx; sa.z = a.sz; sa.g = a.sg;
        sa.reg = d_reg.p; sa.nreg = 1; sa.bv_reg = ng > 1 ? 1 : -1;
        sa.vc4 = a.vc4_uniform; sa.P = d_part.p; sa.n = tot; sa.rank = 0; sa.nranks = 1;
        if (int rc = launch_induce_sym(sa, symR, rsqrt_order(), st)) return rc;
    } else if (sweep) {
        if (int rc = launch_sweep(a, (int)nt, plan.S, uni, st)) return rc;
    } else {
        if (int rc = launch_induce_direct(a, plan, uni, rsqrt_order(), st)) return rc;
    }

    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.pu = a.pu; f.pw = a.pw; f.tstride = ntp; f.S = plan.S;
    f.treg = d_reg.p + 1; f.ntreg = 1;
    f.vx = d_out.p; f.vz = d_out.p + ntp;
    f.fs = nullptr; f.fs_u = fs_u; f.fs_w = fs_w;
    f.accumulate = accumulate;
    if (accumulate) {
        UF_CUDA(cudaMemcpyAsync(f.vx, u_inout, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
        UF_CUDA(cudaMemcpyAsync(f.vz, w_inout, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    }
    if (dt != 0.0) {
        // convect a private copy of the target positions (sources must stay untouched)
        f.x = d_out.p + 2 * ntp; f.z = d_out.p + 3 * ntp; f.dt = dt;
        UF_CUDA(cudaMemcpyAsync(f.x, d_tx, sizeof(double) * nt, cudaMemcpyDeviceToDevice, st));
        UF_CUDA(cudaMemcpyAsync(f.z, d_tz, sizeof(double) * nt, cudaMemcpyDeviceToDevice, st));
    }
    if (sym) {
        // targets = source group 0 = [0, nt) in slab indexing, so the outputs can be indexed the same way
        SymFinishArgs sf;
        memset(&sf, 0, sizeof(sf));
        sf.P = d_part.p; sf.n = tot; sf.ctas = ctas; sf.R = symR;
        sf.reg = d_reg.p; sf.nreg = 1;
        sf.vx = f.vx; sf.vz = f.vz; sf.x = f.x; sf.z = f.z;
        sf.fs_u = fs_u; sf.fs_w = fs_w; sf.dt = f.dt; sf.accumulate = accumulate;
        if (int rc = launch_sym_finish(sf, nt, st)) return rc;
    } else {
        if (int rc = launch_induce_finish(f, nt, st)) return rc;
    }
    if (nt > 0) {
        UF_CUDA(cudaMemcpyAsync(u_inout, f.vx, sizeof(double) * nt, cudaMemcpyDeviceToHost, st));
        UF_CUDA(cudaMemcpyAsync(w_inout, f.vz, sizeof(double) * nt, cudaMemcpyDeviceToHost, st));
        if (dt != 0.0 && x_out) {
            UF_CUDA(cudaMemcpyAsync(x_out, f.x, sizeof(double) * nt, cudaMemcpyDeviceToHost, st));
            UF_CUDA(cudaMemcpyAsync(z_out, f.z, sizeof(double) * nt, cudaMemcpyDeviceToHost, st));
        }
    }
    UF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// measurement kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, long long *clk) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 1.0000000001, c = 1e-12;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void k_rsqrt_probe(long long n, double lqmin, double lqstep, double *mx) {
    double e0 = 0, e2 = 0, e3 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double q = exp(lqmin + lqstep * (double)i);
        const double truth = 1.0 / sqrt(q);
        e0 = fmax(e0, fabs(rsqrt_seed(q) - truth) / truth);
        e2 = fmax(e2, fabs(rsqrt_full<2>(q) - truth) / truth);
        e3 = fmax(e3, fabs(rsqrt_full<3>(q) - truth) / truth);
    }
    atomic_max_pos(&mx[0], e0);
    atomic_max_pos(&mx[1], e2);
    atomic_max_pos(&mx[2], e3);
}

template <int ORDER>
__global__ void k_rsqrt_probe_order(long long n, double lqmin, double lqstep, double *mx) {
    double e = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double q = exp(lqmin + lqstep * (double)i);
        // truth in double-double: one Newton correction of the IEEE value with an exact residual
        const double t0 = 1.0 / sqrt(q);
        const double r = fma(-q * t0, t0, 1.0) - fma(q, t0, -(q * t0)) * t0;     // 1 - q t0^2 to ~1e-32
        const double truth_corr = 0.5 * t0 * r;                                     // truth = t0 + truth_corr
        // ORDER 7: the cubic step on a seed whose low word is donated by another register (rsqrt_seed_lo); the donor
        // here is an unrelated value with all low bits set / random
        const double y = ORDER == 7 ? rsqrt_full_lo<3>(q, __hiloint2double(0x3ff00000, (int)(i * 2654435761u) | 0x80000001), 0.375) : rsqrt_full<ORDER == 7 ? 3 : ORDER>(q);
        e = fmax(e, fabs((y - t0) - truth_corr) / t0);
    }
    atomic_max_pos(&mx[0], e);
}

}  // namespace unsflow

using namespace unsflow;

extern "C" {

const char *unsflow_last_error(void) { return g_err; }
int unsflow_version(void) { return UNSFLOW_VERSION; }

int unsflow_device_count(int *count) {
    UF_CHECK(count != nullptr, "unsflow_device_count: null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    *count = n;
    return 0;
}

int unsflow_set_device(int device) {
    if (int rc = require_device()) return rc;
    UF_CUDA(cudaSetDevice(device));
    return 0;
}

int unsflow_device_info(int *smc, int *clk_khz, int64_t *l2, int64_t *hbm) {
    if (int rc = require_device()) return rc;
    int dev = 0;
    UF_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    UF_CUDA(cudaGetDeviceProperties(&p, dev));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    if (smc) *smc = p.multiProcessorCount;
    if (clk_khz) *clk_khz = khz;
    if (l2) *l2 = p.l2CacheSize;
    if (hbm) *hbm = (int64_t)p.totalGlobalMem;
    return 0;
}

int unsflow_ind_vel(int64_t ns, const double *sx, const double *sz, const double *ss, const double *svc,
                    int64_t nt, const double *tx, const double *tz, double *u_out, double *w_out) {
    UF_CHECK(ns >= 0 && nt >= 0, "unsflow_ind_vel: negative count");
    UF_CHECK(nt == 0 || (tx && tz && u_out && w_out), "unsflow_ind_vel: null target/output pointer");
    UF_CHECK(ns == 0 || (sx && sz && ss && svc), "unsflow_ind_vel: null source pointer");
    if (int rc = require_device()) return rc;
    if (nt == 0) return 0;
    HostGroup g = {ns, sx, sz, ss, svc};
    return induce_host(&g, 1, nt, tx, tz, u_out, w_out, 0, 0.0, 0.0, 0.0, nullptr, nullptr);
}

int unsflow_mutual_ind(int64_t n, const double *x, const double *z, const double *s, const double *vc,
                       double *vx, double *vz) {
    UF_CHECK(n >= 0, "unsflow_mutual_ind: negative count");
    UF_CHECK(n == 0 || (x && z && s && vc && vx && vz), "unsflow_mutual_ind: null pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return 0;
    HostGroup g = {n, x, z, s, vc};
    return induce_host(&g, 1, n, nullptr, nullptr, vx, vz, 1, 0.0, 0.0, 0.0, nullptr, nullptr);
}

int unsflow_wakeroll(int64_t n, double *x, double *z, const double *s, const double *vc, double *vx, double *vz,
                     int64_t nbv, const double *bx, const double *bz, const double *bs, const double *bvc,
                     double u_inf, double w_inf, double dt) {
    UF_CHECK(n >= 0 && nbv >= 0, "unsflow_wakeroll: negative count");
    UF_CHECK(n == 0 || (x && z && s && vc && vx && vz), "unsflow_wakeroll: null wake pointer");
    UF_CHECK(nbv == 0 || (bx && bz && bs && bvc), "unsflow_wakeroll: null bound-vortex pointer");
    if (int rc = require_device()) return rc;
    if (n == 0) return 0;
    HostGroup g[2] = {{n, x, z, s, vc}, {nbv, bx, bz, bs, bvc}};
    // dt == 0 is legal (velocities only); induce_host convects only when dt != 0
    return induce_host(g, 2, n, nullptr, nullptr, vx, vz, 0, u_inf, w_inf, dt, x, z);
}

int unsflow_update_indbound(int64_t n, const double *x, const double *z, const double *s, const double *vc,
                            int64_t ndiv, const double *bnd_x, const double *bnd_z, double *uind, double *wind) {
    return unsflow_ind_vel(n, x, z, s, vc, ndiv, bnd_x, bnd_z, uind, wind);
}

int unsflow_set_precision(int mode) {
    UF_CHECK(mode == UNSFLOW_PRECISION_FULL || mode == UNSFLOW_PRECISION_FAST, "unsflow_set_precision: mode must be 0 (full) or 1 (fast)");
    g_order_override = mode == UNSFLOW_PRECISION_FAST ? 2 : 3;
    return 0;
}

int unsflow_release_scratch(void) {
    g_scratch.src.release(); g_scratch.t.release(); g_scratch.part.release(); g_scratch.out.release(); g_scratch.reg.release();
    return 0;
}

int unsflow_fp64_peak(double *tflops, double *sm_mhz_est) {
    UF_CHECK(tflops != nullptr, "unsflow_fp64_peak: null argument");
    if (int rc = require_device()) return rc;
    const int blocks = sm_count() * 8, threads = 256, iters = 4096;
    DevBuf<double> out;
    DevBuf<long long> clk;
    if (out.alloc((int64_t)blocks * threads) || clk.alloc(1)) return 2;
    cudaEvent_t e0, e1;
    UF_CUDA(cudaEventCreate(&e0));
    UF_CUDA(cudaEventCreate(&e1));
    double best = 0, mhz = 0;
    for (int rep = 0; rep < 5; rep++) {
        UF_CUDA(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(out.p, iters, clk.p);
        UF_CUDA(cudaEventRecord(e1));
        UF_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        UF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) {
            best = tf;
            // a resident block of 8 warps shares 4 FP64 pipes with (blocks/SM) blocks: estimate the
            // clock from the achieved rate instead: rate = SMs * 64 lanes * 2 flop * f
            mhz = tf * 1e12 / (sm_count() * 64.0 * 2.0) / 1e6;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    if (sm_mhz_est) *sm_mhz_est = mhz;
    return 0;
}

// bound-point sweep micro-benchmark: ndiv targets x n synthetic sources resident in HBM; times the
// dedicated k_sweep and the generic one-target-per-thread kernel (CUDA events, best of reps)
__global__ void k_fill_sweep(double *x, double *z, double *g, long long n, long long npad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (long long)gridDim.x * blockDim.x) {
        const bool ok = i < n;
        x[i] = ok ? 0.5 + 0.015 * (double)i : 0.0;
        z[i] = ok ? 0.3 * sin(0.37 * (double)i) : 0.0;
        g[i] = ok ? 0.01 * cos(0.11 * (double)i) : 0.0;
    }
}

int unsflow_bench_sweep(int64_t n, int ndiv, int reps, float *ms_sweep, float *ms_generic) {
    UF_CHECK(n > 0 && ndiv >= 1 && ndiv <= 256 && reps >= 1 && ms_sweep && ms_generic, "unsflow_bench_sweep: bad arguments");
    if (int rc = require_device()) return rc;
    const int64_t npad = round_up(n, kSymTB), ndp = round_up(ndiv, 8);
    DevBuf<double> src, tgt, part;
    DevBuf<Regions> reg;
    const int S = sweep_splits(n, 1, 592);
    InducePlan plan = plan_induce(ndiv, 1, n, 1, sm_count(), 592);
    const int Smax = std::max(S, plan.S);
    if (src.alloc(3 * npad) || tgt.alloc(2 * ndp) || part.alloc(2 * (int64_t)Smax * ndp) || reg.alloc(2)) return 2;
    k_fill_sweep<<<sm_count() * 4, 256>>>(src.p, src.p + npad, src.p + 2 * npad, n, npad);
    std::vector<double> ht((size_t)(2 * ndp), 0.0);
    for (int i = 0; i < ndiv; i++) { ht[i] = -1.0 + (double)i / ndiv; ht[ndp + i] = 0.02 * i / ndiv; }
    UF_CUDA(cudaMemcpy(tgt.p, ht.data(), sizeof(double) * ht.size(), cudaMemcpyHostToDevice));
    Regions hr[2];
    memset(hr, 0, sizeof(hr));
    hr[0].end[0] = n; hr[1].end[0] = ndiv;
    UF_CUDA(cudaMemcpy(reg.p, hr, sizeof(hr), cudaMemcpyHostToDevice));
    InduceArgs a;
    a.sx = src.p; a.sz = src.p + npad; a.sg = src.p + 2 * npad; a.svc4 = nullptr;
    a.tx = tgt.p; a.tz = tgt.p + ndp; a.sreg = reg.p; a.treg = reg.p + 1; a.nsreg = 1; a.ntreg = 1;
    a.vc4_uniform = 1.6e-7; a.pu = part.p; a.pw = part.p + (int64_t)Smax * ndp; a.tstride = ndp;
    cudaEvent_t e0, e1;
    UF_CUDA(cudaEventCreate(&e0));
    UF_CUDA(cudaEventCreate(&e1));
    float best[2] = {1e30f, 1e30f};
    for (int which = 0; which < 2; which++)
        for (int r = 0; r < reps + 2; r++) {
            UF_CUDA(cudaEventRecord(e0));
            if (which == 0) { if (int rc = launch_sweep(a, ndiv, S, true, 0)) return rc; }
            else { if (int rc = launch_induce_direct(a, plan, true, 3, 0)) return rc; }
            UF_CUDA(cudaEventRecord(e1));
            UF_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            UF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) best[which] = std::min(best[which], ms);
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_sweep = best[0];
    *ms_generic = best[1];
    return 0;
}

int unsflow_rsqrt_probe(int64_t nsamples, double qmin, double qmax, double *e_seed, double *e_newton, double *e_cubic) {
    UF_CHECK(nsamples > 1 && qmin > 0 && qmax > qmin, "unsflow_rsqrt_probe: bad arguments");
    if (int rc = require_device()) return rc;
    DevBuf<double> mx;
    if (mx.alloc(3)) return 2;
    UF_CUDA(cudaMemset(mx.p, 0, 3 * sizeof(double)));
    const double l0 = log(qmin), step = (log(qmax) - l0) / (double)(nsamples - 1);
    k_rsqrt_probe<<<sm_count() * 4, 256>>>(nsamples, l0, step, mx.p);
    UF_CUDA(cudaGetLastError());
    double h[3];
    UF_CUDA(cudaMemcpy(h, mx.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (e_seed) *e_seed = h[0];
    if (e_newton) *e_newton = h[1];
    if (e_cubic) *e_cubic = h[2];
    return 0;
}

int unsflow_rsqrt_probe_order(int order, int64_t nsamples, double qmin, double qmax, double *max_rel) {
    UF_CHECK(nsamples > 1 && qmin > 0 && qmax > qmin && max_rel, "unsflow_rsqrt_probe_order: bad arguments");
    UF_CHECK(order >= 2 && order <= 7, "unsflow_rsqrt_probe_order: order must be in [2, 7]");
    if (int rc = require_device()) return rc;
    DevBuf<double> mx;
    if (mx.alloc(1)) return 2;
    UF_CUDA(cudaMemset(mx.p, 0, sizeof(double)));
    const double l0 = log(qmin), step = (log(qmax) - l0) / (double)(nsamples - 1);
    const int grid = sm_count() * 4;
    switch (order) {
        case 2: k_rsqrt_probe_order<2><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
        case 3: k_rsqrt_probe_order<3><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
        case 4: k_rsqrt_probe_order<4><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
        case 5: k_rsqrt_probe_order<5><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
        case 6: k_rsqrt_probe_order<6><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
        default: k_rsqrt_probe_order<7><<<grid, 256>>>(nsamples, l0, step, mx.p); break;
    }
    UF_CUDA(cudaGetLastError());
    UF_CUDA(cudaMemcpy(max_rel, mx.p, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
