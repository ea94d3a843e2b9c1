// wake.cu -- device-resident free-vortex wake: repeated wakeroll steps (calcs.jl:222-294)
// without leaving HBM, optionally target-sliced over ranks (one process per GPU) with an NCCL
// all-gather of the convected positions each step.
#include "../../include/unsflow.h"
#include "induce_sym.cuh"
#include "nccl_dl.h"
#include "runtime.h"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <vector>

namespace unsflow {

const NcclApi *nccl_api() {
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 failed
    if (state == 1) return &api;
    if (state == -1) { set_error("libnccl.so.2 could not be loaded"); return nullptr; }
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { state = -1; set_error("dlopen(libnccl.so.2) failed: %s", dlerror()); return nullptr; }
#define UF_SYM(field, name)                                                    \
    *(void **)(&api.field) = dlsym(h, name);                                   \
    if (!api.field) { state = -1; set_error("dlsym(%s) failed", name); return nullptr; }
    UF_SYM(GetUniqueId, "ncclGetUniqueId")
    UF_SYM(CommInitRank, "ncclCommInitRank")
    UF_SYM(CommDestroy, "ncclCommDestroy")
    UF_SYM(AllGather, "ncclAllGather")
    UF_SYM(AllReduce, "ncclAllReduce")
    UF_SYM(GroupStart, "ncclGroupStart")
    UF_SYM(GroupEnd, "ncclGroupEnd")
    UF_SYM(GetErrorString, "ncclGetErrorString")
#undef UF_SYM
    state = 1;
    return &api;
}

#define UF_NCCL(call)                                                                         \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            set_error("%s failed: %s", #call, nccl_api() ? nccl_api()->GetErrorString(r_) : "?"); \
            return 4;                                                                         \
        }                                                                                     \
    } while (0)

}  // namespace unsflow

using namespace unsflow;

struct unsflow_wake {
    int64_t n = 0, nbv = 0;
    int64_t slice = 0;     // targets per rank (equal on all ranks; last slice may be partly padding)
    int64_t npad = 0;      // wake slab length (>= slice*nranks, multiple of kTS)
    int64_t bvpad = 0;     // bound-vortex slab length
    int64_t tot = 0;       // npad + bvpad
    int rank = 0, nranks = 1, variant = UNSFLOW_KERNEL_DIRECT;
    bool uniform = true;
    double vc4u = 1.0;
    bool have_vc = false, have_bvc = false, uni_w = true, uni_b = true;
    double vc0_w = 0.0, vc0_b = 0.0;   // first core radius of the wake / bound slab (the uniform value when uni_*)
    DevBuf<int> flag;      // device flag of the uniformity test
    HostPinCache pins;     // page-locked caller buffers (upload / download every step)
    DevBuf<double> src;    // [x | z | g | vc4], each `tot`
    DevBuf<double> vel;    // [vx | vz], each npad
    DevBuf<double> part;   // [pu | pw], each S*npad
    DevBuf<double> planes; // symmetric kernel: [ctas][2][tot] + touch records
    DevBuf<double> vsum;   // symmetric kernel on several ranks: packed plane sums [u(n) | w(n)] for the all-reduce
    int ctas = 0, symR = 4;
    bool want_sym = false;
    DevBuf<Regions> reg;   // [0] sources, [1] my targets
    InducePlan plan{};
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> kev;  // pairs around the induction launches (first kMaxTimed steps)
    int ktimed = 0;
    int64_t launches = 0;
    ncclComm_t comm = nullptr;
    static constexpr int kMaxTimed = 64;
};

namespace unsflow {
// v[i] <- v[i]^4 with the host's operation order ((vc*vc)*vc)*vc (multiplies only: nothing to contract);
// *flag is cleared when an entry differs from v0
__global__ void k_vc4_inplace(double *v, long long n, double v0, int *flag) {
    bool same = true;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double c = v[i];
        same = same && c == v0;
        v[i] = __dmul_rn(__dmul_rn(__dmul_rn(c, c), c), c);
    }
    if (!same) *flag = 0;
}
__global__ void k_fill(double *v, long long n, double val) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = val;
}
}  // namespace unsflow

// target slice of the direct kernel: equal slices (NCCL all-gather needs equal counts)
static int64_t slice_len(int64_t n, int nranks) { return round_up((n + nranks - 1) / nranks, 2); }

extern "C" {

int unsflow_partition(int64_t n, int nranks, int rank, int variant, int64_t *begin, int64_t *end) {
    UF_CHECK(n > 0 && nranks >= 1 && rank >= 0 && rank < nranks && begin && end, "unsflow_partition: bad arguments");
    if (variant == UNSFLOW_KERNEL_SYMMETRIC) {
        const long long tb = (long long)kThreads * sym_rows_per_thread(n, nranks);
        const long long nb = (n + tb - 1) / tb, W = nb * (nb + 1) / 2;
        long long b, e;
        split_range(W * kSymSub, rank, nranks, &b, &e);   // in sub-units of 1/kSymSub pair tile
        *begin = b; *end = e;
    } else {
        const int64_t sl = slice_len(n, nranks);
        *begin = std::min<int64_t>(n, rank * sl);
        *end = std::min<int64_t>(n, (rank + 1) * sl);
    }
    return 0;
}

int unsflow_nccl_unique_id(void *id128) {
    UF_CHECK(id128 != nullptr, "unsflow_nccl_unique_id: null argument");
    const NcclApi *nc = nccl_api();
    if (!nc) return 4;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    UF_NCCL(nc->GetUniqueId((ncclUniqueId *)id128));
    return 0;
}

int unsflow_wake_create(unsflow_wake **out, int64_t n, int64_t nbv, int rank, int nranks, const void *uid, int variant) {
    UF_CHECK(out != nullptr, "unsflow_wake_create: null handle pointer");
    *out = nullptr;
    UF_CHECK(n > 0 && nbv >= 0, "unsflow_wake_create: need n > 0, nbv >= 0");
    UF_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "unsflow_wake_create: bad rank %d / %d", rank, nranks);
    UF_CHECK(variant >= 0 && variant <= 2, "unsflow_wake_create: bad kernel variant %d", variant);
    if (int rc = require_device()) return rc;
    unsflow_wake *w = new (std::nothrow) unsflow_wake();
    UF_CHECK(w != nullptr, "out of host memory");
    w->n = n; w->nbv = nbv; w->rank = rank; w->nranks = nranks;
    w->variant = variant;
    w->slice = slice_len(n, nranks);
    w->npad = round_up(w->slice * nranks, kSymTB);
    w->bvpad = round_up(std::max<int64_t>(nbv, 1), kTS);
    w->tot = w->npad + w->bvpad;
    auto fail = [&](int rc) { unsflow_wake_destroy(w); return rc; };
    if (cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(2); }
    cudaEventCreate(&w->ev0);
    cudaEventCreate(&w->ev1);
    w->kev.resize(2 * unsflow_wake::kMaxTimed);
    for (auto &e : w->kev) cudaEventCreate(&e);
    if (w->src.alloc(4 * w->tot) || w->vel.alloc(2 * w->npad) || w->reg.alloc(2) || w->flag.alloc(1)) return fail(2);
    const int64_t my_t = std::max<int64_t>(0, std::min(n, (rank + 1) * w->slice) - rank * w->slice);
    w->plan = plan_induce(my_t, 1, n + nbv, 2, sm_count(), 64);
    if (w->part.alloc(2 * (int64_t)w->plan.S * w->npad)) return fail(2);
    w->want_sym = variant == UNSFLOW_KERNEL_SYMMETRIC || (variant == UNSFLOW_KERNEL_AUTO && n >= sym_threshold());
    w->symR = sym_rows_per_thread(n, nranks);
    w->ctas = sym_ctas(w->symR);
    if (w->want_sym && w->planes.alloc(sym_plane_doubles(w->ctas, w->tot))) return fail(2);
    if (w->want_sym && nranks > 1 && w->vsum.alloc(2 * n)) return fail(2);
    if (w->src.zero(w->st) || w->vel.zero(w->st)) return fail(2);
    // dead / padding entries: strength 0, finite position, vc^4 = 1 > 0
    k_fill<<<sm_count() * 2, 256, 0, w->st>>>(w->src.p + 3 * w->tot, w->tot, 1.0);
    Regions h[2];
    memset(h, 0, sizeof(h));
    h[0].begin[0] = 0; h[0].end[0] = n;
    h[0].begin[1] = w->npad; h[0].end[1] = w->npad + nbv;
    h[1].begin[0] = rank * w->slice; h[1].end[0] = rank * w->slice + my_t;
    if (cudaMemcpyAsync(w->reg.p, h, sizeof(h), cudaMemcpyHostToDevice, w->st) != cudaSuccess) { set_error("region upload failed"); return fail(2); }
    if (nranks > 1) {
        if (uid == nullptr) { set_error("unsflow_wake_create: nccl_unique_id required when nranks > 1"); return fail(1); }
        const NcclApi *nc = nccl_api();
        if (!nc) return fail(4);
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        ncclResult_t r = nc->CommInitRank(&w->comm, nranks, id, rank);
        if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", nc->GetErrorString(r)); return fail(4); }
    }
    if (cudaStreamSynchronize(w->st) != cudaSuccess) { set_error("wake_create: sync failed"); return fail(2); }
    *out = w;
    return 0;
}

int unsflow_wake_upload(unsflow_wake *w, const double *x, const double *z, const double *s, const double *vc,
                        const double *bx, const double *bz, const double *bs, const double *bvc) {
    UF_CHECK(w != nullptr, "unsflow_wake_upload: null handle");
    double *base = w->src.p;
    const int64_t tot = w->tot;
    auto up = [&](int arr, int64_t off, const double *h, int64_t cnt) -> int {
        if (!h || cnt == 0) return 0;
        w->pins.pin(h, sizeof(double) * cnt);
        UF_CUDA(cudaMemcpyAsync(base + arr * tot + off, h, sizeof(double) * cnt, cudaMemcpyHostToDevice, w->st));
        return 0;
    };
    if (up(0, 0, x, w->n) || up(1, 0, z, w->n) || up(2, 0, s, w->n)) return 2;
    if (up(0, w->npad, bx, w->nbv) || up(1, w->npad, bz, w->nbv) || up(2, w->npad, bs, w->nbv)) return 2;
    // core radii: only the sub-range that was passed is rewritten; vc -> vc^4 and the uniformity test run on the device
    auto up_vc = [&](int64_t off, const double *h, int64_t cnt, bool *have, bool *uni, double *v0) -> int {
        if (!h || cnt == 0) return 0;
        if (up(3, off, h, cnt)) return 2;
        int one = 1;
        UF_CUDA(cudaMemcpyAsync(w->flag.p, &one, sizeof(int), cudaMemcpyHostToDevice, w->st));
        k_vc4_inplace<<<sm_count() * 2, 256, 0, w->st>>>(base + 3 * tot + off, cnt, h[0], w->flag.p);
        UF_CUDA(cudaGetLastError());
        int f = 0;
        UF_CUDA(cudaMemcpyAsync(&f, w->flag.p, sizeof(int), cudaMemcpyDeviceToHost, w->st));
        UF_CUDA(cudaStreamSynchronize(w->st));
        *have = true; *uni = f != 0; *v0 = h[0];
        return 0;
    };
    if (up_vc(0, vc, w->n, &w->have_vc, &w->uni_w, &w->vc0_w)) return 2;
    if (up_vc(w->npad, bvc, w->nbv, &w->have_bvc, &w->uni_b, &w->vc0_b)) return 2;
    w->uniform = w->uni_w && (w->nbv == 0 || (w->uni_b && (!w->have_vc || !w->have_bvc || w->vc0_b == w->vc0_w)));
    const double v0 = w->have_vc ? w->vc0_w : w->vc0_b;
    w->vc4u = v0 * v0 * v0 * v0;
    UF_CUDA(cudaStreamSynchronize(w->st));
    return 0;
}

int unsflow_wake_download(unsflow_wake *w, double *x, double *z, double *vx, double *vz) {
    UF_CHECK(w != nullptr, "unsflow_wake_download: null handle");
    const int64_t n = w->n;
    auto down = [&](double *h, const double *d) -> int {
        if (!h) return 0;
        w->pins.pin(h, sizeof(double) * n);
        UF_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * n, cudaMemcpyDeviceToHost, w->st));
        return 0;
    };
    // velocities exist only for this rank's target slice (direct variant); other entries are whatever was there
    if (down(x, w->src.p) || down(z, w->src.p + w->tot) || down(vx, w->vel.p) || down(vz, w->vel.p + w->npad)) return 2;
    UF_CUDA(cudaStreamSynchronize(w->st));
    return 0;
}

int unsflow_wake_step(unsflow_wake *w, int64_t nsteps, double u_inf, double w_inf, double dt) {
    UF_CHECK(w != nullptr, "unsflow_wake_step: null handle");
    UF_CHECK(nsteps >= 0, "unsflow_wake_step: negative nsteps");
    UF_CHECK(w->have_vc && (w->nbv == 0 || w->have_bvc), "unsflow_wake_step: core radii were never uploaded (pass vc%s to unsflow_wake_upload first)", w->nbv ? " and bv_vc" : "");
    const int64_t tot = w->tot, npad = w->npad;
    InduceArgs a;
    a.sx = w->src.p; a.sz = w->src.p + tot; a.sg = w->src.p + 2 * tot; a.svc4 = w->src.p + 3 * tot;
    a.tx = a.sx; a.tz = a.sz;
    a.sreg = w->reg.p; a.treg = w->reg.p + 1;
    a.nsreg = 2; a.ntreg = 1;
    a.vc4_uniform = w->vc4u;
    a.pu = w->part.p; a.pw = w->part.p + (int64_t)w->plan.S * npad;
    a.tstride = npad;
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.pu = a.pu; f.pw = a.pw; f.tstride = npad; f.S = w->plan.S;
    f.treg = w->reg.p + 1; f.ntreg = 1;
    f.vx = w->vel.p; f.vz = w->vel.p + npad;
    f.x = w->src.p; f.z = w->src.p + tot;
    f.fs_u = u_inf; f.fs_w = w_inf; f.dt = dt;
    const int64_t my_t = std::max<int64_t>(0, std::min(w->n, (w->rank + 1) * w->slice) - w->rank * w->slice);
    const bool sym = w->want_sym && w->uniform;
    UF_CHECK(sym || w->variant != UNSFLOW_KERNEL_SYMMETRIC, "unsflow_wake_step: the symmetric kernel needs a uniform core radius");
    SymArgs sa;
    sa.x = a.sx; sa.z = a.sz; sa.g = a.sg;
    sa.reg = w->reg.p; sa.nreg = 1; sa.bv_reg = w->nbv > 0 ? 1 : -1;
    sa.vc4 = w->vc4u;
    sa.P = w->planes.p; sa.n = tot;
    sa.rank = w->rank; sa.nranks = w->nranks;
    w->ktimed = 0;
    w->launches = 0;
    UF_CUDA(cudaEventRecord(w->ev0, w->st));
    for (int64_t is = 0; is < nsteps; is++) {
        const bool timed = is < unsflow_wake::kMaxTimed;
        if (timed) UF_CUDA(cudaEventRecord(w->kev[2 * is], w->st));
        if (sym) {
            // every rank evaluates its share of the unordered pair tiles for ALL vortices
            if (int rc = launch_induce_sym(sa, w->symR, rsqrt_order(), w->st)) return rc;
            if (timed) { UF_CUDA(cudaEventRecord(w->kev[2 * is + 1], w->st)); w->ktimed = (int)is + 1; }
            SymFinishArgs fs;
            memset(&fs, 0, sizeof(fs));
            fs.P = w->planes.p; fs.n = tot; fs.ctas = w->ctas; fs.R = w->symR;
            fs.reg = w->reg.p; fs.nreg = 1;          // all n free vortices
            fs.vx = w->vel.p; fs.vz = w->vel.p + npad; fs.x = w->src.p; fs.z = w->src.p + tot;
            fs.fs_u = u_inf; fs.fs_w = w_inf; fs.dt = dt;
            if (w->nranks == 1) {
                if (int rc = launch_sym_finish(fs, w->n, w->st)) return rc;
                w->launches += 2;
            } else {
                // plain sums of this rank's planes, packed [u | w] -> ONE all-reduce -> 1/2pi, free stream, convection
                fs.raw = 1; fs.packed = 1; fs.packed_stride = w->n; fs.vx = w->vsum.p;
                if (int rc = launch_sym_finish(fs, w->n, w->st)) return rc;
                const NcclApi *nc = nccl_api();
                if (!nc) return 4;
                UF_NCCL(nc->AllReduce(w->vsum.p, w->vsum.p, (size_t)(2 * w->n), ncclDouble, ncclSum, w->comm, w->st));
                FinishArgs f2 = f;
                f2.pu = w->vsum.p; f2.pw = w->vsum.p + w->n; f2.tstride = 0; f2.S = 1; f2.packed = 1;
                f2.treg = w->reg.p; f2.ntreg = 1;
                if (int rc = launch_induce_finish(f2, w->n, w->st)) return rc;
                w->launches += 3;
            }
            continue;
        }
        if (int rc = launch_induce_direct(a, w->plan, w->uniform, rsqrt_order(), w->st)) return rc;
        if (timed) { UF_CUDA(cudaEventRecord(w->kev[2 * is + 1], w->st)); w->ktimed = (int)is + 1; }
        if (int rc = launch_induce_finish(f, my_t, w->st)) return rc;
        w->launches += 2;
        if (w->nranks > 1) {
            const NcclApi *nc = nccl_api();
            if (!nc) return 4;
            double *x = w->src.p, *z = w->src.p + tot;
            UF_NCCL(nc->GroupStart());
            UF_NCCL(nc->AllGather(x + w->rank * w->slice, x, (size_t)w->slice, ncclDouble, w->comm, w->st));
            UF_NCCL(nc->AllGather(z + w->rank * w->slice, z, (size_t)w->slice, ncclDouble, w->comm, w->st));
            UF_NCCL(nc->GroupEnd());
        }
    }
    UF_CUDA(cudaEventRecord(w->ev1, w->st));
    return 0;
}

int unsflow_wake_sync(unsflow_wake *w) {
    UF_CHECK(w != nullptr, "unsflow_wake_sync: null handle");
    UF_CUDA(cudaStreamSynchronize(w->st));
    return 0;
}

int unsflow_wake_timing(unsflow_wake *w, float *ms_total, float *ms_induce, int64_t *n_launches) {
    UF_CHECK(w != nullptr, "unsflow_wake_timing: null handle");
    UF_CUDA(cudaStreamSynchronize(w->st));
    float t = 0;
    UF_CUDA(cudaEventElapsedTime(&t, w->ev0, w->ev1));
    if (ms_total) *ms_total = t;
    float k = 0;
    for (int i = 0; i < w->ktimed; i++) {
        float d = 0;
        UF_CUDA(cudaEventElapsedTime(&d, w->kev[2 * i], w->kev[2 * i + 1]));
        k += d;
    }
    if (ms_induce) *ms_induce = k;
    if (n_launches) *n_launches = w->launches;
    return 0;
}

int unsflow_wake_destroy(unsflow_wake *w) {
    if (!w) return 0;
    if (w->comm && nccl_api()) nccl_api()->CommDestroy(w->comm);
    if (w->st) cudaStreamSynchronize(w->st);
    for (auto &e : w->kev) if (e) cudaEventDestroy(e);
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
    return 0;
}

}  // extern "C"
