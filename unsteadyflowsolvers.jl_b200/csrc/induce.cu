// induce.cu -- launch plumbing for the induction kernels (see induce.cuh).
#define UNSFLOW_INDUCE_KERNELS
#include "induce_sym.cuh"
#include <algorithm>

namespace unsflow {

InducePlan plan_induce(long long nt_ub, int ntreg, long long ns_ub, int nsreg, int sm_count, int max_S) {
    InducePlan p;
    const long long slots = (long long)sm_count * 2;
    const long long tsrc = std::max<long long>(1, (ns_ub + kTS - 1) / kTS + nsreg);
    auto tiles = [&](int R) { return (nt_ub + (long long)kThreads * R - 1) / ((long long)kThreads * R) + ntreg; };
    int R = 1;
    if (tiles(4) * tsrc >= 4 * slots) R = 4;
    else if (tiles(2) * tsrc >= 4 * slots) R = 2;
    p.R = R;
    p.ntiles = (int)std::max<long long>(1, tiles(R));
    long long S = (8 * slots + p.ntiles - 1) / p.ntiles;
    S = std::min<long long>(S, tsrc);
    S = std::min<long long>(S, max_S);
    S = std::min<long long>(S, 65535);
    p.S = (int)std::max<long long>(1, S);
    return p;
}

template <int R, bool U, int O>
static void launch_one(const InduceArgs &a, const InducePlan &p, cudaStream_t st) {
    dim3 grid(p.ntiles, p.S);
    k_induce_direct<R, U, O><<<grid, kThreads, 0, st>>>(a);
}

int launch_induce_direct(const InduceArgs &a, const InducePlan &p, bool uniform, int order, cudaStream_t st) {
#define UF_CASE(R_)                                                                   \
    case R_:                                                                          \
        if (uniform) { if (order == 2) launch_one<R_, true, 2>(a, p, st); else launch_one<R_, true, 3>(a, p, st); } \
        else         { if (order == 2) launch_one<R_, false, 2>(a, p, st); else launch_one<R_, false, 3>(a, p, st); } \
        break;
    switch (p.R) {
        UF_CASE(1)
        UF_CASE(2)
        UF_CASE(4)
        default: set_error("launch_induce_direct: unsupported R=%d", p.R); return 1;
    }
#undef UF_CASE
    UF_CUDA(cudaGetLastError());
    return 0;
}

int launch_induce_finish(const FinishArgs &a, long long nt_ub, cudaStream_t st) {
    if (nt_ub <= 0) return 0;
    const int blocks = (int)((nt_ub + 255) / 256);
    k_induce_finish<<<blocks, 256, 0, st>>>(a);
    UF_CUDA(cudaGetLastError());
    return 0;
}

size_t sym_smem_bytes() { return sizeof(double) * (2 * 3 * kSymTB + 2 * kSymTB) + 2 * sizeof(uint64_t) + 16; }

int launch_induce_sym(const SymArgs &a, int ctas, int order, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sym_smem_bytes();
    if (!attr_set) {
        UF_CUDA(cudaFuncSetAttribute(k_induce_sym<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UF_CUDA(cudaFuncSetAttribute(k_induce_sym<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (order == 2) k_induce_sym<2><<<ctas, kThreads, smem, st>>>(a);
    else k_induce_sym<3><<<ctas, kThreads, smem, st>>>(a);
    UF_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace unsflow
