// common.cuh -- shared host/device helpers for libunsflow_cuda (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace unsflow {

// ---------------------------------------------------------------------------------------
// error plumbing: no exception crosses the C ABI; every entry returns int and keeps a
// thread-local message for unsflow_last_error().
// ---------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local char g_err[512];

#define UF_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            unsflow::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
                               __FILE__, __LINE__);                                        \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

#define UF_CHECK(cond, ...)                                                                \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            unsflow::set_error(__VA_ARGS__);                                               \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int require_device();   // 0 when a CUDA device is usable, else error (there is NO CPU fallback)

constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kInvTwoPi = 0.15915494309189533576888376337251;
constexpr double kPi = 3.141592653589793238462643383279502884;

// source tile (doubles per array per TMA bulk copy) and block size of the induction kernels
constexpr int kTS = 512;
constexpr int kThreads = 256;

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// 1/sqrt(q) in full double precision from one MUFU.RSQ64H seed.
//   ORDER 3: Halley (cubic) step  y = y0*(1 + e/2 + 3e^2/8), e = 1 - q*y0^2   -> 5 FP64 instr
//   ORDER 2: Newton (quadratic), bias-centred: y = y0*(3/2 + k - q*y0^2/2)      -> 3 FP64 instr
//            plain Newton always falls short by (3/8)e^2 in [0, 1.29e-12] (seed error 9.3e-7 measured);
//            adding the mid-range constant k = 6.4e-13 makes the error two-sided, |err| <= 6.5e-13.
//            OPT-IN fast mode only (UNSFLOW_RSQRT_ORDER=2): the default stays the cubic step.
// This replaces the reference's sqrt AND its divide (calcs.jl:472-473, :499-500).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_seed(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    return y;
}

// Same seed, but the low word of the result is taken from `donor` instead of being zeroed.  MUFU.RSQ64H only
// writes the high word; PTX's rsqrt.approx.ftz.f64 adds a MOV of 0 into the low one.  The refinement does not care
// what the low 32 mantissa bits of the seed are (they move it by < 2^-20 relative, like the seed's own error), so
// a register that dies here anyway (d^2 after q = d^4 + vc^4) donates them and the MOV disappears: the induction
// loops are issue-bound (a DFMA takes two issue slots, every other instruction one), so each one removed counts.
// Deterministic: the donor is data, not leftover register content.
__device__ __forceinline__ double rsqrt_seed_lo(double q, double donor) {
    double y;
    asm("{\n"
        ".reg .f64 t;\n"
        ".reg .b32 tlo, thi, dlo, dhi;\n"
        "rsqrt.approx.ftz.f64 t, %1;\n"
        "mov.b64 {tlo, thi}, t;\n"
        "mov.b64 {dlo, dhi}, %2;\n"
        "mov.b64 %0, {dlo, thi};\n"
        "}\n"
        : "=d"(y)
        : "d"(q), "d"(donor));
    return y;
}

// 1/sqrt(q), seed low word donated by `donor` (see rsqrt_seed_lo); k375 = 0.375 held in a (uniform) register by
// the caller so that the compiler does not rematerialise the constant inside the loop
template <int ORDER>
__device__ __forceinline__ double rsqrt_full_lo(double q, double donor, double k375) {
    const double y0 = rsqrt_seed_lo(q, donor);
    if (ORDER == 3) {
        const double s = q * y0;
        const double e = fma(-s, y0, 1.0);
        const double p = fma(k375, e, 0.5);
        const double pe = p * e;
        return fma(y0, pe, y0);
    } else {
        const int hi = (__double2hiint(y0) - 0x00100000) ^ 0x80000000;
        const double yh = __hiloint2double(hi, __double2loint(y0));
        const double s = q * y0;
        const double c = fma(s, yh, 1.5 + 6.4e-13);
        return y0 * c;
    }
}

template <int ORDER>
__device__ __forceinline__ double rsqrt_full(double q) {
    const double y0 = rsqrt_seed(q);
    if (ORDER == 2) {
        // -y0/2 by integer ops on the high word (ALU pipe, not the FP64 pipe)
        const int hi = (__double2hiint(y0) - 0x00100000) ^ 0x80000000;
        const double yh = __hiloint2double(hi, __double2loint(y0));
        const double s = q * y0;
        const double c = fma(s, yh, 1.5 + 6.4e-13);
        return y0 * c;
    } else if (ORDER == 4 || ORDER == 5 || ORDER == 6) {
        // Halley step with its second-order term in FP32: y = y0*(1 + e/2 + 3e^2/8).  |e| <= 2e-6, so
        // 3e^2/8 <= 1.5e-12 needs a relative accuracy of only ~1e-4 to stay below 1e-16 of y: e is rounded
        // to float (or its top 20 mantissa bits are re-packed as a float by integer ops), squared and
        // scaled on the FP32 pipe and widened again -- 4 FP64-pipe instructions instead of 5, same
        // 2.7e-16 accuracy class as ORDER 3 (unsflow_rsqrt_probe reports all of them).
        const double s = q * y0;
        const double e = fma(-s, y0, 1.0);
        float ef;
        if (ORDER == 4) {
            ef = __double2float_rn(e);                                   // F2F.F32.F64
        } else {
            // |e| as a float from the high word: rebias the exponent (1023 -> 127), keep 20 mantissa bits;
            // |e| < 2^-126 is clamped to 0 (its square is irrelevant)
            const int a = max(__double2hiint(e) & 0x7fffffff, 0x38100000);
            ef = __int_as_float((a - 0x38000000) << 3);
        }
        const float e2f = (ef * ef) * 0.375f;
        double e2;
        if (ORDER == 6) {
            // float -> double by integer ops (e2f >= 0; 0 maps to 2^-896, harmless)
            e2 = __hiloint2double((int)((unsigned)__float_as_int(e2f) >> 3) + 0x38000000, 0);
        } else {
            e2 = (double)e2f;                                            // F2F.F64.F32
        }
        const double pe = fma(e, 0.5, e2);
        return fma(y0, pe, y0);
    } else {
        const double s = q * y0;
        const double e = fma(-s, y0, 1.0);
        const double p = fma(0.375, e, 0.5);
        const double pe = p * e;
        return fma(y0, pe, y0);
    }
}

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // __CUDACC__

}  // namespace unsflow
