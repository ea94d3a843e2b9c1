// FP64 issue probe (sm_100a): achieved DFMA rate per SM as a function of resident warps per scheduler and
// independent chains per thread.  Answers: how many warps / chains does the FP64 pipe need to stay busy?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_issue_probe fp64_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_chain(double *out, int iters, double m, double c) {
    double a[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) a[k] = 1.0 + threadIdx.x * 1e-9 + k * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < CH; k++) a[k] = fma(a[k], m, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 chains; every NF DFMAs per chain-set one MUFU.RSQ64H (the seed instruction of the induction kernels) or one
// F2F.F32.F64 + F2F.F64.F32 pair is issued: what does it cost in FP64-pipe time?
template <int KIND, int NF>
__global__ void k_mix(double *out, int iters, double m, double c) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + threadIdx.x * 1e-9 + k * 1e-3;
    double acc = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < NF; k++) a[k % 8] = fma(a[k % 8], m, c);
            if (KIND == 1) {
                double y;
                asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[r]));
                acc += __hiloint2double(0, __double2hiint(y)) == 12345.0 ? 1.0 : 0.0;   // keeps the MUFU alive at ~no FP64 cost? (1 DSETP) -- see KIND 3
            } else if (KIND == 2) {
                float f;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(a[r]));
                double y;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(f));
                acc += __hiloint2double(0, __double2hiint(y)) == 12345.0 ? 1.0 : 0.0;
            } else if (KIND == 3) {
                acc += __hiloint2double(0, __double2hiint(a[r])) == 12345.0 ? 1.0 : 0.0;       // the bookkeeping alone
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int NF>
static double run_mix(int warps_per_sm, int sms) {
    const int threads = warps_per_sm * 32, blocks = sms, iters = 2000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_mix<KIND, NF><<<blocks, threads>>>(out, iters, 1.0000000001, 1e-12);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaFree(out);
    // SM cycles per group of (NF DFMA + 1 extra) per scheduler, assuming 1.92 GHz
    const double groups = 8.0 * iters * (warps_per_sm / 4.0);
    return best * 1e-3 * 1.92e9 / groups;
}

template <int CH>
static double run(int warps_per_sm, int sms) {
    // one CTA per SM with warps_per_sm warps (<= 32), or several CTAs when more
    const int threads = warps_per_sm <= 32 ? warps_per_sm * 32 : 1024;
    const int blocks = sms * (warps_per_sm <= 32 ? 1 : warps_per_sm / 32);
    const int iters = 20000 / CH;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_chain<CH><<<blocks, threads>>>(out, iters, 1.0000000001, 1e-12);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaFree(out);
    const double fl = 2.0 * 8 * CH * (double)iters * blocks * threads;
    return fl / (best * 1e-3) / 1e12;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("# DFMA TFLOP/s (2 flop per DFMA) on %d SMs: rows = warps per SM (4 schedulers), columns = independent chains per thread\n", sms);
    printf("| warps/SM | 1 | 2 | 4 | 8 | 16 |\n|---|---|---|---|---|---|\n");
    const int ws[] = {4, 8, 12, 16, 24, 32, 64};
    for (int w : ws)
        printf("| %d | %.2f | %.2f | %.2f | %.2f | %.2f |\n", w, run<1>(w, sms), run<2>(w, sms), run<4>(w, sms), run<8>(w, sms), run<16>(w, sms));
    printf("\n# cycles (at 1.92 GHz) per scheduler for one group of NF DFMAs (+ the extra instruction), 8 warps/SM, 8 chains\n");
    printf("| NF | DFMA only + bookkeeping | + MUFU.RSQ64H | + F2F.F32.F64 and F2F.F64.F32 |\n|---|---|---|---|\n");
    printf("| 13 | %.2f | %.2f | %.2f |\n", run_mix<3, 13>(8, sms), run_mix<1, 13>(8, sms), run_mix<2, 13>(8, sms));
    printf("| 16 | %.2f | %.2f | %.2f |\n", run_mix<3, 16>(8, sms), run_mix<1, 16>(8, sms), run_mix<2, 16>(8, sms));
    printf("| 32 | %.2f | %.2f | %.2f |\n", run_mix<3, 32>(8, sms), run_mix<1, 32>(8, sms), run_mix<2, 32>(8, sms));
    return 0;
}
