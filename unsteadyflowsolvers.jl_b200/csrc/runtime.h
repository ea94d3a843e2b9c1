// runtime.h -- small host-side runtime pieces shared by the .cu files.
#pragma once
#include "common.cuh"

namespace unsflow {

int rsqrt_order();   // 3 (cubic, default) or 2 (Newton) -- env UNSFLOW_RSQRT_ORDER
int sm_count();

// RAII device buffer (no exceptions: alloc returns non-zero and sets the error message)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    int64_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // grow-only variant for scratch that is reused across calls (keeps the allocation when it is
    // already large enough)
    int ensure(int64_t count) {
        if (p && n >= count) return 0;
        return alloc(count);
    }
    int alloc(int64_t count) {
        release();
        if (count <= 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, sizeof(T) * (size_t)count);
        if (e != cudaSuccess) {
            set_error("cudaMalloc of %lld bytes failed: %s", (long long)(sizeof(T) * (size_t)count),
                      cudaGetErrorString(e));
            p = nullptr;
            return 2;
        }
        n = count;
        return 0;
    }
    int zero(cudaStream_t st = 0) {
        cudaError_t e = cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); return 2; }
        return 0;
    }
};

}  // namespace unsflow
