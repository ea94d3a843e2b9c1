// runtime.h -- small host-side runtime pieces shared by the .cu files.
#pragma once
#include "common.cuh"
#include <cstdlib>

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

// Page-locks caller buffers that are handed to a handle again and again (device-resident wake: upload /
// download of the same host arrays every step), so the copies run as DMA at PCIe speed instead of through the
// driver's pageable staging path.  A registration is kept until the handle dies or the range is seen with a
// different size; a failed registration is not an error (the copy falls back to the pageable path).
// UNSFLOW_NO_HOST_REGISTER=1 disables it.
struct HostPinCache {
    struct Ent { const void *p; size_t bytes; };
    Ent ent[16];
    int n = 0;
    HostPinCache() = default;
    HostPinCache(const HostPinCache &) = delete;
    HostPinCache &operator=(const HostPinCache &) = delete;
    ~HostPinCache() { clear(); }
    void clear() {
        for (int i = 0; i < n; i++) cudaHostUnregister(const_cast<void *>(ent[i].p));
        n = 0;
        (void)cudaGetLastError();
    }
    void pin(const void *p, size_t bytes) {
        static const bool off = [] { const char *e = getenv("UNSFLOW_NO_HOST_REGISTER"); return e && atoi(e) != 0; }();
        if (off || !p || bytes < (1u << 16)) return;
        const char *b = static_cast<const char *>(p), *e = b + bytes;
        for (int i = 0; i < n; i++) {
            const char *eb = static_cast<const char *>(ent[i].p), *ee = eb + ent[i].bytes;
            if (eb == b && ee == e) return;                      // already pinned
            if (b < ee && eb < e) {                              // overlaps an older registration: replace it
                cudaHostUnregister(const_cast<void *>(ent[i].p));
                ent[i] = ent[--n];
                i--;
            }
        }
        if (n == 16) { cudaHostUnregister(const_cast<void *>(ent[0].p)); ent[0] = ent[--n]; }
        if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) ent[n++] = Ent{p, bytes};
        (void)cudaGetLastError();
    }
};

// RAII CUDA event (early error returns must not leak it)
struct DevEvent {
    cudaEvent_t e = nullptr;
    DevEvent() = default;
    DevEvent(const DevEvent &) = delete;
    DevEvent &operator=(const DevEvent &) = delete;
    ~DevEvent() { if (e) cudaEventDestroy(e); }
    int create() {
        if (e) return 0;
        cudaError_t r = cudaEventCreate(&e);
        if (r != cudaSuccess) { set_error("cudaEventCreate failed: %s", cudaGetErrorString(r)); e = nullptr; return 2; }
        return 0;
    }
};

}  // namespace unsflow
