// Shared host/device plumbing for libsolid_gpu.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/solid_gpu.h"

#define SGPU_EXPORT extern "C" __attribute__((visibility("default")))

namespace sgpu {

// thread-local last error text (sgpu_last_error)
char *err_buf();
int fail(int status, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SGPU_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::sgpu::fail(SGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define SGPU_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return ::sgpu::fail(SGPU_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                                cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

// Requires a Blackwell-class device; there is no CPU fallback anywhere in this library.
int require_device(int *device_out, int *sm_count_out);

// Growable device scratch owned by a handle, used by the SGPU_HOST convenience path.
struct Staging {
    void *in = nullptr, *out = nullptr;
    size_t in_bytes = 0, out_bytes = 0;
    int ensure(size_t need_in, size_t need_out);
    void release();
};

// RAII guard: make the handle's device current for the duration of a call.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            switched = true;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace sgpu
