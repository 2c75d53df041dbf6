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

// SGPU_HOST convenience path, pipelined: the call is cut into chunks along time; chunk k+1 is copied
// host->device on one stream while chunk k runs on the caller's stream and chunk k-1 is copied
// device->host on a third, through two staging buffers per direction owned by the handle.  The
// per-chunk `run` is the same code the SGPU_DEVICE path executes, so streaming state (history tails,
// decimator phase, IIR state) carries across chunks exactly as it does across calls.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_in[2] = {nullptr, nullptr}, e_comp[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};
    void *d_in[2] = {nullptr, nullptr};
    void *d_out[2] = {nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
    // pinned staging for PAGEABLE caller buffers (the reference's callers hand over ordinary `&[In]` slices and get a
    // fresh `Vec<Out>`, filter/mod.rs:14): chunk k + 1 is gathered into h_in by host threads and chunk k - 1 scattered
    // from h_out while chunk k is on the PCIe bus.  Left to itself the driver stages pageable copies synchronously, one
    // direction at a time (measured 9 GB/s in total against 2 x 47 GB/s from pinned memory).
    void *h_in[2] = {nullptr, nullptr};
    void *h_out[2] = {nullptr, nullptr};
    size_t h_in_bytes = 0, h_out_bytes = 0;
    int ensure(size_t need_in, size_t need_out);
    int ensure_staging(size_t need_in, size_t need_out);
    void release();
};

// SGPU_HOST_STAGING (default below): the library's own staging of pageable caller memory
bool host_staging_enabled();
// true when the CUDA runtime knows `p` as pinned (cudaHostAlloc / cudaHostRegister) or managed memory
bool host_ptr_is_pinned(const void *p);
// rows x width bytes from (src, src_pitch) to (dst, dst_pitch) on several host threads (one below 4 MiB)
void par_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows);

// chunk length (samples per channel) for a host call of n_in samples on C channels
size_t host_chunk_len(size_t C, size_t n_in);

template <class OutLen, class Run>
int host_pipeline_body(HostPipe &hp, size_t C, const float *in, size_t n_in, size_t in_stride, float *out,
                       size_t out_stride, size_t max_out_per_in, OutLen out_len, Run run, cudaStream_t s);

// On any error the three streams are drained before the call returns: no copy into the caller's `out` (or out of
// `in`) is still in flight when the caller sees the status.
template <class OutLen, class Run>
int host_pipeline(HostPipe &hp, size_t C, const float *in, size_t n_in, size_t in_stride, float *out,
                  size_t out_stride, size_t max_out_per_in, OutLen out_len, Run run, cudaStream_t s) {
    const int st = host_pipeline_body(hp, C, in, n_in, in_stride, out, out_stride, max_out_per_in, out_len, run, s);
    if (st != SGPU_OK) {
        if (hp.s_in) cudaStreamSynchronize(hp.s_in);
        cudaStreamSynchronize(s);
        if (hp.s_out) cudaStreamSynchronize(hp.s_out);
        (void)cudaGetLastError();
    }
    return st;
}

template <class OutLen, class Run>
int host_pipeline_body(HostPipe &hp, size_t C, const float *in, size_t n_in, size_t in_stride, float *out,
                       size_t out_stride, size_t max_out_per_in, OutLen out_len, Run run, cudaStream_t s) {
    const size_t chunk = host_chunk_len(C, n_in);
    const size_t max_out = chunk * max_out_per_in + 1;
    int st = hp.ensure(C * chunk * 8, C * max_out * 8);
    if (st) return st;
    // pageable caller memory: own pinned staging + host threads (small calls stay on the driver's staged copies)
    const bool big = C * n_in * 8 >= ((size_t)4 << 20) && host_staging_enabled();
    const bool stage_in = big && !host_ptr_is_pinned(in);
    const bool stage_out = big && out && !host_ptr_is_pinned(out);
    if (stage_in || stage_out) {
        st = hp.ensure_staging(stage_in ? C * chunk * 8 : 0, stage_out ? C * max_out * 8 : 0);
        if (st) return st;
    }
    const float *src = in;
    size_t out_off = 0;
    size_t k = 0;
    int pend_b = -1;  // staged outputs of the previous chunk still to be scattered into `out`
    size_t pend_off = 0, pend_nout = 0;
    auto scatter_pending = [&]() -> int {
        if (pend_b < 0) return SGPU_OK;
        SGPU_CUDA(cudaEventSynchronize(hp.e_out[pend_b]));
        par_copy2d(out + 2 * pend_off, out_stride * 8, hp.h_out[pend_b], pend_nout * 8, pend_nout * 8, C);
        pend_b = -1;
        return SGPU_OK;
    };
    for (size_t done = 0; done < n_in; done += chunk, ++k) {
        const int b = (int)(k & 1);
        const size_t nc = n_in - done < chunk ? n_in - done : chunk;
        if (k >= 2) SGPU_CUDA(cudaStreamWaitEvent(hp.s_in, hp.e_comp[b], 0));
        if (stage_in) {
            if (k >= 2) SGPU_CUDA(cudaEventSynchronize(hp.e_in[b]));  // the copy of chunk k - 2 has left h_in[b]
            par_copy2d(hp.h_in[b], nc * 8, src + 2 * done, in_stride * 8, nc * 8, C);
            SGPU_CUDA(cudaMemcpyAsync(hp.d_in[b], hp.h_in[b], C * nc * 8, cudaMemcpyHostToDevice, hp.s_in));
        } else {
            SGPU_CUDA(cudaMemcpy2DAsync(hp.d_in[b], nc * 8, src + 2 * done, in_stride * 8, nc * 8, C,
                                        cudaMemcpyHostToDevice, hp.s_in));
        }
        SGPU_CUDA(cudaEventRecord(hp.e_in[b], hp.s_in));
        SGPU_CUDA(cudaStreamWaitEvent(s, hp.e_in[b], 0));
        if (k >= 2) SGPU_CUDA(cudaStreamWaitEvent(s, hp.e_out[b], 0));
        const size_t nout = out_len(nc);
        st = run(reinterpret_cast<const float2 *>(hp.d_in[b]), nc, (long long)nc,
                 reinterpret_cast<float2 *>(hp.d_out[b]), (long long)(nout ? nout : 1), nout, s);
        if (st) return st;
        SGPU_CUDA(cudaEventRecord(hp.e_comp[b], s));
        SGPU_CUDA(cudaStreamWaitEvent(hp.s_out, hp.e_comp[b], 0));
        if (nout) {
            if (stage_out)  // h_out[b] is free: chunk k - 2 was scattered out of it in iteration k - 1
                SGPU_CUDA(cudaMemcpyAsync(hp.h_out[b], hp.d_out[b], C * nout * 8, cudaMemcpyDeviceToHost, hp.s_out));
            else
                SGPU_CUDA(cudaMemcpy2DAsync(out + 2 * out_off, out_stride * 8, hp.d_out[b], (nout ? nout : 1) * 8, nout * 8,
                                            C, cudaMemcpyDeviceToHost, hp.s_out));
        }
        SGPU_CUDA(cudaEventRecord(hp.e_out[b], hp.s_out));
        if (stage_out) {
            st = scatter_pending();  // the previous chunk, while this one is on the bus
            if (st) return st;
            pend_b = b;
            pend_off = out_off;
            pend_nout = nout;
        }
        out_off += nout;
    }
    st = scatter_pending();
    if (st) return st;
    SGPU_CUDA(cudaStreamSynchronize(hp.s_out));
    SGPU_CUDA(cudaStreamSynchronize(s));
    return SGPU_OK;
}

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
