// Roofline denominators measured on the box: FP32 FMA issue peak (scalar FFMA, packed FFMA2)
// and a streaming copy.  MEASURED_PEAKS.json only records HBM and bf16 tensor peaks; four of
// the five BASELINE configs are bound by the FP32 FMA pipe, so bench.py measures that here.
// Separate from libsolid_gpu.so on purpose: this is measurement tooling, not product API.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#define PB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

constexpr int kAcc = 16;     // independent accumulators per thread (ILP)
constexpr int kInner = 64;   // FMA "rounds" per loop iteration

// variant 0: scalar FFMA, three register operands, 2*kAcc independent chains
__global__ void __launch_bounds__(256) fma_scalar_kernel(float *out, int iters, float a0, float b0) {
    float acc[2 * kAcc];
#pragma unroll
    for (int i = 0; i < 2 * kAcc; ++i) acc[i] = threadIdx.x * 1e-6f + i;
    float a = a0 + threadIdx.x * 1e-9f, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < 2 * kAcc; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * kAcc; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

// variant 1: packed fma.rn.f32x2 (FFMA2), kAcc independent register-pair chains
__global__ void __launch_bounds__(256) fma_packed_kernel(float *out, int iters, float a0, float b0) {
    float2 acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = make_float2(threadIdx.x * 1e-6f + i, i * 0.5f);
    const float2 a = make_float2(a0 + threadIdx.x * 1e-9f, a0);
    const float2 b = make_float2(b0, b0 * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) out[0] = s;
}

// variant 2: FIR-shaped FFMA2 -- acc[i] += w[j] * g with the accumulator as the addend (the
// operand pattern of fir_core: two distinct register-pair multiplicands plus the accumulator)
__global__ void __launch_bounds__(256) fma_packed_fir_kernel(float *out, int iters, float a0, float b0) {
    float2 acc[kAcc];
    float2 w[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
        acc[i] = make_float2(0.f, 0.f);
        w[i] = make_float2(a0 * (i + 1) + threadIdx.x * 1e-9f, b0 * (i + 2));
    }
    float2 g = make_float2(a0, a0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) acc[i] = __ffma2_rn(w[(i + k) % kAcc], g, acc[i]);
        }
        g.x += 1e-7f;
        g.y = g.x;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) out[0] = s;
}

// variant 3: same FIR-shaped pattern with scalar FFMA
__global__ void __launch_bounds__(256) fma_scalar_fir_kernel(float *out, int iters, float a0, float b0) {
    float2 acc[kAcc];
    float2 w[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
        acc[i] = make_float2(0.f, 0.f);
        w[i] = make_float2(a0 * (i + 1) + threadIdx.x * 1e-9f, b0 * (i + 2));
    }
    float g = a0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) {
                acc[i].x = fmaf(w[(i + k) % kAcc].x, g, acc[i].x);
                acc[i].y = fmaf(w[(i + k) % kAcc].y, g, acc[i].y);
            }
        }
        g += 1e-7f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const float4 *__restrict__ in, float4 *__restrict__ out,
                                                   size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        out[i] = a;
        out[i + stride] = b;
        out[i + 2 * stride] = c;
        out[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) out[i] = in[i];
}

}  // namespace

// Runs `reps` timed launches (after one warm-up) of FMA variant `variant` with
// blocks_per_sm * SMs blocks of 256 threads; returns the best time and the FP32 flop count of
// one launch (2 flops per FMA lane-op).
PB_EXPORT int sgpu_peak_fma_ex(int variant, int blocks_per_sm, int threads, int iters, int reps, double *best_ms,
                               double *flops_per_launch);

PB_EXPORT int sgpu_peak_fma(int variant, int blocks_per_sm, int iters, int reps, double *best_ms,
                            double *flops_per_launch) {
    return sgpu_peak_fma_ex(variant, blocks_per_sm, 256, iters, reps, best_ms, flops_per_launch);
}

// same with an explicit block size (<= 256): occupancy sweeps (how many warps per SM sub-partition the
// FMA pipe needs before it saturates)
PB_EXPORT int sgpu_peak_fma_ex(int variant, int blocks_per_sm, int threads, int iters, int reps, double *best_ms,
                               double *flops_per_launch) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1;
    const int blocks = p.multiProcessorCount * blocks_per_sm;
    float *out = nullptr;
    if (cudaMalloc(&out, 256) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r <= reps; ++r) {
        cudaEventRecord(e0);
        switch (variant) {
            case 0: fma_scalar_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); break;
            case 1: fma_packed_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); break;
            case 2: fma_packed_fir_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); break;
            case 3: fma_scalar_fir_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); break;
            default: cudaFree(out); return -2;
        }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -3; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    // lane-FMAs per thread per iteration: 2*kAcc*kInner for every variant (packed = 2 lanes/instr)
    const double fmas = (double)blocks * (double)threads * (double)iters * (2.0 * kAcc * kInner);
    if (best_ms) *best_ms = best;
    if (flops_per_launch) *flops_per_launch = 2.0 * fmas;
    return 0;
}

// Streaming copy of `bytes` (read + write counted), best of reps.
PB_EXPORT int sgpu_peak_copy(size_t bytes, int reps, double *best_ms, double *bytes_moved) {
    float4 *a = nullptr, *b = nullptr;
    const size_t n4 = bytes / 16;
    if (cudaMalloc(&a, n4 * 16) != cudaSuccess) return -1;
    if (cudaMalloc(&b, n4 * 16) != cudaSuccess) { cudaFree(a); return -1; }
    cudaMemset(a, 1, n4 * 16);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r <= reps; ++r) {
        cudaEventRecord(e0);
        copy_kernel<<<p.multiProcessorCount * 16, 256>>>(a, b, n4);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(a); cudaFree(b); return -3; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    if (best_ms) *best_ms = best;
    if (bytes_moved) *bytes_moved = 2.0 * (double)n4 * 16.0;
    return 0;
}
