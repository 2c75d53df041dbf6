// DotProduct<T> + Execute::execute (dot_product/mod.rs:37-87,153-171) on sm_100a.
// sum_{i < min(len_c, len_x)} c[i] * x[i]; one block per sample vector, warp-shuffle reduce.
#include "sgpu_common.cuh"

using namespace sgpu;

namespace {

template <bool COMPLEX>
__global__ void __launch_bounds__(256) dot_kernel(const float *__restrict__ coefs, int n_terms,
                                                  const float2 *__restrict__ x, long long x_stride,
                                                  float2 *__restrict__ result) {
    const float2 *v = x + (long long)blockIdx.x * x_stride;
    float2 acc = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < n_terms; i += blockDim.x) {
        const float2 s = v[i];
        if constexpr (COMPLEX) {
            const float2 c = reinterpret_cast<const float2 *>(coefs)[i];
            acc.x = fmaf(c.x, s.x, acc.x);
            acc.x = fmaf(-c.y, s.y, acc.x);
            acc.y = fmaf(c.x, s.y, acc.y);
            acc.y = fmaf(c.y, s.x, acc.y);
        } else {
            const float c = coefs[i];
            acc.x = fmaf(c, s.x, acc.x);
            acc.y = fmaf(c, s.y, acc.y);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    __shared__ float2 part[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) part[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = l < (int)(blockDim.x >> 5) ? part[l] : make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (l == 0) result[blockIdx.x] = acc;
    }
}

}  // namespace

struct sgpu_dot {
    int device = 0;
    size_t n = 0;
    bool complex_coefs = false;
    std::vector<float> stored;  // stored order (dot_product/mod.rs:75-84), x2 floats when complex
    float *d_coefs = nullptr;
    Staging stage;
};

SGPU_EXPORT int sgpu_dot_create(const double *coefs, size_t n, sgpu_tapkind kind, sgpu_direction dir,
                                sgpu_dot **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (n && !coefs) return fail(SGPU_ERR_INVALID_ARGUMENT, "coefs is NULL");
    int dev = 0;
    int st = require_device(&dev, nullptr);
    if (st) return st;
    sgpu_dot *d = new (std::nothrow) sgpu_dot();
    if (!d) return fail(SGPU_ERR_ALLOC, "out of host memory");
    d->device = dev;
    d->n = n;
    d->complex_coefs = kind == SGPU_TAPS_COMPLEX;
    const size_t w = d->complex_coefs ? 2 : 1;
    d->stored.resize(n * w);
    for (size_t i = 0; i < n; ++i) {
        const size_t src = dir == SGPU_REVERSE ? n - 1 - i : i;
        for (size_t c = 0; c < w; ++c) d->stored[i * w + c] = (float)coefs[src * w + c];
    }
    if (n) {
        if (cudaMalloc(&d->d_coefs, n * w * sizeof(float)) != cudaSuccess) {
            delete d;
            return fail(SGPU_ERR_CUDA, "cudaMalloc(dot coefs) failed");
        }
        cudaMemcpy(d->d_coefs, d->stored.data(), n * w * sizeof(float), cudaMemcpyHostToDevice);
    }
    *out = d;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_dot_destroy(sgpu_dot *d) {
    if (!d) return SGPU_OK;
    DeviceGuard g(d->device);
    if (d->d_coefs) cudaFree(d->d_coefs);
    d->stage.release();
    delete d;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_dot_len(const sgpu_dot *d) { return d ? d->n : 0; }

SGPU_EXPORT int sgpu_dot_coefficients(const sgpu_dot *d, double *out) {
    if (!d || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    for (size_t i = 0; i < d->stored.size(); ++i) out[i] = (double)d->stored[i];
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_dot_execute(sgpu_dot *d, const float *x, size_t n_x, size_t x_stride, size_t n_vec,
                                 float *result, sgpu_mem mem, void *stream) {
    if (!d || !result) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    if (n_vec == 0) return SGPU_OK;
    if (n_x && !x) return fail(SGPU_ERR_INVALID_ARGUMENT, "null samples");
    DeviceGuard g(d->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t terms = n_x < d->n ? n_x : d->n;  // dot_product/mod.rs:160
    const float2 *d_x = reinterpret_cast<const float2 *>(x);
    float2 *d_r = reinterpret_cast<float2 *>(result);
    long long xs = (long long)x_stride;
    if (mem == SGPU_HOST) {
        int st = d->stage.ensure(n_vec * (n_x ? n_x : 1) * sizeof(float2), n_vec * sizeof(float2));
        if (st) return st;
        if (n_x)
            SGPU_CUDA(cudaMemcpy2DAsync(d->stage.in, n_x * sizeof(float2), x, x_stride * sizeof(float2),
                                        n_x * sizeof(float2), n_vec, cudaMemcpyHostToDevice, s));
        d_x = (const float2 *)d->stage.in;
        d_r = (float2 *)d->stage.out;
        xs = (long long)n_x;
    }
    if (d->complex_coefs)
        dot_kernel<true><<<(unsigned)n_vec, 256, 0, s>>>(d->d_coefs, (int)terms, d_x, xs, d_r);
    else
        dot_kernel<false><<<(unsigned)n_vec, 256, 0, s>>>(d->d_coefs, (int)terms, d_x, xs, d_r);
    SGPU_LAUNCH_CHECK();
    count_launch();
    if (mem == SGPU_HOST) {
        SGPU_CUDA(cudaMemcpyAsync(result, d_r, n_vec * sizeof(float2), cudaMemcpyDeviceToHost, s));
        SGPU_CUDA(cudaStreamSynchronize(s));
    }
    return SGPU_OK;
}
