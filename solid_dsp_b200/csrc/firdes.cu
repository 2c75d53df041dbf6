// firdes_kaiser on the device: h[i] = sinc(2 fc t_i) * I0(beta sqrt(1 - r_i^2)) / I0(beta) in f64, one
// thread per tap of one design -- firdes/mod.rs:243-253 (kaiser_beta), :278-305 (firdes_kaiser), windows/kaiser.rs:33-46,
// math/mod.rs:17-27 (sinc), :41-100 (besseli / lnbesseli, 64-term log-domain series), :171-183 (lngamma).  The same
// operations in the same order as the host restatement (solid_dsp_b200/filter/firdes.py); only the last bits of the
// device's log / exp / sin / cos differ from the host's libm.  A bank of per-channel filters (one design per channel,
// sgpu_fir_create_per_channel) is where this pays: 4096 designs x 256 taps x 64 series terms.
#include "sgpu_common.cuh"

using namespace sgpu;

namespace {

constexpr int kBesselIterations = 64;  // math/mod.rs:8
constexpr double kPi = 3.14159265358979323846;

__device__ double d_sinc(double x) {  // math/mod.rs:17-27
    if (fabs(x) < 0.01) return cos(kPi * x / 2.0) * cos(kPi * x / 4.0) * cos(kPi * x / 8.0);
    return sin(kPi * x) / (kPi * x);
}

// math/mod.rs:171-183.  The reference recurses lngamma(x) = lngamma(x + 1) - ln(x) for x < 10; unrolled here with the
// subtractions in the order the recursion returns them (innermost first).
__device__ double d_lngamma(double x) {
    if (x < 0.0) return 0.0;
    double top = x;
    int k = 0;
    while (top < 10.0) {
        top += 1.0;
        ++k;
    }
    // x + k is built by the same k additions of 1.0 the recursion performs
    double g = 0.5 * (log(2.0 * kPi) - log(top));
    g = g + top * (log(top + (1.0 / (12.0 * top - 0.1 / top))) - 1.0);
    for (int j = k - 1; j >= 0; --j) {
        double xj = x;
        for (int a = 0; a < j; ++a) xj += 1.0;
        g -= log(xj);
    }
    return g;
}

__device__ double d_lnbesseli0(double z) {  // math/mod.rs:66-100 with nu = 0 (the early outs are taken by the caller)
    const double t0 = 0.0 * log(0.5 * z);
    double y = 0.0;
    for (int k = 0; k < kBesselIterations; ++k) {
        const double t1 = 2.0 * (double)k * log(0.5 * z);
        const double t2 = d_lngamma((double)k + 1.0);
        const double t3 = d_lngamma(0.0 + (double)k + 1.0);
        y += exp(t1 - t2 - t3);
    }
    return t0 + log(y);
}

__device__ double d_besseli0(double z) {  // math/mod.rs:41-64 with nu = 0
    if (z == 0.0) return 1.0;
    if (z < 0.001 * sqrt(0.0 + 1.0)) return pow(0.5 * z, 0.0) / exp(d_lngamma(0.0 + 1.0));
    return exp(d_lnbesseli0(z));
}

__device__ double d_kaiser_beta(double as) {  // firdes/mod.rs:243-253
    const double a = fabs(as);
    if (a > 50.0) return 0.1102 * (a - 8.7);
    if (a > 21.0) return 0.5842 * pow(a - 21.0, 0.4) + 0.07886 * (a - 21.0);
    return 0.0;
}

// blockIdx.y = design; params[d] = (cutoff, stop-band attenuation, mu)
__global__ void __launch_bounds__(128) firdes_kaiser_kernel(const double *__restrict__ params, int n, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double fc = params[3 * blockIdx.y], as = params[3 * blockIdx.y + 1], mu = params[3 * blockIdx.y + 2];
    const double beta = d_kaiser_beta(as);
    const double t = (double)i - (double)(n - 1) / 2.0 + mu;      // firdes/mod.rs:294
    const double h1 = d_sinc(2.0 * fc * t);                       // :296
    const double tw = (double)i - (double)(n - 1) / 2.0;          // windows/kaiser.rs:41-45
    const double r = 2.0 * tw / (double)(n - 1);
    const double h2 = d_besseli0(beta * sqrt(1.0 - r * r)) / d_besseli0(beta);
    out[(size_t)blockIdx.y * n + i] = h1 * h2;
}

}  // namespace

SGPU_EXPORT int sgpu_firdes_kaiser(size_t filter_length, const double *cutoff_frequency, const double *stop_band_attenuation,
                                   const double *fractional_sample_offset, size_t n_designs, double *out, sgpu_mem mem,
                                   void *stream) {
    if (!cutoff_frequency || !stop_band_attenuation || !out)
        return fail(SGPU_ERR_INVALID_ARGUMENT, "firdes_kaiser: null argument");
    if (n_designs == 0 || filter_length == 0) return SGPU_OK;  // an empty Vec in the reference
    if (filter_length > (1u << 24) || n_designs > 65535)
        return fail(SGPU_ERR_UNSUPPORTED, "firdes_kaiser: filter_length / n_designs beyond supported range");
    std::vector<double> params(3 * n_designs);
    for (size_t d = 0; d < n_designs; ++d) {
        const double mu = fractional_sample_offset ? fractional_sample_offset[d] : 0.0;
        const double fc = cutoff_frequency[d], as = stop_band_attenuation[d];
        // the checks and their order: firdes/mod.rs:284-290 (NaN fails the range tests like `contains` does)
        if (!(mu >= -0.5 && mu <= 0.5)) return fail(SGPU_ERR_FIRDES_MU, "Firdes Error: Invalid Mu Range [-0.5, 0.5]");
        if (!(fc >= 0.0 && fc <= 0.5)) return fail(SGPU_ERR_FIRDES_BANDWIDTH, "Firdes Error: Invalid Bandwidth [0, 0.5]");
        if (as <= 0.0) return fail(SGPU_ERR_FIRDES_STOP_BAND_LEVEL, "Firdes Error: Invalid Stop Band Attenuation (0, inf)");
        params[3 * d] = fc;
        params[3 * d + 1] = as;
        params[3 * d + 2] = mu;
    }
    int dev = 0;
    int st = require_device(&dev, nullptr);
    if (st) return st;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t n_out = n_designs * filter_length;
    double *d_params = nullptr, *d_out = mem == SGPU_DEVICE ? out : nullptr;
    SGPU_CUDA(cudaMallocAsync(&d_params, params.size() * sizeof(double), s));
    if (mem != SGPU_DEVICE && cudaMallocAsync(&d_out, n_out * sizeof(double), s) != cudaSuccess) {
        cudaFreeAsync(d_params, s);
        return fail(SGPU_ERR_CUDA, "cudaMallocAsync(%zu taps) failed", n_out);
    }
    auto cleanup = [&]() {
        cudaFreeAsync(d_params, s);
        if (mem != SGPU_DEVICE) cudaFreeAsync(d_out, s);
    };
    // `params` is pageable: the copy has returned from the staging buffer before this function does
    cudaError_t e = cudaMemcpyAsync(d_params, params.data(), params.size() * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        const dim3 grid((unsigned)ceil_div(filter_length, 128), (unsigned)n_designs);
        firdes_kaiser_kernel<<<grid, 128, 0, s>>>(d_params, (int)filter_length, d_out);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && mem != SGPU_DEVICE) {
        e = cudaMemcpyAsync(out, d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    cleanup();
    if (e != cudaSuccess) return fail(SGPU_ERR_CUDA, "firdes_kaiser: %s", cudaGetErrorString(e));
    return SGPU_OK;
}
