// NCO (nco/mod.rs): numerically controlled oscillator bank.  One 32-bit phase accumulator per channel, a 1024-entry
// sine table, mix_up / mix_down of a block of samples (nco/mod.rs:141-172; the reference's *_block functions index an
// empty Vec and panic -- SURVEY Appendix A -- so the block form here is the per-sample loop they were meant to be:
// y[i] = mix(x[i]); step()).
#include "nco.cuh"

#include <cmath>
#include <mutex>

namespace sgpu {
namespace {

template <bool UP>
__global__ void __launch_bounds__(256) nco_mix_kernel(const float2 *__restrict__ in, const long long in_stride,
                                                      float2 *__restrict__ out, const long long out_stride,
                                                      const long long n, const unsigned *__restrict__ nco,
                                                      const unsigned nco_pos, const float2 *__restrict__ lut) {
    __shared__ float2 lut_s[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) lut_s[i] = lut[i];
    __syncthreads();
    const int ch = blockIdx.y;
    const unsigned delta = nco[2 * ch + 1];
    const unsigned theta0 = nco[2 * ch] + nco_pos * delta + (1u << 21);  // the index rounding folded in (nco/mod.rs:99-101)
    const float4 *__restrict__ x = reinterpret_cast<const float4 *>(in + (long long)ch * in_stride);
    float4 *__restrict__ y = reinterpret_cast<float4 *>(out + (long long)ch * out_stride);
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    auto mix = [&](const float vx, const float vy, const long long i, float &ox, float &oy) {
        const float2 cs = lut_s[(theta0 + (unsigned)i * delta) >> 22];
        if (UP) {  // (c + j s) x   (nco/mod.rs:142-145)
            ox = fmaf(cs.x, vx, -cs.y * vy);
            oy = fmaf(cs.x, vy, cs.y * vx);
        } else {   // (c - j s) x   (nco/mod.rs:148-151)
            ox = fmaf(cs.x, vx, cs.y * vy);
            oy = fmaf(cs.x, vy, -cs.y * vx);
        }
    };
    const long long pairs = vec ? n / 2 : 0;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < pairs; p += (long long)gridDim.x * 256) {
        const float4 v = x[p];
        float4 o;
        mix(v.x, v.y, 2 * p, o.x, o.y);
        mix(v.z, v.w, 2 * p + 1, o.z, o.w);
        y[p] = o;
    }
    const float2 *__restrict__ xs = in + (long long)ch * in_stride;
    float2 *__restrict__ ys = out + (long long)ch * out_stride;
    for (long long i = 2 * pairs + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float2 v = xs[i];
        float2 o;
        mix(v.x, v.y, i, o.x, o.y);
        ys[i] = o;
    }
}

struct LutCache {
    std::mutex m;
    float2 *d[64] = {};
};
LutCache g_lut;

// nco/mod.rs:176-188: the fractional part of theta / 2 pi, made positive, times 0xffffffff, truncated
uint32_t constrain(double theta) {
    double ip;
    double frac = std::modf(theta / (2.0 * M_PI), &ip);
    if (frac < 0.0) frac += 1.0;
    const double v = frac * 4294967295.0;
    if (!(v > 0.0)) return 0u;  // NaN and negatives saturate to 0 like Rust's `as u32`
    return v >= 4294967295.0 ? 0xffffffffu : (uint32_t)v;
}

// fold the stream position into the phase words so that they can be edited
void nco_rebase(sgpu_nco *n) {
    if (n->pos == 0) return;
    for (size_t c = 0; c < n->C; ++c) n->tab[2 * c] += n->pos * n->tab[2 * c + 1];
    n->pos = 0;
    n->dirty = true;
}

template <class F>
int nco_edit(sgpu_nco *n, size_t channel, F f) {
    if (!n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (channel != SGPU_ALL_CHANNELS && channel >= n->C) return fail(SGPU_ERR_INVALID_ARGUMENT, "channel %zu >= %zu", channel, n->C);
    nco_rebase(n);
    const size_t lo = channel == SGPU_ALL_CHANNELS ? 0 : channel, hi = channel == SGPU_ALL_CHANNELS ? n->C : channel + 1;
    for (size_t c = lo; c < hi; ++c) f(n->tab[2 * c], n->tab[2 * c + 1]);
    n->dirty = true;
    return SGPU_OK;
}

}  // namespace

int nco_lut(int device, const float2 **out) {
    if (device < 0 || device >= 64) return fail(SGPU_ERR_UNSUPPORTED, "device index %d", device);
    std::lock_guard<std::mutex> lock(g_lut.m);
    if (!g_lut.d[device]) {
        double table[1024];
        for (int i = 0; i < 1024; ++i) table[i] = std::sin(2.0 * M_PI * (double)i / 1024.0);  // nco/mod.rs:38-40
        std::vector<float2> h(1024);
        for (int i = 0; i < 1024; ++i) h[i] = make_float2((float)table[(i + 256) & 1023], (float)table[i]);
        float2 *d = nullptr;
        SGPU_CUDA(cudaMalloc(&d, 1024 * sizeof(float2)));
        SGPU_CUDA(cudaMemcpy(d, h.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
        g_lut.d[device] = d;
    }
    *out = g_lut.d[device];
    return SGPU_OK;
}

int nco_sync_table(sgpu_nco *n, cudaStream_t s) {
    if (!n->dirty) return SGPU_OK;
    // the table is pageable host memory: the copy has left it when cudaMemcpyAsync returns
    SGPU_CUDA(cudaMemcpyAsync(n->d_tab, n->tab.data(), n->tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    n->dirty = false;
    return SGPU_OK;
}

int nco_mix_launch(bool up, const float2 *in, long long in_stride, float2 *out, long long out_stride, long long n,
                   size_t C, const unsigned *d_tab, unsigned pos, const float2 *d_lut, cudaStream_t s) {
    if (n <= 0) return SGPU_OK;
    for (size_t c0 = 0; c0 < C; c0 += 65535) {
        const size_t cb = std::min<size_t>(65535, C - c0);
        const long long want = (n / 2 + 255) / 256;
        const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(want, std::max<long long>(1, (long long)(148 * 16 / cb))));
        dim3 grid(gx, (unsigned)cb);
        if (up) nco_mix_kernel<true><<<grid, 256, 0, s>>>(in + (long long)c0 * in_stride, in_stride, out + (long long)c0 * out_stride, out_stride, n, d_tab + 2 * c0, pos, d_lut);
        else nco_mix_kernel<false><<<grid, 256, 0, s>>>(in + (long long)c0 * in_stride, in_stride, out + (long long)c0 * out_stride, out_stride, n, d_tab + 2 * c0, pos, d_lut);
        SGPU_LAUNCH_CHECK();
        count_launch();
    }
    return SGPU_OK;
}

}  // namespace sgpu

using namespace sgpu;

SGPU_EXPORT int sgpu_nco_create(size_t n_channels, sgpu_nco **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "nco_create: out is NULL");
    *out = nullptr;
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "nco_create: n_channels == 0");
    int dev = 0;
    int st = require_device(&dev, nullptr);
    if (st) return st;
    sgpu_nco *n = new (std::nothrow) sgpu_nco();
    if (!n) return fail(SGPU_ERR_ALLOC, "out of host memory");
    n->device = dev;
    n->C = n_channels;
    n->tab.assign(2 * n_channels, 0u);  // theta = 0, delta_theta = 0 (nco/mod.rs:45-46)
    st = nco_lut(dev, &n->d_lut);
    if (st == SGPU_OK && cudaMalloc(&n->d_tab, 2 * n_channels * sizeof(uint32_t)) != cudaSuccess)
        st = fail(SGPU_ERR_CUDA, "cudaMalloc(nco phase words) failed");
    if (st) {
        delete n;
        return st;
    }
    *out = n;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_nco_destroy(sgpu_nco *n) {
    if (!n) return SGPU_OK;
    DeviceGuard g(n->device);
    if (n->d_tab) cudaFree(n->d_tab);
    n->stage.release();
    delete n;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_nco_clone(const sgpu_nco *n, sgpu_nco **out) {
    if (!n || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(n->device);
    int st = sgpu_nco_create(n->C, out);
    if (st) return st;
    (*out)->tab = n->tab;
    (*out)->pos = n->pos;
    (*out)->dirty = true;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_nco_channels(const sgpu_nco *n) { return n ? n->C : 0; }

SGPU_EXPORT int sgpu_nco_reset(sgpu_nco *n) {  // nco/mod.rs:53-56
    if (!n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    n->pos = 0;
    std::fill(n->tab.begin(), n->tab.end(), 0u);
    n->dirty = true;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_nco_set_frequency(sgpu_nco *n, size_t channel, double delta_theta) {  // nco/mod.rs:59-61
    const uint32_t v = constrain(delta_theta);
    return nco_edit(n, channel, [v](uint32_t &, uint32_t &d) { d = v; });
}
SGPU_EXPORT int sgpu_nco_adjust_frequency(sgpu_nco *n, size_t channel, double dt) {  // nco/mod.rs:64-66 (wrapping)
    const uint32_t v = constrain(dt);
    return nco_edit(n, channel, [v](uint32_t &, uint32_t &d) { d += v; });
}
SGPU_EXPORT int sgpu_nco_set_phase(sgpu_nco *n, size_t channel, double phi) {  // nco/mod.rs:79-81
    const uint32_t v = constrain(phi);
    return nco_edit(n, channel, [v](uint32_t &t, uint32_t &) { t = v; });
}
SGPU_EXPORT int sgpu_nco_adjust_phase(sgpu_nco *n, size_t channel, double delta_phi) {  // nco/mod.rs:84-86 (wrapping)
    const uint32_t v = constrain(delta_phi);
    return nco_edit(n, channel, [v](uint32_t &t, uint32_t &) { t += v; });
}
SGPU_EXPORT int sgpu_nco_set(sgpu_nco *n, size_t channel, uint32_t theta, uint32_t delta_theta) {
    return nco_edit(n, channel, [=](uint32_t &t, uint32_t &d) {
        t = theta;
        d = delta_theta;
    });
}
SGPU_EXPORT int sgpu_nco_get(const sgpu_nco *n, size_t channel, uint32_t *theta, uint32_t *delta_theta) {
    if (!n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (channel >= n->C) return fail(SGPU_ERR_INVALID_ARGUMENT, "channel %zu >= %zu", channel, n->C);
    if (theta) *theta = n->tab[2 * channel] + n->pos * n->tab[2 * channel + 1];
    if (delta_theta) *delta_theta = n->tab[2 * channel + 1];
    return SGPU_OK;
}
SGPU_EXPORT uint32_t sgpu_nco_constrain(double theta) { return constrain(theta); }  // nco/mod.rs:176-188

SGPU_EXPORT int sgpu_nco_step(sgpu_nco *n, uint64_t count) {  // `count` times NCO::step (nco/mod.rs:93-96)
    if (!n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    n->pos += (uint32_t)count;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_nco_mix_block(sgpu_nco *n, int up, const float *in, size_t n_in, size_t in_stride, float *out,
                                   size_t out_stride, sgpu_mem mem, void *stream) {
    if (!n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (n_in == 0) return SGPU_OK;
    if (!in || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (n->C > 1 && (in_stride < n_in || out_stride < n_in)) return fail(SGPU_ERR_INVALID_ARGUMENT, "stride < n_in");
    DeviceGuard g(n->device);
    cudaStream_t s = (cudaStream_t)stream;
    int st = nco_sync_table(n, s);
    if (st) return st;
    if (mem == SGPU_DEVICE) {
        st = nco_mix_launch(up != 0, reinterpret_cast<const float2 *>(in), (long long)in_stride, reinterpret_cast<float2 *>(out),
                            (long long)out_stride, (long long)n_in, n->C, n->d_tab, n->pos, n->d_lut, s);
    } else {
        const size_t bytes = n->C * n_in * sizeof(float2);
        st = n->stage.ensure(bytes, bytes);
        if (st) return st;
        SGPU_CUDA(cudaMemcpy2DAsync(n->stage.in, n_in * sizeof(float2), in, in_stride * sizeof(float2), n_in * sizeof(float2), n->C,
                                    cudaMemcpyHostToDevice, s));
        st = nco_mix_launch(up != 0, (const float2 *)n->stage.in, (long long)n_in, (float2 *)n->stage.out, (long long)n_in,
                            (long long)n_in, n->C, n->d_tab, n->pos, n->d_lut, s);
        if (st == SGPU_OK) {
            SGPU_CUDA(cudaMemcpy2DAsync(out, out_stride * sizeof(float2), n->stage.out, n_in * sizeof(float2), n_in * sizeof(float2),
                                        n->C, cudaMemcpyDeviceToHost, s));
            SGPU_CUDA(cudaStreamSynchronize(s));
        }
    }
    if (st) return st;
    n->pos += (uint32_t)n_in;  // one step() per sample
    return SGPU_OK;
}
