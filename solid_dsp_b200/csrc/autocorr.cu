// AutoCorrelator on sm_100a (the sliding-window step that sits in front of the FIR).
//
// Reference loops replaced (relative to the reference's src/):
//   AutoCorrelator::push / write / execute / execute_block / get_energy
//                                     filter/auto_correlator/mod.rs:99-111,130-141,165-191,214-216
//   Window::new(capacity, delay) / push / to_vec       window/mod.rs:17-34,44-51,63-71
//
// What the reference computes (the Window delay quirk included, see oracle/solid_oracle.c): with
// W = window_size, d = delay, Wd = max(W - d, 0),
//     r[n] = sum_{i < Wd} x[n-i] * conj(x[n-d-i]),     energy = sum_{i < W} |x[n-i]|^2
// over the stream of everything pushed so far (zeros before it).
//
// Kernel: HBM-bound (16 B per sample, 8 + 4*Wd/R flop).  A block takes 2048 outputs of one channel:
// the samples (plus W-1 of history) are staged in shared memory, every lag product
// p[j] = x[j] conj(x[j-d]) is formed ONCE, and a thread then slides the window sum over its R = 8
// consecutive outputs: r[n+1] = r[n] + p[n+1] - p[n+1-Wd], restarted from a direct sum (through sums
// of 8 for long windows) every 8 outputs so that rounding never accumulates over more than 7 updates.
// p lives in a skewed layout (one pad slot per 8) so that lanes 8 samples apart hit different banks; the
// samples arrive by 8-byte cp.async and a run's 8 outputs leave straight from registers.  Runs are aligned to the absolute stream position, which makes the results
// independent of how the stream is cut into calls (bit for bit).
#include <algorithm>

#include "sgpu_common.cuh"

using namespace sgpu;

namespace {

constexpr int kNT = 256, kR = 8, kTileOut = kNT * kR;
constexpr int kMaxWindow = 8192;
constexpr int kHistExtra = 7;  // the run that contains a call's first output may start 7 samples early

struct AcArgs {
    const float2 *in;
    float2 *out;
    const float2 *hist;  // [C][HW] last HW = W + 7 samples before this call, oldest first
    long long in_stride, out_stride, n_in;
    int W, d, Wd, HW;
    int shift;  // samples pushed before this call, mod 8
    int vec_out;  // out base 16-byte aligned and out_stride even
};

__device__ __forceinline__ int skew(int j) { return j + (j >> 3); }

__device__ __forceinline__ float2 ac_fetch(const float2 *__restrict__ x, const float2 *__restrict__ hist, long long i,
                                           long long n_in, int W) {
    if (i >= 0) return i < n_in ? x[i] : make_float2(0.f, 0.f);
    const long long h = (long long)W + i;
    return h >= 0 ? hist[h] : make_float2(0.f, 0.f);
}

__global__ void __launch_bounds__(kNT) autocorr_kernel(const AcArgs a) {
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, ch = blockIdx.y;
    const int H = a.W - 1;                    // samples in front of the tile that its outputs reach
    const int NX = kTileOut + H;              // staged samples: xs[k] = x[n_base - H + k]
    const int NP = kTileOut + a.Wd - 1;       // lag products:   p[k]  = product at n_base - (Wd-1) + k
    const int NB = NP / 8;                    // whole blocks of 8 lag products
    float2 *xs = sm;
    float2 *ps = sm + ((NX + 1) & ~1);
    float2 *bs = ps + ((NP + NP / 8 + 2) & ~1);
    // Tiles are aligned to the ABSOLUTE stream position (a.shift = samples pushed so far mod 8), so
    // the association order of every output's sum -- and with it the rounding -- does not depend on
    // where the stream was cut into calls: concat(execute_block(a), execute_block(b)) is bit-identical
    // to execute_block(a ++ b).  n_base may be negative for the first tile; those outputs are dropped.
    const long long n_base = (long long)blockIdx.x * kTileOut - a.shift;
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    const float2 *__restrict__ hist = a.hist + (long long)ch * a.HW;
    if (n_base - H >= 0 && n_base + kTileOut <= a.n_in) {  // interior tile: straight 8-byte LDGSTS
        const float2 *src = x + (n_base - H);
        for (int k = tid; k < NX; k += kNT) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(xs + k);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + k) : "memory");
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        for (int k = tid; k < NX; k += kNT) xs[k] = ac_fetch(x, hist, n_base - H + k, a.n_in, a.HW);
    }
    __syncthreads();
    // p at stream position n uses x[n] and x[n-d]; position of p[k] is n_base - (Wd-1) + k, i.e.
    // xs index k + H - (Wd-1) = k + d   (H - Wd + 1 = d)  and, d earlier, xs index k
    for (int k = tid; k < NP; k += kNT) {
        const float2 u = xs[k + a.d], v = xs[k];  // x[n], x[n-d]
        ps[skew(k)] = make_float2(u.x * v.x + u.y * v.y, u.y * v.x - u.x * v.y);  // u * conj(v)
    }
    __syncthreads();
    const bool blocked = a.Wd >= 16;  // long windows: first sum of a run from sums of 8
    if (blocked) {
        for (int m = tid; m < NB; m += kNT) {
            float2 t = ps[skew(8 * m)];
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const float2 q = ps[skew(8 * m + j)];
                t.x += q.x;
                t.y += q.y;
            }
            bs[m] = t;
        }
        __syncthreads();
    }
    // outputs n_base + tid*R + r: window of p indices [tid*R + r, tid*R + r + Wd)
    const int k0 = tid * kR;
    float2 s = make_float2(0.f, 0.f);
    if (blocked) {
        const int nb = a.Wd >> 3;
        for (int m = 0; m < nb; ++m) {
            const float2 q = bs[tid + m];
            s.x += q.x;
            s.y += q.y;
        }
        for (int i = 8 * nb; i < a.Wd; ++i) {
            const float2 q = ps[skew(k0 + i)];
            s.x += q.x;
            s.y += q.y;
        }
    } else {
        for (int i = a.Wd - 1; i >= 0; --i) {  // newest first, like the reference's zip/sum
            const float2 q = ps[skew(k0 + i)];
            s.x += q.x;
            s.y += q.y;
        }
    }
    float2 y[kR];
    y[0] = s;
#pragma unroll
    for (int r = 1; r < kR; ++r) {
        const float2 add = ps[skew(k0 + r + a.Wd - 1)], sub = ps[skew(k0 + r - 1)];
        s.x += add.x - sub.x;
        s.y += add.y - sub.y;
        y[r] = s;
    }
    if (a.Wd == 0) {
#pragma unroll
        for (int r = 0; r < kR; ++r) y[r] = make_float2(0.f, 0.f);
    }
    // the run's 8 outputs are 64 contiguous bytes: straight from registers (the L2 merges the sectors)
    const long long n0 = n_base + k0;
    float2 *__restrict__ out = a.out + (long long)ch * a.out_stride + n0;
    if (n0 >= 0 && n0 + kR <= a.n_in) {
        if (a.vec_out && (a.shift & 1) == 0) {
#pragma unroll
            for (int r = 0; r < kR; r += 2)
                *reinterpret_cast<float4 *>(out + r) = make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y);
        } else {
#pragma unroll
            for (int r = 0; r < kR; ++r) out[r] = y[r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < kR; ++r)
            if (n0 + r >= 0 && n0 + r < a.n_in) out[r] = y[r];
    }
}

// new history = last W samples of (old history ++ x[0..n_in))
__global__ void ac_hist_update_kernel(const float2 *__restrict__ in, long long in_stride, long long n_in,
                                      const float2 *__restrict__ hist_old, float2 *__restrict__ hist_new, int W) {
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W) return;
    const long long s = n_in - W + i;
    float2 v;
    if (s >= 0) v = in[(long long)ch * in_stride + s];
    else v = hist_old[(long long)ch * W + (W + s)];
    hist_new[(long long)ch * W + i] = v;
}

// execute() / get_energy() of the current state: one block per channel over the history
// out[ch] = (re, im, energy, 0) in double
__global__ void __launch_bounds__(256) ac_point_kernel(const float2 *__restrict__ hist, int W, int d, int Wd,
                                                       double4 *__restrict__ out) {
    __shared__ double red[3][256];
    const float2 *h = hist + (long long)blockIdx.x * (W + kHistExtra) + kHistExtra;  // h[W-1] = newest
    double re = 0.0, im = 0.0, en = 0.0;
    for (int i = threadIdx.x; i < W; i += 256) {
        const float2 u = h[W - 1 - i];
        en += (double)u.x * u.x + (double)u.y * u.y;
        if (i < Wd) {
            const float2 v = h[W - 1 - i - d];
            re += (double)u.x * v.x + (double)u.y * v.y;
            im += (double)u.y * v.x - (double)u.x * v.y;
        }
    }
    red[0][threadIdx.x] = re; red[1][threadIdx.x] = im; red[2][threadIdx.x] = en;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = make_double4(red[0][0], red[1][0], red[2][0], 0.0);
}

}  // namespace

struct sgpu_autocorr {
    int device = 0, sm_count = 0;
    size_t W = 0, d = 0, C = 0;
    float2 *d_hist[2] = {nullptr, nullptr};  // ping-pong: a launch reads one while the update writes the other
    int cur = 0;
    uint64_t pos = 0;  // samples pushed so far (tile alignment, see autocorr_kernel)
    double4 *d_point = nullptr;
    HostPipe pipe;
    Staging stage;
};

namespace {

int ac_launch(sgpu_autocorr *f, const float2 *d_in, long long n_in, long long istr, float2 *d_out, long long ostr,
              cudaStream_t s) {
    AcArgs a{};
    a.in = d_in; a.out = d_out; a.hist = f->d_hist[f->cur];
    a.in_stride = istr; a.out_stride = ostr; a.n_in = n_in;
    a.W = (int)f->W; a.d = (int)f->d; a.Wd = f->d < f->W ? (int)(f->W - f->d) : 0;
    a.HW = (int)f->W + kHistExtra;
    if (a.Wd == 0) a.d = 0;  // delay >= window: all outputs are zero; keep the index math in range
    a.shift = (int)(f->pos & 7);
    a.vec_out = ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0) && (ostr % 2 == 0);
    const size_t nx = kTileOut + f->W - 1, np = kTileOut + (size_t)a.Wd;
    const size_t smem = (((nx + 1) & ~(size_t)1) + ((np + np / 8 + 4) & ~(size_t)1) + np / 8 + 2) * sizeof(float2);
    constexpr size_t kMaxGridY = 65535;  // channels ride in grid.y: larger handles go in channel blocks
    const float2 *hist_all = f->d_hist[f->cur];
    if (d_out) SGPU_CUDA(cudaFuncSetAttribute(autocorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (size_t c0 = 0; c0 < f->C; c0 += kMaxGridY) {
        const unsigned nc = (unsigned)std::min<size_t>(kMaxGridY, f->C - c0);
        if (d_out) {
            a.in = d_in + (long long)c0 * istr;
            a.out = d_out + (long long)c0 * ostr;
            a.hist = hist_all + c0 * (size_t)a.HW;
            dim3 grid((unsigned)ceil_div((size_t)n_in + a.shift, kTileOut), nc);
            autocorr_kernel<<<grid, kNT, smem, s>>>(a);
            SGPU_LAUNCH_CHECK();
            count_launch();
        }
        dim3 hg((unsigned)ceil_div(f->W + kHistExtra, 128), nc);
        ac_hist_update_kernel<<<hg, 128, 0, s>>>(d_in + (long long)c0 * istr, istr, n_in, hist_all + c0 * (size_t)a.HW,
                                                 f->d_hist[f->cur ^ 1] + c0 * (size_t)a.HW, a.HW);
        SGPU_LAUNCH_CHECK();
        count_launch();
    }
    f->cur ^= 1;
    f->pos += (uint64_t)n_in;
    return SGPU_OK;
}

int ac_point(sgpu_autocorr *f, double *host4) {  // host4: [C][4] doubles
    const int Wd = f->d < f->W ? (int)(f->W - f->d) : 0;
    ac_point_kernel<<<(unsigned)f->C, 256>>>(f->d_hist[f->cur], (int)f->W, Wd ? (int)f->d : 0, Wd, f->d_point);
    SGPU_LAUNCH_CHECK();
    count_launch();
    SGPU_CUDA(cudaMemcpy(host4, f->d_point, f->C * sizeof(double4), cudaMemcpyDeviceToHost));
    return SGPU_OK;
}

}  // namespace

SGPU_EXPORT int sgpu_autocorr_create(size_t window_size, size_t delay, size_t n_channels, sgpu_autocorr **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (window_size == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "window_size == 0 (Window::new asserts capacity > 0)");
    if (window_size > (size_t)kMaxWindow) return fail(SGPU_ERR_UNSUPPORTED, "window_size above %d", kMaxWindow);
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    int dev = 0, sms = 0;
    int st = require_device(&dev, &sms);
    if (st) return st;
    sgpu_autocorr *f = new (std::nothrow) sgpu_autocorr();
    if (!f) return fail(SGPU_ERR_ALLOC, "out of host memory");
    f->device = dev; f->sm_count = sms;
    f->W = window_size; f->d = delay; f->C = n_channels;
    const size_t hb = n_channels * (window_size + kHistExtra) * sizeof(float2);
    if (cudaMalloc(&f->d_hist[0], hb) != cudaSuccess || cudaMalloc(&f->d_hist[1], hb) != cudaSuccess ||
        cudaMalloc(&f->d_point, n_channels * sizeof(double4)) != cudaSuccess) {
        sgpu_autocorr_destroy(f);
        return fail(SGPU_ERR_CUDA, "cudaMalloc(autocorrelator state) failed");
    }
    cudaMemset(f->d_hist[0], 0, hb);
    cudaMemset(f->d_hist[1], 0, hb);
    *out = f;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_autocorr_destroy(sgpu_autocorr *f) {
    if (!f) return SGPU_OK;
    DeviceGuard g(f->device);
    for (int i = 0; i < 2; ++i)
        if (f->d_hist[i]) cudaFree(f->d_hist[i]);
    if (f->d_point) cudaFree(f->d_point);
    f->pipe.release();
    f->stage.release();
    delete f;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_autocorr_window_size(const sgpu_autocorr *f) { return f ? f->W : 0; }
SGPU_EXPORT size_t sgpu_autocorr_delay(const sgpu_autocorr *f) { return f ? f->d : 0; }
SGPU_EXPORT size_t sgpu_autocorr_channels(const sgpu_autocorr *f) { return f ? f->C : 0; }

SGPU_EXPORT int sgpu_autocorr_execute_block(sgpu_autocorr *f, const float *in, size_t n_in, size_t in_stride, float *out,
                                            size_t out_stride, size_t *n_out_p, sgpu_mem mem, void *stream) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (n_out_p) *n_out_p = n_in;  // one output per input (auto_correlator/mod.rs:184-191)
    if (n_in == 0) return SGPU_OK;
    if (!in || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (f->C > 1 && in_stride < n_in) return fail(SGPU_ERR_INVALID_ARGUMENT, "in_stride < n_in");
    if (out_stride < n_in) return fail(SGPU_ERR_CAPACITY, "out capacity %zu < %zu outputs", out_stride, n_in);
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    auto run = [f](const float2 *d_in, size_t nc, long long istr, float2 *d_out, long long ostr, size_t /*nout*/,
                   cudaStream_t st_) -> int { return ac_launch(f, d_in, (long long)nc, istr, d_out, ostr, st_); };
    if (mem == SGPU_DEVICE)
        return run(reinterpret_cast<const float2 *>(in), n_in, (long long)in_stride, reinterpret_cast<float2 *>(out),
                   (long long)out_stride, n_in, s);
    return host_pipeline(f->pipe, f->C, in, n_in, in_stride, out, out_stride, 1, [](size_t nc) { return nc; }, run, s);
}

// write(): push samples without producing outputs (auto_correlator/mod.rs:130-141)
SGPU_EXPORT int sgpu_autocorr_write(sgpu_autocorr *f, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
                                    void *stream) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (n_in == 0) return SGPU_OK;
    if (!in) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    const float2 *d_in = reinterpret_cast<const float2 *>(in);
    long long istr = (long long)in_stride;
    if (mem == SGPU_HOST) {
        int st = f->stage.ensure(f->C * n_in * sizeof(float2), 0);
        if (st) return st;
        SGPU_CUDA(cudaMemcpy2DAsync(f->stage.in, n_in * sizeof(float2), in, in_stride * sizeof(float2),
                                    n_in * sizeof(float2), f->C, cudaMemcpyHostToDevice, s));
        d_in = (const float2 *)f->stage.in;
        istr = (long long)n_in;
    }
    int st = ac_launch(f, d_in, (long long)n_in, istr, nullptr, 0, s);
    if (st) return st;
    if (mem == SGPU_HOST) SGPU_CUDA(cudaStreamSynchronize(s));
    return SGPU_OK;
}

// execute(): the correlator output of the current window, [n_channels] complex doubles (HOST)
SGPU_EXPORT int sgpu_autocorr_execute(sgpu_autocorr *f, double *out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    std::vector<double> tmp(f->C * 4);
    int st = ac_point(f, tmp.data());
    if (st) return st;
    for (size_t c = 0; c < f->C; ++c) { out[2 * c] = tmp[4 * c]; out[2 * c + 1] = tmp[4 * c + 1]; }
    return SGPU_OK;
}

// get_energy(): sum of |x|^2 over the last window_size samples, [n_channels] doubles (HOST)
SGPU_EXPORT int sgpu_autocorr_get_energy(sgpu_autocorr *f, double *out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    std::vector<double> tmp(f->C * 4);
    int st = ac_point(f, tmp.data());
    if (st) return st;
    for (size_t c = 0; c < f->C; ++c) out[c] = tmp[4 * c + 2];
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_autocorr_reset(sgpu_autocorr *f) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_hist[f->cur], 0, f->C * (f->W + kHistExtra) * sizeof(float2)));
    f->pos = 0;
    return SGPU_OK;
}

// state = the last window_size samples per channel, oldest first, [n_channels][window_size] cf32 (HOST)
SGPU_EXPORT int sgpu_autocorr_get_state(sgpu_autocorr *f, float *state) {
    if (!f || !state) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemcpy2D(state, f->W * sizeof(float2), f->d_hist[f->cur] + kHistExtra, (f->W + kHistExtra) * sizeof(float2),
                           f->W * sizeof(float2), f->C, cudaMemcpyDeviceToHost));
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_autocorr_set_state(sgpu_autocorr *f, const float *state) {
    if (!f || !state) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_hist[f->cur], 0, f->C * (f->W + kHistExtra) * sizeof(float2)));
    SGPU_CUDA(cudaMemcpy2D(f->d_hist[f->cur] + kHistExtra, (f->W + kHistExtra) * sizeof(float2), state, f->W * sizeof(float2),
                           f->W * sizeof(float2), f->C, cudaMemcpyHostToDevice));
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_autocorr_clone(const sgpu_autocorr *f, sgpu_autocorr **out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    sgpu_autocorr *c = nullptr;
    int st = sgpu_autocorr_create(f->W, f->d, f->C, &c);
    if (st) return st;
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemcpy(c->d_hist[c->cur], f->d_hist[f->cur], f->C * (f->W + kHistExtra) * sizeof(float2),
                         cudaMemcpyDeviceToDevice));
    c->pos = f->pos;
    *out = c;
    return SGPU_OK;
}
