// In-library multi-GPU context (SURVEY 8e, Appendix D `sgpu_ctx_create`): ONE caller thread, host buffers, every GPU of
// the box behind one filter object -- what a Rust caller like the reference's main.rs:39-41 (one Vec in, one Vec out)
// needs to use more than one GPU without writing a launcher.
//
//   * independent channels (decimator / interpolator / IIR batches): contiguous channel ranges per device, no exchange;
//   * one FIR / decimating-FIR stream: contiguous time segments; the T-1 sample halo of segment d > 0 is sliced out of
//     the caller's host buffer ("halo via host slicing") and written as that device's filter history; the
//     decimator's segments start where the stream's phase counter is 0 (fir/decim.rs:221-228);
//   * the outputs of every device land in the caller's one host buffer: that is the gather -- no collective is needed
//     on this path (north star: "NCCL ... only to gather outputs where required"; device-resident multi-process
//     sharding lives in solid_dsp_b200/sharding.py over torch.distributed).
//
// Built on the public C ABI only (sgpu_fir_* / sgpu_interp_* / sgpu_iir_*): one handle per device, one host thread per
// device for the duration of a call (each SGPU_HOST call pipelines H2D / kernels / D2H on its own streams).
#include <algorithm>
#include <string>
#include <thread>

#include "sgpu_common.cuh"

using namespace sgpu;

struct sgpu_ctx {
    std::vector<int> devices;
};

namespace {
enum ShardKind { kFir = 0, kInterp = 1, kIir = 2 };
struct Shard {
    int device = 0;
    void *handle = nullptr;
    size_t ch_first = 0, ch_count = 0;
};
}  // namespace

struct sgpu_sharded {
    ShardKind kind = kFir;
    size_t C = 0, T = 0, M = 1, L = 1;
    bool is_decim = false;
    bool by_stream = false;  // one FIR stream cut into time segments (C == 1)
    std::vector<Shard> shards;
    int last_segments = 0;   // segments the last call really used
};

namespace {

int destroy_handle(ShardKind k, void *h) {
    if (!h) return SGPU_OK;
    switch (k) {
        case kFir: return sgpu_fir_destroy((sgpu_fir *)h);
        case kInterp: return sgpu_interp_destroy((sgpu_interp *)h);
        default: return sgpu_iir_destroy((sgpu_iir *)h);
    }
}

// run fn(shard index) on one host thread per shard, each with its device current; first error wins
template <class F>
int for_each_shard(size_t n, const std::vector<Shard> &shards, F fn) {
    std::vector<int> status(n, SGPU_OK);
    std::vector<std::string> messages(n);
    auto body = [&](size_t i) {
        if (cudaSetDevice(shards[i].device) != cudaSuccess) {
            status[i] = SGPU_ERR_CUDA;
            messages[i] = "cudaSetDevice failed";
            return;
        }
        status[i] = fn(i);
        if (status[i] != SGPU_OK) messages[i] = sgpu_last_error();  // the message is thread-local: carry it over
    };
    if (n == 1) {
        int prev = 0;
        cudaGetDevice(&prev);
        body(0);
        cudaSetDevice(prev);
    } else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (size_t i = 0; i < n; ++i) th.emplace_back(body, i);
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < n; ++i)
        if (status[i] != SGPU_OK) return fail(status[i], "shard %zu (device %d): %s", i, shards[i].device, messages[i].c_str());
    return SGPU_OK;
}

// channel ranges: contiguous, the remainder over the first shards; shards that would be empty are not created
void plan_channels(const sgpu_ctx *ctx, size_t C, std::vector<Shard> &out) {
    const int world = (int)std::min<size_t>(ctx->devices.size(), C);
    for (int r = 0; r < world; ++r) {
        size_t first = 0, count = 0;
        sgpu_shard_channels(C, world, r, &first, &count);
        if (count == 0) continue;
        Shard s;
        s.device = ctx->devices[r];
        s.ch_first = first;
        s.ch_count = count;
        out.push_back(s);
    }
}

template <class Create>
int build(sgpu_ctx *ctx, sgpu_sharded *sh, Create create, sgpu_sharded **out) {
    int prev = 0;
    cudaGetDevice(&prev);
    int st = SGPU_OK;
    for (auto &s : sh->shards) {
        if (cudaSetDevice(s.device) != cudaSuccess) {
            st = fail(SGPU_ERR_CUDA, "cudaSetDevice(%d) failed", s.device);
            break;
        }
        st = create(s);
        if (st) break;
    }
    cudaSetDevice(prev);
    if (st) {
        sgpu_sharded_destroy(sh);
        return st;
    }
    (void)ctx;
    *out = sh;
    return SGPU_OK;
}

}  // namespace

SGPU_EXPORT int sgpu_ctx_create_devices(const int *devices, int n, sgpu_ctx **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SGPU_ERR_NO_DEVICE, "no CUDA device; libsolid_gpu has no CPU fallback");
    if (n <= 0 || !devices) return fail(SGPU_ERR_INVALID_ARGUMENT, "ctx_create: empty device list");
    sgpu_ctx *c = new (std::nothrow) sgpu_ctx();
    if (!c) return fail(SGPU_ERR_ALLOC, "out of host memory");
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) {
            delete c;
            return fail(SGPU_ERR_INVALID_ARGUMENT, "ctx_create: device %d of %d", devices[i], ndev);
        }
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, devices[i]) != cudaSuccess || p.major < 10) {
            delete c;
            return fail(SGPU_ERR_NO_DEVICE, "device %d is not sm_100 class", devices[i]);
        }
        c->devices.push_back(devices[i]);
    }
    *out = c;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_ctx_create(int n_gpus, sgpu_ctx **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SGPU_ERR_NO_DEVICE, "no CUDA device; libsolid_gpu has no CPU fallback");
    if (n_gpus <= 0) n_gpus = ndev;
    if (n_gpus > ndev) return fail(SGPU_ERR_INVALID_ARGUMENT, "ctx_create: %d GPUs asked, %d visible", n_gpus, ndev);
    std::vector<int> d(n_gpus);
    for (int i = 0; i < n_gpus; ++i) d[i] = i;
    return sgpu_ctx_create_devices(d.data(), n_gpus, out);
}

SGPU_EXPORT int sgpu_ctx_destroy(sgpu_ctx *c) {
    delete c;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_ctx_devices(const sgpu_ctx *c) { return c ? (int)c->devices.size() : 0; }

SGPU_EXPORT int sgpu_ctx_fir_create(sgpu_ctx *ctx, const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                    double scale_re, double scale_im, int is_decimator, size_t decimation,
                                    sgpu_sharded **out) {
    if (!ctx || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    sgpu_sharded *sh = new (std::nothrow) sgpu_sharded();
    if (!sh) return fail(SGPU_ERR_ALLOC, "out of host memory");
    sh->kind = kFir;
    sh->C = n_channels;
    sh->T = n_taps;
    sh->is_decim = is_decimator != 0;
    sh->M = sh->is_decim ? decimation : 1;
    if (n_channels == 1) {  // one stream: a one-channel handle on every device, time segments per call
        sh->by_stream = true;
        for (int d : ctx->devices) {
            Shard s;
            s.device = d;
            s.ch_first = 0;
            s.ch_count = 1;
            sh->shards.push_back(s);
        }
    } else {
        plan_channels(ctx, n_channels, sh->shards);
    }
    return build(ctx, sh, [&](Shard &s) {
        return sgpu_fir_create(taps, n_taps, kind, s.ch_count, scale_re, scale_im, is_decimator, decimation, (sgpu_fir **)&s.handle);
    }, out);
}

SGPU_EXPORT int sgpu_ctx_interp_create(sgpu_ctx *ctx, const double *taps, size_t n_taps, sgpu_tapkind kind,
                                       size_t n_channels, size_t interpolation, sgpu_sharded **out) {
    if (!ctx || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    sgpu_sharded *sh = new (std::nothrow) sgpu_sharded();
    if (!sh) return fail(SGPU_ERR_ALLOC, "out of host memory");
    sh->kind = kInterp;
    sh->C = n_channels;
    sh->T = n_taps;
    sh->L = interpolation;
    plan_channels(ctx, n_channels, sh->shards);
    return build(ctx, sh, [&](Shard &s) {
        return sgpu_interp_create(taps, n_taps, kind, s.ch_count, interpolation, (sgpu_interp **)&s.handle);
    }, out);
}

SGPU_EXPORT int sgpu_ctx_iir_create(sgpu_ctx *ctx, sgpu_iirtype type, const double *ff, size_t n_ff, const double *fb,
                                    size_t n_fb, size_t n_channels, sgpu_iirwrap wrap, size_t factor, sgpu_sharded **out) {
    if (!ctx || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    sgpu_sharded *sh = new (std::nothrow) sgpu_sharded();
    if (!sh) return fail(SGPU_ERR_ALLOC, "out of host memory");
    sh->kind = kIir;
    sh->C = n_channels;
    plan_channels(ctx, n_channels, sh->shards);  // one stream (C == 1) stays on the first device: the recurrence is serial
    return build(ctx, sh, [&](Shard &s) {
        return sgpu_iir_create(type, ff, n_ff, fb, n_fb, s.ch_count, wrap, factor, (sgpu_iir **)&s.handle);
    }, out);
}

SGPU_EXPORT int sgpu_sharded_destroy(sgpu_sharded *sh) {
    if (!sh) return SGPU_OK;
    for (auto &s : sh->shards) destroy_handle(sh->kind, s.handle);
    delete sh;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_sharded_shards(const sgpu_sharded *sh) { return sh ? (int)sh->shards.size() : 0; }
SGPU_EXPORT int sgpu_sharded_last_segments(const sgpu_sharded *sh) { return sh ? sh->last_segments : 0; }
SGPU_EXPORT int sgpu_sharded_shard_info(const sgpu_sharded *sh, int index, int *device, size_t *first_channel, size_t *n_channels) {
    if (!sh || index < 0 || index >= (int)sh->shards.size()) return fail(SGPU_ERR_INVALID_ARGUMENT, "shard index");
    if (device) *device = sh->shards[index].device;
    if (first_channel) *first_channel = sh->shards[index].ch_first;
    if (n_channels) *n_channels = sh->shards[index].ch_count;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_sharded_out_len(const sgpu_sharded *sh, size_t n_in) {
    if (!sh || sh->shards.empty()) return 0;
    void *h = sh->shards[0].handle;  // shard 0 carries the stream's counter; channel shards advance in lock-step
    switch (sh->kind) {
        case kFir: return sgpu_fir_out_len((sgpu_fir *)h, n_in);
        case kInterp: return n_in * sh->L;
        default: return sgpu_iir_out_len((sgpu_iir *)h, n_in);
    }
}

SGPU_EXPORT int sgpu_sharded_reset(sgpu_sharded *sh) {
    if (!sh) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    for (auto &s : sh->shards) {
        int st;
        switch (sh->kind) {
            case kFir: st = sgpu_fir_reset((sgpu_fir *)s.handle); break;
            case kInterp: st = sgpu_interp_reset((sgpu_interp *)s.handle); break;
            default: st = sgpu_iir_reset((sgpu_iir *)s.handle); break;
        }
        if (st) return st;
    }
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_sharded_execute_block(sgpu_sharded *sh, const float *in, size_t n_in, size_t in_stride, float *out,
                                           size_t out_stride, size_t *n_out_p) {
    if (!sh) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    const size_t n_out = sgpu_sharded_out_len(sh, n_in);
    if (n_out_p) *n_out_p = n_out;
    sh->last_segments = 0;
    if (n_in == 0) return SGPU_OK;
    if (!in || (n_out && !out)) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (sh->C > 1 && in_stride < n_in) return fail(SGPU_ERR_INVALID_ARGUMENT, "in_stride < n_in");
    if (out_stride < n_out) return fail(SGPU_ERR_CAPACITY, "out capacity %zu < %zu outputs", out_stride, n_out);

    if (!sh->by_stream) {  // ---- channel ranges: every shard filters its rows of the caller's buffers
        sh->last_segments = (int)sh->shards.size();
        return for_each_shard(sh->shards.size(), sh->shards, [&](size_t i) -> int {
            const Shard &s = sh->shards[i];
            const float *xi = in + 2 * s.ch_first * in_stride;
            float *yi = out + 2 * s.ch_first * out_stride;
            size_t got = 0;
            switch (sh->kind) {
                case kFir: return sgpu_fir_execute_block((sgpu_fir *)s.handle, xi, n_in, in_stride, yi, out_stride, &got, SGPU_HOST, nullptr);
                case kInterp: return sgpu_interp_execute_block((sgpu_interp *)s.handle, xi, n_in, in_stride, yi, out_stride, &got, SGPU_HOST, nullptr);
                default: return sgpu_iir_execute_block((sgpu_iir *)s.handle, xi, n_in, in_stride, yi, out_stride, &got, SGPU_HOST, nullptr);
            }
        });
    }

    // ---- one FIR stream: time segments.  Invariant between calls: shard 0 holds the stream's history and counter.
    sgpu_fir *f0 = (sgpu_fir *)sh->shards[0].handle;
    const size_t H = sh->T > 0 ? sh->T - 1 : 0, M = sh->M;
    uint64_t c0 = 0;
    int st = sgpu_fir_get_state(f0, nullptr, &c0);
    if (st) return st;
    // first sample at which the decimator's counter is 0 again (decim.rs:221-228: emits when (count + 1) % M == 0)
    const size_t off = sh->is_decim ? (size_t)((M - c0 % M) % M) : 0;
    // segments: shard 0 takes [0, first_1); the rest of the input is cut at multiples of M from `off` on.  Every segment must
    // be at least as long as the halo (it is sliced from THIS call's input) and worth a launch: else fewer segments.
    const size_t min_seg = std::max<size_t>(std::max<size_t>(H, M), (size_t)1 << 16);
    int world = (int)sh->shards.size();
    while (world > 1 && (n_in <= off || (n_in - off) / (size_t)world < min_seg)) --world;
    struct Seg {
        size_t first, count, out_first;
    };
    std::vector<Seg> segs((size_t)world);
    for (int r = 0; r < world; ++r) {
        size_t first = 0, count = 0;
        if (world == 1) {
            first = 0;
            count = n_in;
        } else {
            sgpu_shard_stream(n_in - off, M, world, r, &first, &count);
            first += off;
            if (r == 0) {  // shard 0 also takes the samples in front of the first aligned position
                count += first;
                first = 0;
            }
        }
        segs[(size_t)r] = {first, count, sh->is_decim ? (size_t)((c0 + first) / M) : first};
    }
    sh->last_segments = world;
    st = for_each_shard((size_t)world, sh->shards, [&](size_t i) -> int {
        sgpu_fir *f = (sgpu_fir *)sh->shards[i].handle;
        const Seg &sg = segs[i];
        if (i > 0) {  // fresh filter + the halo as its history; the halo does not move the decimator's counter
            int s2 = sgpu_fir_reset(f);
            if (s2 == SGPU_OK && H > 0) s2 = sgpu_fir_write(f, in + 2 * (sg.first - H), H, H, SGPU_HOST, nullptr);
            if (s2 == SGPU_OK) s2 = sgpu_fir_set_state(f, nullptr, 0);
            if (s2) return s2;
        }
        size_t got = 0;
        return sgpu_fir_execute_block(f, in + 2 * sg.first, sg.count, sg.count, out + 2 * sg.out_first,
                                      out_stride - sg.out_first, &got, SGPU_HOST, nullptr);
    });
    if (st) return st;
    if (world > 1) {  // hand the stream's end state back to shard 0: the last T-1 samples and the counter
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(sh->shards[0].device);
        if (H > 0) st = sgpu_fir_write(f0, in + 2 * (n_in - H), H, H, SGPU_HOST, nullptr);
        if (st == SGPU_OK) st = sgpu_fir_set_state(f0, nullptr, (c0 + n_in) % M);
        cudaSetDevice(prev);
    }
    return st;
}
