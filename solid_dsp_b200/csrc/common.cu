// Library-level entry points: errors, device probe, launch counter, sharding arithmetic.
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <thread>

#include "sgpu_common.cuh"

namespace sgpu {

std::atomic<uint64_t> g_launches{0};

char *err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return status;
}

int require_device(int *device_out, int *sm_count_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SGPU_ERR_NO_DEVICE, "no CUDA device (%s); libsolid_gpu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = 0;
    SGPU_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SGPU_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major < 10)
        return fail(SGPU_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    dev, p.major, p.minor);
    if (device_out) *device_out = dev;
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    return SGPU_OK;
}

int Staging::ensure(size_t need_in, size_t need_out) {
    if (need_in > in_bytes) {
        if (in) cudaFree(in);
        in = nullptr;
        in_bytes = 0;
        SGPU_CUDA(cudaMalloc(&in, need_in));
        in_bytes = need_in;
    }
    if (need_out > out_bytes) {
        if (out) cudaFree(out);
        out = nullptr;
        out_bytes = 0;
        SGPU_CUDA(cudaMalloc(&out, need_out));
        out_bytes = need_out;
    }
    return SGPU_OK;
}

void Staging::release() {
    if (in) cudaFree(in);
    if (out) cudaFree(out);
    in = out = nullptr;
    in_bytes = out_bytes = 0;
}

int HostPipe::ensure(size_t need_in, size_t need_out) {
    if (!s_in) {
        SGPU_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        SGPU_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SGPU_CUDA(cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming));
            SGPU_CUDA(cudaEventCreateWithFlags(&e_comp[i], cudaEventDisableTiming));
            SGPU_CUDA(cudaEventCreateWithFlags(&e_out[i], cudaEventDisableTiming));
        }
    }
    if (need_in > in_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            d_in[i] = nullptr;
        }
        in_bytes = 0;
        for (int i = 0; i < 2; ++i) SGPU_CUDA(cudaMalloc(&d_in[i], need_in));
        in_bytes = need_in;
    }
    if (need_out > out_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (d_out[i]) cudaFree(d_out[i]);
            d_out[i] = nullptr;
        }
        out_bytes = 0;
        for (int i = 0; i < 2; ++i) SGPU_CUDA(cudaMalloc(&d_out[i], need_out));
        out_bytes = need_out;
    }
    return SGPU_OK;
}

int HostPipe::ensure_staging(size_t need_in, size_t need_out) {
    if (need_in > h_in_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (h_in[i]) cudaFreeHost(h_in[i]);
            h_in[i] = nullptr;
        }
        h_in_bytes = 0;
        for (int i = 0; i < 2; ++i) SGPU_CUDA(cudaHostAlloc(&h_in[i], need_in, cudaHostAllocDefault));
        h_in_bytes = need_in;
    }
    if (need_out > h_out_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (h_out[i]) cudaFreeHost(h_out[i]);
            h_out[i] = nullptr;
        }
        h_out_bytes = 0;
        for (int i = 0; i < 2; ++i) SGPU_CUDA(cudaHostAlloc(&h_out[i], need_out, cudaHostAllocDefault));
        h_out_bytes = need_out;
    }
    return SGPU_OK;
}

void HostPipe::release() {
    for (int i = 0; i < 2; ++i) {
        if (h_in[i]) cudaFreeHost(h_in[i]);
        if (h_out[i]) cudaFreeHost(h_out[i]);
        h_in[i] = h_out[i] = nullptr;
    }
    h_in_bytes = h_out_bytes = 0;
    for (int i = 0; i < 2; ++i) {
        if (d_in[i]) cudaFree(d_in[i]);
        if (d_out[i]) cudaFree(d_out[i]);
        if (e_in[i]) cudaEventDestroy(e_in[i]);
        if (e_comp[i]) cudaEventDestroy(e_comp[i]);
        if (e_out[i]) cudaEventDestroy(e_out[i]);
        d_in[i] = d_out[i] = nullptr;
        e_in[i] = e_comp[i] = e_out[i] = nullptr;
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    s_in = s_out = nullptr;
    in_bytes = out_bytes = 0;
}

bool host_staging_enabled() {
    const char *e = getenv("SGPU_HOST_STAGING");
    return e ? atoi(e) != 0 : true;  // measured: 1.56 vs 0.42 Gsamp/s on a 512-tap stream of 2^27 pageable samples
}

bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void par_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows) {
    const size_t total = width * rows;
    if (total == 0) return;
    // bytes [lo, hi) of the rows x width index space
    auto copy_range = [=](size_t lo, size_t hi) {
        size_t r = lo / width, off = lo - r * width;
        while (lo < hi) {
            const size_t n = std::min(width - off, hi - lo);
            memcpy(static_cast<char *>(dst) + r * dst_pitch + off, static_cast<const char *>(src) + r * src_pitch + off, n);
            lo += n;
            ++r;
            off = 0;
        }
    };
    unsigned hc = std::thread::hardware_concurrency();
    // measured on the GPU box (16 host threads; tools/pageable_probe.py, 2^27 pageable samples in, a fresh array out):
    // 1 thread 0.47, 4 threads 1.14, 8 threads 1.5, 16 threads 1.80 Gsamp/s (the driver's own staged copies: 0.42)
    size_t nt = total < ((size_t)4 << 20) ? 1 : std::min<size_t>(16, std::max<unsigned>(1, hc));
    if (const char *e = getenv("SGPU_HOST_COPY_THREADS")) nt = std::max(1, atoi(e));
    if (nt <= 1) {
        copy_range(0, total);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    const size_t per = (total + nt - 1) / nt;
    for (size_t t = 1; t < nt; ++t) {
        const size_t lo = std::min(total, t * per), hi = std::min(total, (t + 1) * per);
        if (lo < hi) th.emplace_back(copy_range, lo, hi);
    }
    copy_range(0, std::min(total, per));
    for (auto &x : th) x.join();
}

size_t host_chunk_len(size_t C, size_t n_in) {
    // ~64 MiB of input per chunk: large enough for PCIe efficiency and full-chip kernels, small
    // enough that the first copy-in and the last copy-out (the un-overlapped ends) stay short
    size_t target = (size_t)64 << 20;
    const char *e = getenv("SGPU_HOST_CHUNK_MB");
    if (e && atoi(e) > 0) target = (size_t)atoi(e) << 20;
    size_t chunk = target / 8 / (C ? C : 1);
    if (chunk < 4096) chunk = 4096;
    chunk = (chunk + 63) / 64 * 64;
    return chunk < n_in ? chunk : (n_in ? n_in : 1);
}

}  // namespace sgpu

// ---------------------------------------------------------------------------------------------
// Pinned host memory on the NUMA node of a GPU.  The SGPU_HOST path is bounded by the host <-> device copies; on a
// multi-socket box a pinned buffer that lives on the other socket costs a hop over the inter-socket link per byte
// (tools/pcie_probe.py measures both placements).  The pages are first-touched while the calling thread is bound to
// the CPUs the kernel reports as local to the GPU's PCIe root (/sys/bus/pci/devices/<bdf>/local_cpulist).
namespace {
bool gpu_local_cpus(int device, cpu_set_t *set) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) return false;
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bdf);
    FILE *fp = fopen(path, "r");
    if (!fp) return false;
    char buf[4096] = {0};
    const size_t got = fread(buf, 1, sizeof(buf) - 1, fp);
    fclose(fp);
    if (got == 0) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) {
            for (int i = a; i <= b && i < CPU_SETSIZE; ++i) { CPU_SET(i, set); ++n; }
        } else if (sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) {
            CPU_SET(a, set);
            ++n;
        }
    }
    return n > 0;
}
}  // namespace

SGPU_EXPORT int sgpu_host_alloc(size_t bytes, int device, void **out) {
    if (!out) return sgpu::fail(SGPU_ERR_INVALID_ARGUMENT, "host_alloc: out is NULL");
    *out = nullptr;
    if (bytes == 0) return SGPU_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return sgpu::fail(SGPU_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0) cudaGetDevice(&device);
    cpu_set_t old_set, gpu_set;
    const bool have_old = sched_getaffinity(0, sizeof(old_set), &old_set) == 0;
    const bool bound = have_old && gpu_local_cpus(device, &gpu_set) && sched_setaffinity(0, sizeof(gpu_set), &gpu_set) == 0;
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) {
        // first touch under the GPU-local binding (cudaHostAlloc usually faults the pages in itself; this makes sure)
        volatile char *c = static_cast<volatile char *>(p);
        for (size_t i = 0; i < bytes; i += 4096) c[i] = 0;
    }
    if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return sgpu::fail(SGPU_ERR_ALLOC, "cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    *out = p;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_host_free(void *p) {
    if (!p) return SGPU_OK;
    if (cudaFreeHost(p) != cudaSuccess) {
        (void)cudaGetLastError();
        return sgpu::fail(SGPU_ERR_CUDA, "cudaFreeHost failed");
    }
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_abi_version(void) { return SGPU_ABI_VERSION; }

SGPU_EXPORT const char *sgpu_last_error(void) { return sgpu::err_buf(); }

SGPU_EXPORT const char *sgpu_status_name(int s) {
    switch (s) {
        case SGPU_OK: return "SGPU_OK";
        case SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO: return "FIRErrorCode::CoefficientsLengthZero";
        case SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE: return "FIRErrorCode::DecimationLessThanOne";
        case SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE: return "FIRErrorCode::InterpolationLessThanOne";
        case SGPU_ERR_FIR_NOT_ENOUGH_FILTERS: return "FIRErrorCode::NotEnoughFilters";
        case SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO: return "IIRErrorCode::NumeratorLengthZero";
        case SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO: return "IIRErrorCode::DenominatorLengthZero";
        case SGPU_ERR_IIR_SOS_SIZE_ZERO: return "IIRErrorCode::SecondOrderSectionSizeZero";
        case SGPU_ERR_IIR_SOS_SIZE_MISMATCH: return "IIRErrorCode::SecondOrderSectionSizeMismatch";
        case SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3: return "IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3";
        case SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE: return "IIRErrorCode::DecimationLessThanOne";
        case SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE: return "IIRErrorCode::InterpolationLessThanOne";
        case SGPU_ERR_SOS_COEFFICIENTS_NOT_IN_RANGE: return "SecondOrderErrorCode::CoefficientsNotInRange";
        case SGPU_ERR_FIRDES_BANDWIDTH: return "FirdesErrorCode::Bandwidth";
        case SGPU_ERR_FIRDES_STOP_BAND_LEVEL: return "FirdesErrorCode::StopBandLevel";
        case SGPU_ERR_FIRDES_MU: return "FirdesErrorCode::Mu";
        case SGPU_ERR_INVALID_ARGUMENT: return "SGPU_ERR_INVALID_ARGUMENT";
        case SGPU_ERR_CAPACITY: return "SGPU_ERR_CAPACITY";
        case SGPU_ERR_CUDA: return "SGPU_ERR_CUDA";
        case SGPU_ERR_UNSUPPORTED: return "SGPU_ERR_UNSUPPORTED";
        case SGPU_ERR_NO_DEVICE: return "SGPU_ERR_NO_DEVICE";
        case SGPU_ERR_ALLOC: return "SGPU_ERR_ALLOC";
        default: return "SGPU_ERR_UNKNOWN";
    }
}

SGPU_EXPORT int sgpu_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor,
                                 size_t *total_mem) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return sgpu::fail(SGPU_ERR_NO_DEVICE, "no CUDA device (%s)",
                          e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = 0;
    SGPU_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SGPU_CUDA(cudaGetDeviceProperties(&p, dev));
    if (device) *device = dev;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return SGPU_OK;
}

SGPU_EXPORT uint64_t sgpu_launch_count(void) {
    return sgpu::g_launches.load(std::memory_order_relaxed);
}

SGPU_EXPORT int sgpu_shard_channels(size_t n_channels, int world, int rank, size_t *first,
                                    size_t *count) {
    if (world < 1 || rank < 0 || rank >= world || !first || !count)
        return sgpu::fail(SGPU_ERR_INVALID_ARGUMENT, "shard_channels: bad world/rank");
    size_t base = n_channels / (size_t)world, rem = n_channels % (size_t)world;
    size_t r = (size_t)rank;
    *count = base + (r < rem ? 1 : 0);
    *first = r * base + (r < rem ? r : rem);
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_shard_stream(size_t n_samples, size_t align, int world, int rank,
                                  size_t *first, size_t *count) {
    if (world < 1 || rank < 0 || rank >= world || !first || !count || align < 1)
        return sgpu::fail(SGPU_ERR_INVALID_ARGUMENT, "shard_stream: bad world/rank/align");
    // segment boundaries at multiples of `align`; the last rank takes the ragged tail
    size_t units = n_samples / align;
    size_t base = units / (size_t)world, rem = units % (size_t)world;
    size_t r = (size_t)rank;
    size_t u0 = r * base + (r < rem ? r : rem);
    size_t u1 = u0 + base + (r < rem ? 1 : 0);
    *first = u0 * align;
    size_t end = (rank == world - 1) ? n_samples : u1 * align;
    *count = end - *first;
    return SGPU_OK;
}
