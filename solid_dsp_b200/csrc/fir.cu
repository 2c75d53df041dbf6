// FIRFilter / DecimatingFIRFilter / InterpolatingFIRFilter (PolyPhaseFilterBank) on sm_100a.
//
// Reference loops replaced (relative to the reference's src/):
//   FIRFilter::execute_block            filter/fir/mod.rs:235-241  -> :209-212
//   DecimatingFIRFilter::execute_block  filter/fir/decim.rs:250-256 -> :221-228
//   InterpolatingFIRFilter::execute_block filter/fir/interp.rs:102-111 -> pfb.rs:85-90
//   Window::push / to_vec               window/mod.rs:63-71,44-51  (history: hist_update_kernel)
//   DotProduct::execute                 dot_product/mod.rs:159-170 (fir_core.cuh)
#include <map>
#include <mutex>

#include "fir_core.cuh"
#include "fir_tc.cuh"
#include "nco.cuh"
#include "sgpu_common.cuh"

namespace sgpu {
namespace {

constexpr int kMaxSmem = 227 * 1024;
constexpr size_t kMaxGridY = 65535;

struct FirArgs {
    const float2 *in;
    float2 *out;
    const float2 *hist;  // [C][T-1], oldest first (state entering this call)
    float2 *hist_new;    // [C][T-1]: state leaving this call, written by the last block of every channel (NULL: not here)
    const float *taps;   // [nsets][Qpad] taps (x2 floats for complex taps), device; channel c reads taps + c * tap_stride
    long long tap_stride;  // floats between the tap images of two channels (0: all channels share one image)
    long long in_stride, out_stride;
    long long n_in, n_out;  // per channel
    int T;                  // taps of the full filter
    int M;                  // decimation (decim kernel) or interpolation L (interp kernel)
    int c0;                 // decimator phase counter on entry (current_item)
    int Qpad;               // taps per phase, padded to a multiple of R
    int RS;                 // row stride of a plane in float4, odd
    int vec_in, vec_out;    // 16-byte vector access allowed
    float scale_re, scale_im;
    // NCO mix-down fused in front of the filter (DDC, nco/mod.rs:141-172): sample i of this call on channel c is
    // multiplied by conj(phasor(theta0[c] + (nco_pos + i) * delta[c])) before the filter sees it (32-bit wrapping phase,
    // nco/mod.rs:93-96).  nco = [C][2] words (theta0, delta_theta); lut = 1024 x (cos, sin); NULL = no mixing
    const unsigned *nco;
    unsigned nco_pos;
    const float2 *lut;
};

// phase of sample 0 of this call and the phase step, channel ch
__device__ __forceinline__ void nco_channel(const FirArgs &a, const int ch, unsigned &theta, unsigned &delta) {
    delta = a.nco[2 * ch + 1];
    theta = a.nco[2 * ch] + a.nco_pos * delta;
}

// NCO phasor of a 32-bit phase (nco/mod.rs:98-114): table index = ((theta + 2^21) >> 22) & 1023, sin = table[index],
// cos = table[(index + 256) & 1023]; the table here holds the (cos, sin) pair per index
__device__ __forceinline__ float2 nco_mix_down(const float2 x, const unsigned theta, const float2 *__restrict__ lut) {
    const float2 cs = lut[(theta + (1u << 21)) >> 22];
    // conj(cos + j sin) * x  (nco/mod.rs:147-151)
    return make_float2(fmaf(cs.x, x.x, cs.y * x.y), fmaf(cs.x, x.y, -cs.y * x.x));
}

// sample `i` of this call's logical input: i < 0 reads the history, beyond either end is 0
__device__ __forceinline__ float2 fetch_sample(const float2 *__restrict__ x,
                                               const float2 *__restrict__ hist, const long long i,
                                               const long long n_in, const int T) {
    if (i >= 0) return i < n_in ? x[i] : make_float2(0.f, 0.f);
    const long long h = (long long)(T - 1) + i;
    return h >= 0 ? hist[h] : make_float2(0.f, 0.f);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// new history = last T-1 samples of (old history ++ x[0..n_in))   (Window::push, window/mod.rs:63-71), written by the
// LAST block of every channel's grid row before it starts on its tile: execute_block is one launch.  The old and the
// new tail are different buffers (ping-pong), so the blocks that still read the old one are not disturbed.
__device__ __forceinline__ void hist_tail_update(const FirArgs &a, const int ch, const int tid, const int nthr) {
    if (a.hist_new == nullptr || blockIdx.x != gridDim.x - 1) return;
    const int H = a.T - 1;
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    const float2 *__restrict__ ho = a.hist + (long long)ch * H;
    float2 *__restrict__ hn = a.hist_new + (long long)ch * H;
    unsigned nco_theta = 0, nco_delta = 0;
    if (a.lut) nco_channel(a, ch, nco_theta, nco_delta);
    for (int i = tid; i < H; i += nthr) {
        const long long s = a.n_in - H + i;
        float2 v;
        if (s >= 0) {
            v = x[s];
            if (a.lut) v = nco_mix_down(v, nco_theta + (unsigned)s * nco_delta, a.lut);  // the history holds MIXED samples
        } else {
            const long long h = (long long)H + s;
            v = h >= 0 ? ho[h] : make_float2(0.f, 0.f);
        }
        hn[i] = v;
    }
}

// --------------------------------------------------------------------------------------------
// FIR (M1 = true) and decimating FIR.  Block = NT threads, tile = NT*R outputs of one channel.
// Output m of this call is produced by input n_m = m*M + (M-1-c0); with k = q*M + p,
//   y[m] = scale * sum_p sum_q g[q*M+p] * x[(m-q)*M + (M-1-c0) - p]
// i.e. M short FIRs (taps g_p[q]) over the phase sequences x_p[m'] = x[m'*M + (M-1-c0) - p],
// which the loader de-interleaves into M shared-memory planes.
//
// PS = phase split: PS adjacent lanes share one run of R outputs and each walks M/PS of the phase
// planes; their partial sums meet in a warp-shuffle butterfly.  This keeps R = 16 (one LDS.128 per
// 16 complex MACs) while a tile needs only NT/PS * R * M input samples of shared memory, so 3-6
// blocks of 4 warps stay resident per SM and tile loads overlap other blocks' arithmetic.
template <int R, bool PACKED, bool M1, int NT, int MINB, int PS, bool CT = false>
__global__ void __launch_bounds__(NT, MINB) fir_decim_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    static_assert(M1 ? PS == 1 : true, "the plain FIR has a single phase");
    const int tid = threadIdx.x;
    const int M = M1 ? 1 : a.M;
    const int Qpad = a.Qpad;
    const int HR = Qpad / R;
    constexpr int OT = NT / PS;  // output-owning thread groups per block
    const int rows = HR + OT;
    const int RS = a.RS;
    const int plane_f4 = (R / 2) * RS + 1;  // +1: consecutive planes are skewed by 16 bytes
    constexpr int TW = CT ? 2 : 1;
    constexpr int NACC = CT ? 2 * R : R;
    float *taps_s = reinterpret_cast<float *>(smem + (size_t)M * plane_f4);

    const int ch = blockIdx.y;
    const long long m_base = (long long)blockIdx.x * (OT * R);
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
    hist_tail_update(a, ch, tid, NT);

    {  // taps image -> shared memory
        const int n4 = M * (Qpad * TW + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps + (long long)ch * a.tap_stride);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }

    // ---- tile load: rows*R*M consecutive input samples starting at i_lo, two per thread-step
    const long long i_lo = (m_base - Qpad) * M - a.c0;
    const int total_pairs = rows * R * M / 2;
    if constexpr (M1) {
        // asynchronous (LDGSTS) for everything inside the input; history / edges take the guarded path
        for (int pe = tid; pe < total_pairs; pe += NT) {
            const long long i = i_lo + 2 * pe;
            const int rho = pe / (R / 2), jj = pe % (R / 2);
            float4 *dst = smem + jj * RS + rho;
            if (i >= 0 && i + 1 < a.n_in) {
                if (a.vec_in) {
                    cp_async16(dst, x + i);
                } else {
                    cp_async8(dst, x + i);
                    cp_async8(reinterpret_cast<float2 *>(dst) + 1, x + i + 1);
                }
            } else {
                const float2 s0 = fetch_sample(x, hist, i, a.n_in, a.T);
                const float2 s1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
                *dst = make_float4(s0.x, s0.y, s1.x, s1.y);
            }
        }
    } else if (NT % (R * M) == 0) {
        // Fast de-interleave: R*M consecutive input samples form one row of every phase plane and
        // NT is a multiple of that, so a thread keeps its (phase, position-in-row) for the whole
        // tile and only walks down the rows: one LDGSTS.64 plus two adds per sample.
        const int rm = R * M;
        const int e_in = tid % rm;
        const int rem = e_in % M, j = e_in / M;
        const int row_step = NT / rm;
        const int p = M - 1 - rem;
        float2 *dst = reinterpret_cast<float2 *>(smem + (size_t)p * plane_f4 + (j >> 1) * RS + tid / rm) + (j & 1);
        const long long total = (long long)rows * rm;
        const bool interior = i_lo >= 0 && i_lo + total <= a.n_in;
        if (interior) {
            const float2 *src = x + i_lo + tid;
            for (int rho = tid / rm; rho < rows; rho += row_step) {
                cp_async8(dst, src);
                dst += 2 * row_step;  // next row = next float4 of the plane
                src += NT;
            }
        } else {
            long long i = i_lo + tid;
            for (int rho = tid / rm; rho < rows; rho += row_step) {
                if (i >= 0 && i < a.n_in) cp_async8(dst, x + i);
                else *dst = fetch_sample(x, hist, i, a.n_in, a.T);
                dst += 2 * row_step;
                i += NT;
            }
        }
    } else {
        int e = 2 * tid;
        int q = e / M, rem = e - q * M;
        const int qs = (2 * NT) / M, rs = (2 * NT) - qs * M;
        for (int pe = tid; pe < total_pairs; pe += NT) {
            const long long i = i_lo + 2 * pe;
            float2 *d0, *d1;
            {
                const int p = M - 1 - rem, rho = q / R, j = q % R;
                d0 = reinterpret_cast<float2 *>(smem + (size_t)p * plane_f4 + (j >> 1) * RS + rho) + (j & 1);
            }
            {
                int q1 = q, rem1 = rem + 1;
                if (rem1 == M) { rem1 = 0; q1++; }
                const int p = M - 1 - rem1, rho = q1 / R, j = q1 % R;
                d1 = reinterpret_cast<float2 *>(smem + (size_t)p * plane_f4 + (j >> 1) * RS + rho) + (j & 1);
            }
            if (i >= 0 && i + 1 < a.n_in) {  // the de-interleave rides on 8-byte LDGSTS
                cp_async8(d0, x + i);
                cp_async8(d1, x + i + 1);
            } else {
                *d0 = fetch_sample(x, hist, i, a.n_in, a.T);
                *d1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
            }
            q += qs;
            rem += rs;
            if (rem >= M) { rem -= M; q++; }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- compute
    float2 acc[NACC];
#pragma unroll
    for (int r = 0; r < NACC; ++r) acc[r] = make_float2(0.f, 0.f);
    const int ot = tid / PS, part = tid % PS;
    const int row0 = HR + ot;
    const int nchunks = Qpad / R;
    const int Mp = (M + PS - 1) / PS;  // phases per lane of a split group (contiguous block)
    for (int sidx = 0; sidx < Mp; ++sidx) {
        const int p = part * Mp + sidx;
        if (p < M)
            fir_core<R, PACKED, CT>(acc, smem + (size_t)p * plane_f4, RS, row0,
                                    taps_s + (size_t)p * (Qpad * TW + kTapSkew), nchunks);
    }
    if constexpr (PS > 1) {  // butterfly over the PS lanes of a group: everyone ends with the full sums
#pragma unroll
        for (int o = 1; o < PS; o <<= 1) {
#pragma unroll
            for (int r = 0; r < NACC; ++r) {
                acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, o);
                acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, o);
            }
        }
    }
    // ---- epilogue: scale (fir/mod.rs:211: Out * Coef), stage through plane 0, coalesced store
    float2 yv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if constexpr (CT) {
            // (a+bi)(c+di) sums: re = sum(ac) - sum(bd), im = sum(ad) + sum(bc); then the complex scale
            const float yr = acc[r].x - acc[R + r].y, yi = acc[r].y + acc[R + r].x;
            yv[r] = make_float2(yr * a.scale_re - yi * a.scale_im, yr * a.scale_im + yi * a.scale_re);
        } else {
            yv[r] = make_float2(acc[r].x * a.scale_re, acc[r].y * a.scale_re);
        }
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < R / 2; ++jj)
        if (jj / (R / 2 / PS) == part)  // each lane of a group stages its share of the run
            smem[jj * RS + ot] = make_float4(yv[2 * jj].x, yv[2 * jj].y, yv[2 * jj + 1].x, yv[2 * jj + 1].y);
    __syncthreads();
    float2 *__restrict__ y = a.out + (long long)ch * a.out_stride;
    for (int idx = tid; idx < OT * R / 2; idx += NT) {
        const int rho = idx / (R / 2), jj = idx % (R / 2);
        const long long o = m_base + (long long)rho * R + 2 * jj;
        if (o >= a.n_out) continue;
        const float4 v = smem[jj * RS + rho];
        if (a.vec_out && o + 1 < a.n_out) {
            *reinterpret_cast<float4 *>(y + o) = v;
        } else {
            y[o] = make_float2(v.x, v.y);
            if (o + 1 < a.n_out) y[o + 1] = make_float2(v.z, v.w);
        }
    }
}

// --------------------------------------------------------------------------------------------
// Interpolator: y[n*L + p] = sum_{j<S} hp[p][j] * x[n-j], hp[p][j] = hpad[p + (S-1-j)*L].
// One input plane, L tap sets; a thread runs the L phases one after the other over the same
// R input positions and stages the interleaved outputs in shared memory.  a.M carries L.
// PS adjacent lanes share one run of R input positions and split the L output phases.
template <int R, bool PACKED, int NT, int MINB, int PS, bool CT = false>
__global__ void __launch_bounds__(NT, MINB) fir_interp_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    const int tid = threadIdx.x;
    const int L = a.M;
    const int Qpad = a.Qpad;
    const int HR = Qpad / R;
    constexpr int OT = NT / PS;
    const int rows = HR + OT;
    const int RS = a.RS;
    const int plane_f4 = (R / 2) * RS + 1;
    constexpr int TW = CT ? 2 : 1;
    constexpr int NACC = CT ? 2 * R : R;
    float *taps_s = reinterpret_cast<float *>(smem + plane_f4);
    // staging: NT*R*L outputs, thread t's run skewed by t float2 (bank spread)
    float2 *stage = reinterpret_cast<float2 *>(taps_s + (size_t)L * (Qpad * TW + kTapSkew));

    const int ch = blockIdx.y;
    const long long n_base = (long long)blockIdx.x * (OT * R);
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);  // a.T - 1 = S samples kept
    hist_tail_update(a, ch, tid, NT);

    {
        const int n4 = L * (Qpad * TW + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps + (long long)ch * a.tap_stride);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }
    const long long i_lo = n_base - Qpad;
    const int total_pairs = rows * R / 2;
    for (int pe = tid; pe < total_pairs; pe += NT) {
        const long long i = i_lo + 2 * pe;
        const int rho = pe / (R / 2), jj = pe % (R / 2);
        float4 *dst = smem + jj * RS + rho;
        if (i >= 0 && i + 1 < a.n_in) {
            if (a.vec_in) {
                cp_async16(dst, x + i);
            } else {
                cp_async8(dst, x + i);
                cp_async8(reinterpret_cast<float2 *>(dst) + 1, x + i + 1);
            }
        } else {
            const float2 s0 = fetch_sample(x, hist, i, a.n_in, a.T);
            const float2 s1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
            *dst = make_float4(s0.x, s0.y, s1.x, s1.y);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int ot = tid / PS, part = tid % PS;
    const int row0 = HR + ot;
    const int nchunks = Qpad / R;
    float2 *my = stage + (size_t)ot * (R * L + 1);
    const int Lp = (L + PS - 1) / PS;
    for (int sidx = 0; sidx < Lp; ++sidx) {
        const int p = part * Lp + sidx;
        if (p >= L) break;
        float2 acc[NACC];
#pragma unroll
        for (int r = 0; r < NACC; ++r) acc[r] = make_float2(0.f, 0.f);
        fir_core<R, PACKED, CT>(acc, smem, RS, row0, taps_s + (size_t)p * (Qpad * TW + kTapSkew), nchunks);
#pragma unroll
        for (int r = 0; r < R; ++r) {  // no scale: pfb.rs:85-90
            if constexpr (CT) my[r * L + p] = make_float2(acc[r].x - acc[R + r].y, acc[r].y + acc[R + r].x);
            else my[r * L + p] = acc[r];
        }
    }
    __syncthreads();
    float2 *__restrict__ y = a.out + (long long)ch * a.out_stride;
    const int per_thread = R * L;
    const long long o_base = n_base * L;
    // warp w drains the runs of groups w, w+NT/32, ...: 32 lanes stride over one run's R*L outputs
    for (int g = tid >> 5; g < OT; g += NT / 32) {
        const float2 *run = stage + (size_t)g * (per_thread + 1);
        const long long og = o_base + (long long)g * per_thread;
        for (int k = tid & 31; k < per_thread; k += 32)
            if (og + k < a.n_out) y[og + k] = run[k];
    }
}

// new history = last T-1 samples of (old history ++ x[0..n_in))   (Window::push, window/mod.rs:63-71)
__global__ void hist_update_kernel(const float2 *__restrict__ in, long long in_stride, long long n_in,
                                   const float2 *__restrict__ hist_old, float2 *__restrict__ hist_new,
                                   int H /* = T-1 */) {
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H) return;
    const long long s = n_in - H + i;
    float2 v;
    if (s >= 0) v = in[(long long)ch * in_stride + s];
    else {
        const long long h = (long long)H + s;  // index into old history
        v = h >= 0 ? hist_old[(long long)ch * H + h] : make_float2(0.f, 0.f);
    }
    hist_new[(long long)ch * H + i] = v;
}

// PolyPhaseFilterBank::execute(index): one dot product per channel over the history ++ nothing.
// hist holds the last S-1 samples; the window's newest element is hist[S-2] ... the PFB window
// has capacity S, so the bank keeps S samples: we store S-1 "history" plus the newest in `last`.
__global__ void pfb_phase_kernel(const float2 *__restrict__ hist, int S, const float *__restrict__ taps,
                                 long long tap_stride, int Qpad, int tw, int phase, float2 *__restrict__ out, int C) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= C) return;
    // window (newest first) = hist[S-1], hist[S-2], ..., hist[0]  with hist of S samples
    const float2 *h = hist + (long long)ch * S;
    const float *g = taps + (long long)ch * tap_stride + (size_t)phase * (Qpad * tw + kTapSkew);
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < S; ++j) {
        const float2 w = h[S - 1 - j];
        if (tw == 2) {  // complex taps
            const float ga = g[2 * j], gb = g[2 * j + 1];
            acc.x = fmaf(ga, w.x, fmaf(-gb, w.y, acc.x));
            acc.y = fmaf(ga, w.y, fmaf(gb, w.x, acc.y));
        } else {
            const float gj = g[j];
            acc.x = fmaf(gj, w.x, acc.x);
            acc.y = fmaf(gj, w.y, acc.y);
        }
    }
    out[ch] = acc;
}

// --------------------------------------------------------------------------------------------
// Direct form, one thread per output: the path for shapes whose tile does not fit the 227 KB of shared memory (the
// reference takes any decimation / interpolation factor and any length: fir/decim.rs:27-42, fir/interp.rs:27-54),
// e.g. 256 taps at M = 64.  Sequential dot product in the reference's order (dot_product/mod.rs:159-170) over the
// tap image in global memory.  INTERP = false: y[m] = scale * sum_k g[k] x[m M + (M-1-c0) - k], g[q M + p] at
// image[p][q];  INTERP = true: y[n L + p] = sum_j image[p][j] x[n - j] (a.M = L, a.T - 1 = S).
template <bool INTERP, bool CT>
__global__ void __launch_bounds__(128) fir_direct_kernel(const FirArgs a) {
    constexpr int TW = CT ? 2 : 1;
    const int ch = blockIdx.y;
    hist_tail_update(a, ch, threadIdx.x, 128);
    const long long o = (long long)blockIdx.x * 128 + threadIdx.x;
    if (o >= a.n_out) return;
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
    const float *__restrict__ img = a.taps + (long long)ch * a.tap_stride;
    const int rs = a.Qpad * TW + kTapSkew;
    float yr = 0.f, yi = 0.f;
    auto fetch = [&](const long long i) { return fetch_sample(x, hist, i, a.n_in, a.T); };
    auto mac = [&](const float *g, const float2 w) {
        if constexpr (CT) {  // (gr + j gi)(wx + j wy)
            yr += g[0] * w.x - g[1] * w.y;
            yi += g[0] * w.y + g[1] * w.x;
        } else {
            yr = fmaf(g[0], w.x, yr);
            yi = fmaf(g[0], w.y, yi);
        }
    };
    if constexpr (INTERP) {
        const long long n = o / a.M;
        const float *g = img + (size_t)(o - n * a.M) * rs;
        const int S = a.T - 1;
        for (int j = 0; j < S; ++j) mac(g + (size_t)j * TW, fetch(n - j));
        a.out[(long long)ch * a.out_stride + o] = make_float2(yr, yi);  // no scale: pfb.rs:85-90
    } else {
        const long long n = o * a.M + (a.M - 1 - a.c0);
        int p = 0, q = 0;
        for (int k = 0; k < a.T; ++k) {
            mac(img + (size_t)p * rs + (size_t)q * TW, fetch(n - k));
            if (++p == a.M) { p = 0; ++q; }
        }
        float2 y;
        if constexpr (CT) y = make_float2(yr * a.scale_re - yi * a.scale_im, yr * a.scale_im + yi * a.scale_re);
        else y = make_float2(yr * a.scale_re, yi * a.scale_re);
        a.out[(long long)ch * a.out_stride + o] = y;
    }
}

#include "fir_walk.cuh"

}  // namespace
}  // namespace sgpu

// =============================================================================================
// Handles
// =============================================================================================
using namespace sgpu;

namespace {

constexpr int kR = 16;
constexpr int kNT = 128;  // 4 warps per block: one per SM sub-partition

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool packed_default() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SGPU_FIR_SCALAR_FMA");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// Host-side image of the taps as the kernels consume them: nsets phase filters of Qpad taps, tw
// floats per tap (1 real, 2 complex), rows Qpad*tw + kTapSkew floats apart.
void build_tap_image(const std::vector<float> &phase_taps /*[nsets][Q][tw]*/, int nsets, int Q, int Qpad, int tw,
                     std::vector<float> &img) {
    const size_t rs = (size_t)Qpad * tw + kTapSkew;
    img.assign((size_t)nsets * rs, 0.f);
    for (int p = 0; p < nsets; ++p)
        for (int q = 0; q < Q; ++q)
            for (int c = 0; c < tw; ++c) img[p * rs + (size_t)q * tw + c] = phase_taps[((size_t)p * Q + q) * tw + c];
}

}  // namespace

struct sgpu_fir {
    int device = 0, sm_count = 0;
    size_t T = 0, C = 0, M = 1;
    bool is_decim = false, complex_taps = false, packed = true;
    bool per_channel = false;     // every channel owns its taps (one reference object per channel, fir/mod.rs:79-88)
    double scale_re = 1.0, scale_im = 0.0;
    std::vector<float> taps_f32;  // caller order h[0..T), rounded to f32 (x2 when complex); [C][T] when per_channel
    uint64_t current_item = 0;    // fir/decim.rs:8
    size_t ch_off = 0;            // first channel of the block being launched (grids carry <= 65535 channels in y)
    int Q = 0, Qpad = 0;          // taps per phase
    size_t img_floats = 0;        // floats of one tap image
    float *d_taps = nullptr;      // tap image(s)
    float2 *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    bool hist_written = false;    // the launch of this call wrote the new history tail into d_hist[cur ^ 1]
    Staging stage;
    HostPipe pipe;
    FirTcState *tc = nullptr;     // tensor-core path for long filters (fir_tc.cu), built on first use
    bool tc_tried = false;
    int last_path = 0;            // 0 = FFMA2 kernels, 1 = tensor cores
    // NCO mix-down in front of the filter (set by an sgpu_ddc around its calls): [C][2] device words (theta0, delta),
    // the 1024 x (cos, sin) table and the stream position of the call's first sample
    const unsigned *nco_tab = nullptr;
    const float2 *nco_lut = nullptr;
    unsigned nco_pos = 0;
    float2 *d_mix = nullptr;      // pre-mixed input of the shapes without a fused kernel
    size_t mix_cap = 0;           // samples
    int last_mix_fused = 0;
};

// Is this call long enough for the tensor kernel?  A tensor tile (16384 outputs, one CTA) takes 16-30 us from the first
// sample load to the last store whatever the call size, and the band load, TMEM allocation and the second launch add to
// it, so calls that do not fill the chip are served faster by the FFMA2 kernel, whose short-call geometry is one 512-output
// tile per warp.  Measured per-call completion times, back-to-back calls with device pointers
// (tools/call_size_crossover.py; tensor / FFMA2 in us):
//   128 taps:  2^19 24.8 / 10.5   2^20 24.9 / 16.7   2^21 26.9 / 25.8   2^22 41.2 /  47.2   2^23  66 /   89
//   512 taps:  2^17 28.7 / 18.5   2^18 30.6 / 18.5   2^19 31.0 / 26.8   2^20 31.1 /  47.4   2^21  33 /   78   2^23 78 / 301
//  2048 taps:  2^17 57.2 / 51.3   2^18 57.4 / 51.3   2^19 57.6 / 88.3   2^20 59.7 / 168.5   2^21  62 /  291
// -> tensor from 2^19 samples AND 2^29 sample-taps per call on (complex taps count twice).  SGPU_FIR_TC_MIN_SAMPLES, when
// set, replaces the rule by a plain per-channel sample count (the GPU tests pin 2^15 to exercise the tensor kernel on
// short streams).
static bool tc_call_is_long_enough(const sgpu_fir *f, long long n_in) {
    if (const char *e = getenv("SGPU_FIR_TC_MIN_SAMPLES")) return n_in >= atoll(e);
    const long long total = n_in * (long long)f->C;
    return n_in >= (1LL << 15) && total >= (1LL << 19) && total * (long long)f->T * (f->complex_taps ? 2 : 1) >= (1LL << 29);
}


static int fir_R(const sgpu_fir *f) { return f->complex_taps ? 8 : kR; }

static int fir_upload_taps(sgpu_fir *f) {
    const int T = (int)f->T, M = (int)f->M, tw = f->complex_taps ? 2 : 1;
    f->Q = (T + M - 1) / M;
    f->Qpad = (int)round_up((size_t)f->Q, 2 * fir_R(f));
    // sub-filters of <= 16 taps: single-chunk instantiations of the warp-private kernels (no zero padding to 32)
    if (f->Q <= kR && !f->complex_taps && f->packed && (f->M == 1 || f->M == 2 || f->M == 4 || f->M == 8 || f->M == 16 || f->M == 32))
        f->Qpad = kR;
    const size_t nimg = f->per_channel ? f->C : 1;
    std::vector<float> all, img, ph((size_t)M * f->Q * tw);
    for (size_t c = 0; c < nimg; ++c) {
        // g[k] = h[T-1-k] (REVERSE, fir/mod.rs:86); phase p filter: g_p[q] = g[q*M + p]
        std::fill(ph.begin(), ph.end(), 0.f);
        const float *h = f->taps_f32.data() + c * (size_t)T * tw;
        for (int k = 0; k < T; ++k)
            for (int cc = 0; cc < tw; ++cc) ph[((size_t)(k % M) * f->Q + k / M) * tw + cc] = h[(size_t)(T - 1 - k) * tw + cc];
        build_tap_image(ph, M, f->Q, f->Qpad, tw, img);
        f->img_floats = img.size();
        all.insert(all.end(), img.begin(), img.end());
    }
    if (f->d_taps) cudaFree(f->d_taps);
    f->d_taps = nullptr;
    SGPU_CUDA(cudaMalloc(&f->d_taps, all.size() * sizeof(float)));
    SGPU_CUDA(cudaMemcpy(f->d_taps, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice));
    return SGPU_OK;
}

static int fir_create_impl(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels, bool per_channel,
                           double scale_re, double scale_im, int is_decimator, size_t decimation, sgpu_fir **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "fir_create: out is NULL");
    *out = nullptr;
    if (n_taps == 0 || !taps)  // fir/mod.rs:80-82, decim.rs:28-29
        return fail(SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO, "FIR Filter Error CoefficientsLengthZero");
    if (is_decimator && decimation < 1)  // decim.rs:30-31
        return fail(SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE, "FIR Filter Error DecimationLessThanOne");
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "fir_create: n_channels == 0");
    if (n_taps > (1u << 20) || (is_decimator && decimation > (1u << 20)))
        return fail(SGPU_ERR_UNSUPPORTED, "fir_create: n_taps/decimation beyond supported range");
    int dev = 0, sms = 0;
    int st = require_device(&dev, &sms);
    if (st) return st;
    sgpu_fir *f = new (std::nothrow) sgpu_fir();
    if (!f) return fail(SGPU_ERR_ALLOC, "out of host memory");
    f->device = dev;
    f->sm_count = sms;
    f->T = n_taps;
    f->C = n_channels;
    f->is_decim = is_decimator != 0;
    f->M = f->is_decim ? decimation : 1;
    f->scale_re = scale_re;
    f->scale_im = scale_im;
    f->complex_taps = kind == SGPU_TAPS_COMPLEX;
    f->per_channel = per_channel;
    f->packed = f->complex_taps ? true : packed_default();
    const size_t tw = f->complex_taps ? 2 : 1;
    const size_t nt = n_taps * tw * (per_channel ? n_channels : 1);
    f->taps_f32.resize(nt);
    for (size_t i = 0; i < nt; ++i) f->taps_f32[i] = (float)taps[i];
    st = fir_upload_taps(f);
    if (st) { sgpu_fir_destroy(f); return st; }
    const size_t hbytes = n_channels * (n_taps > 1 ? n_taps - 1 : 1) * sizeof(float2);
    for (int i = 0; i < 2; ++i) {
        if (cudaMalloc(&f->d_hist[i], hbytes) != cudaSuccess) {
            sgpu_fir_destroy(f);
            return fail(SGPU_ERR_CUDA, "cudaMalloc(history %zu bytes) failed", hbytes);
        }
        cudaMemset(f->d_hist[i], 0, hbytes);
    }
    *out = f;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                double scale_re, double scale_im, int is_decimator, size_t decimation,
                                sgpu_fir **out) {
    return fir_create_impl(taps, n_taps, kind, n_channels, false, scale_re, scale_im, is_decimator, decimation, out);
}

SGPU_EXPORT int sgpu_fir_create_per_channel(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                            double scale_re, double scale_im, int is_decimator, size_t decimation,
                                            sgpu_fir **out) {
    return fir_create_impl(taps, n_taps, kind, n_channels, true, scale_re, scale_im, is_decimator, decimation, out);
}

SGPU_EXPORT int sgpu_fir_destroy(sgpu_fir *f) {
    if (!f) return SGPU_OK;
    DeviceGuard g(f->device);
    if (f->d_taps) cudaFree(f->d_taps);
    for (int i = 0; i < 2; ++i)
        if (f->d_hist[i]) cudaFree(f->d_hist[i]);
    fir_tc_destroy(f->tc);
    if (f->d_mix) cudaFree(f->d_mix);
    f->stage.release();
    f->pipe.release();
    delete f;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_fir_out_len(const sgpu_fir *f, size_t n_in) {
    if (!f) return 0;
    return f->is_decim ? (size_t)((f->current_item + n_in) / f->M) : n_in;
}
SGPU_EXPORT size_t sgpu_fir_len(const sgpu_fir *f) { return f ? f->T : 0; }
SGPU_EXPORT size_t sgpu_fir_decimation(const sgpu_fir *f) { return f ? f->M : 0; }
SGPU_EXPORT size_t sgpu_fir_channels(const sgpu_fir *f) { return f ? f->C : 0; }
SGPU_EXPORT int sgpu_fir_last_path(const sgpu_fir *f) { return f ? f->last_path : 0; }
SGPU_EXPORT int sgpu_fir_taps_per_channel(const sgpu_fir *f) { return f && f->per_channel ? 1 : 0; }

SGPU_EXPORT int sgpu_fir_set_scale(sgpu_fir *f, double re, double im) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    f->scale_re = re;
    f->scale_im = im;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_fir_get_scale(const sgpu_fir *f, double *re, double *im) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (re) *re = f->scale_re;
    if (im) *im = f->scale_im;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_fir_channel_coefficients(const sgpu_fir *f, size_t channel, double *out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    if (channel >= f->C) return fail(SGPU_ERR_INVALID_ARGUMENT, "channel %zu >= %zu", channel, f->C);
    const size_t tw = f->complex_taps ? 2 : 1;
    const float *h = f->taps_f32.data() + (f->per_channel ? channel * f->T * tw : 0);
    for (size_t i = 0; i < f->T; ++i)  // stored (reversed) order
        for (size_t c = 0; c < tw; ++c) out[i * tw + c] = (double)h[(f->T - 1 - i) * tw + c];
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_fir_coefficients(const sgpu_fir *f, double *out) { return sgpu_fir_channel_coefficients(f, 0, out); }

namespace {

// cudaFuncSetAttribute is a driver round trip (~1 us): remember the largest size set per (device, kernel)
template <typename K>
int set_smem(K kernel, size_t bytes) {
    static std::mutex m;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_pair(dev, reinterpret_cast<const void *>(kernel));
    std::lock_guard<std::mutex> lock(m);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return SGPU_OK;
    SGPU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done[key] = bytes;
    return SGPU_OK;
}

// [ib, ib + span_in) and [ob, ob + span_out) share a byte
bool ranges_overlap(const void *in, size_t span_in, const void *out, size_t span_out) {
    const char *ib = reinterpret_cast<const char *>(in), *ob = reinterpret_cast<const char *>(out);
    return span_in && span_out && ib < ob + span_out && ob < ib + span_in;
}

// Enqueue the stand-alone history update for a handle (ping-pong) on `s`: write / push, and calls that produce no output.
int enqueue_hist_update(const float2 *d_in, long long in_stride, long long n_in, float2 *hist[2], int &cur,
                        size_t C, size_t H, cudaStream_t s) {
    if (H == 0 || n_in == 0) return SGPU_OK;
    for (size_t c0 = 0; c0 < C; c0 += kMaxGridY) {
        dim3 grid((unsigned)ceil_div(H, 128), (unsigned)std::min<size_t>(kMaxGridY, C - c0));
        hist_update_kernel<<<grid, 128, 0, s>>>(d_in + (long long)c0 * in_stride, in_stride, n_in, hist[cur] + c0 * H,
                                                hist[cur ^ 1] + c0 * H, (int)H);
        SGPU_LAUNCH_CHECK();
        count_launch();
    }
    cur ^= 1;
    return SGPU_OK;
}

int fir_launch_block(sgpu_fir *f, const float2 *d_in, long long n_in, long long in_stride, float2 *d_out,
                     long long out_stride, long long n_out, cudaStream_t s);

constexpr size_t kNcoLutBytes = 1024 * sizeof(float2);

// Launch geometry of the warp-private decimator (fir_walk.cuh) for this handle; false when that kernel does not serve it.
struct DwarpGeom {
    int PS, NW, RS;
    bool one;
    size_t stage_b, taps_b;
};
bool dwarp_geometry(const sgpu_fir *f, DwarpGeom &g) {
    if (!((f->M == 2 || f->M == 4 || f->M == 8 || f->M == 16 || f->M == 32) && f->packed && !f->complex_taps &&
          env_int("SGPU_DEC_WARP", 1)))
        return false;
    int PS = env_int("SGPU_DEC_PS", f->M >= 4 ? 4 : 2);
    if (PS != 1 && PS != 2 && PS != 4) PS = 4;
    if (PS > (int)f->M) PS = (int)f->M;
    g.one = f->Qpad == kR;  // sub-filters of <= 16 taps: single tap chunk
    if (g.one) PS = f->M >= 4 ? 4 : 2;
    g.PS = PS;
    g.NW = g.one ? 4 : (PS == 1 ? 1 : (PS == 2 ? 2 : 4));
    const int G = 32 / PS;
    const int rows = f->Qpad / kR + G;
    g.RS = rows | 1;
    g.stage_b = std::max<size_t>((size_t)f->M * ((size_t)(kR / 2) * g.RS + 1), 32 * (kR / 2 + 1)) * sizeof(float4);
    g.taps_b = (size_t)f->M * (f->Qpad + kTapSkew) * sizeof(float);
    return true;
}
// The decimator kernel that serves this handle has a variant with the NCO mix-down fused into its tile loader
bool fir_mix_fusable(const sgpu_fir *f) {
    DwarpGeom g;
    if (!(f->M == 2 || f->M == 4 || f->M == 8) || !dwarp_geometry(f, g)) return false;
    return (size_t)g.NW * g.stage_b + g.taps_b + kNcoLutBytes <= (size_t)kMaxSmem && env_int("SGPU_DDC_FUSED", 1);
}

// Channels ride in grid.y (<= 65535): larger handles are launched in channel blocks.
int fir_launch(sgpu_fir *f, const float2 *d_in, long long n_in, long long in_stride, float2 *d_out,
               long long out_stride, long long n_out, cudaStream_t s) {
    const size_t Ctot = f->C;
    int st = SGPU_OK;
    for (size_t c0 = 0; c0 < Ctot && st == SGPU_OK; c0 += kMaxGridY) {
        f->C = std::min<size_t>(kMaxGridY, Ctot - c0);
        f->ch_off = c0;
        st = fir_launch_block(f, d_in + (long long)c0 * in_stride, n_in, in_stride, d_out + (long long)c0 * out_stride,
                              out_stride, n_out, s);
    }
    f->C = Ctot;
    f->ch_off = 0;
    return st;
}

int fir_launch_block(sgpu_fir *f, const float2 *d_in, long long n_in, long long in_stride, float2 *d_out,
                     long long out_stride, long long n_out, cudaStream_t s) {
    FirArgs a{};
    a.in = d_in;
    a.out = d_out;
    a.hist = f->d_hist[f->cur] + f->ch_off * (f->T - 1);
    a.hist_new = f->d_hist[f->cur ^ 1] + f->ch_off * (f->T - 1);  // every kernel below writes the new tail itself
    a.taps = f->d_taps + (f->per_channel ? f->ch_off * f->img_floats : 0);
    a.tap_stride = f->per_channel ? (long long)f->img_floats : 0;
    a.in_stride = in_stride;
    a.out_stride = out_stride;
    a.n_in = n_in;
    a.n_out = n_out;
    a.T = (int)f->T;
    a.M = (int)f->M;
    a.c0 = (int)f->current_item;
    a.Qpad = f->Qpad;
    a.vec_in = ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) && (in_stride % 2 == 0);
    a.vec_out = ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0) && (out_stride % 2 == 0);
    a.scale_re = (float)f->scale_re;
    a.scale_im = (float)f->scale_im;
    a.nco = f->nco_tab ? f->nco_tab + 2 * f->ch_off : nullptr;
    a.lut = f->nco_tab ? f->nco_lut : nullptr;
    a.nco_pos = f->nco_pos;
    f->last_path = 0;
    if (n_out <= 0) return SGPU_OK;  // a decimator call shorter than one period: the caller enqueues the history update
    if (f->M == 1 && !f->per_channel && (f->complex_taps || f->scale_im == 0.0) &&
        (long long)f->T >= env_int("SGPU_FIR_TC_MIN_TAPS", f->complex_taps ? 56 : 112) && f->T <= 16384 /* band matrix: 512 B per tap */ &&
        tc_call_is_long_enough(f, n_in) && env_int("SGPU_FIR_TC", 1)) {
        // Long filters (real or complex taps): banded-Toeplitz product on the tcgen05 tensor cores (fir_tc.cu, DESIGN
        // 4.9).  Thresholds from measurements: tools/tc_taps_crossover.py (taps) and tools/call_size_crossover.py (call
        // size, see tc_call_is_long_enough).
        if (!f->tc_tried) {
            f->tc_tried = true;
            // a failed set-up is not fatal: the FP32 kernels below serve the call, the reason stays in sgpu_last_error
            if (fir_tc_create(&f->tc, f->taps_f32.data(), (int)f->T, f->complex_taps) != SGPU_OK) f->tc = nullptr;
        }
        if (f->tc) {
            int st = fir_tc_run(f->tc, d_in, n_in, in_stride, a.hist, (int)f->T - 1, f->T > 1 ? a.hist_new : nullptr, d_out,
                                out_stride, f->C, (float)f->scale_re, (float)f->scale_im, f->sm_count, s);
            if (st) return st;
            f->last_path = 1;
            f->hist_written = true;
            return SGPU_OK;
        }
    }
    if (f->M == 1 && f->packed && !f->complex_taps && env_int("SGPU_FIR_WARP", 1)) {
        // plain FIR, warp-private tiles (fir_walk.cuh)
        const int rows = f->Qpad / kR + 32;
        a.RS = rows | 1;
        const size_t stage_b = std::max<size_t>((size_t)(kR / 2) * a.RS + 1, 32 * (kR / 2 + 1)) * sizeof(float4);
        const size_t taps_b = (size_t)(f->Qpad + kTapSkew) * sizeof(float);
        const int ns = env_int("SGPU_FIR_NS", 1);
        int st = SGPU_OK;
        bool done = false;
#define LAUNCH_FWARP(NWV, MB, NSV, ONEV, TPWV)                                                \
    do {                                                                                      \
        const size_t smem = (size_t)NWV * NSV * stage_b + taps_b;                             \
        if (smem <= (size_t)kMaxSmem) {                                                       \
            auto kern = fir_warp_kernel<kR, NWV, MB, TPWV, NSV, ONEV>;                        \
            st = set_smem(kern, smem);                                                        \
            if (st) return st;                                                                \
            const long long per_block = 32LL * kR * TPWV * NWV;                               \
            dim3 grid((unsigned)((n_out + per_block - 1) / per_block), (unsigned)f->C);       \
            if (pdl) {                                                                        \
                cudaLaunchConfig_t cfg{};                                                     \
                cfg.gridDim = grid;                                                           \
                cfg.blockDim = dim3(NWV * 32);                                                \
                cfg.dynamicSmemBytes = smem;                                                  \
                cfg.stream = s;                                                               \
                cudaLaunchAttribute at[1];                                                    \
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                \
                at[0].val.programmaticStreamSerializationAllowed = 1;                         \
                cfg.attrs = at;                                                               \
                cfg.numAttrs = 1;                                                             \
                SGPU_CUDA(cudaLaunchKernelEx(&cfg, kern, a));                                 \
            } else {                                                                          \
                kern<<<grid, NWV * 32, smem, s>>>(a);                                         \
            }                                                                                 \
            done = true;                                                                      \
        }                                                                                     \
    } while (0)
        // Eight tiles per warp amortise the tap staging on long streams, but a short call (BASELINE config 1: 64 taps
        // x 2^20 samples = 64 blocks of 16384 outputs) would leave most of the chip idle: one tile per warp there.
        const bool small = (n_out + 16383) / 16384 * (long long)f->C < 4LL * f->sm_count && env_int("SGPU_FIR_SMALL", 1);
        // short calls are bounded by the launch-to-launch floor of a stream (6.4 us per dependent kernel):
        // programmatic dependent launch lets call k + 1 be scheduled and stage its taps while call k runs
        const bool pdl = small && env_int("SGPU_FIR_PDL", 1);
        if (f->Qpad == kR) {  // <= 16 taps: single tap chunk
            if (small) LAUNCH_FWARP(4, 4, 1, true, 1);
            else LAUNCH_FWARP(4, 4, 1, true, 8);
        } else if (ns == 2) LAUNCH_FWARP(4, 3, 2, false, 8);
        else if (small) LAUNCH_FWARP(4, 4, 1, false, 1);
        else LAUNCH_FWARP(4, 4, 1, false, 8);
        if (!done && f->Qpad != kR) LAUNCH_FWARP(1, 1, 1, false, 8);  // very long filters: one warp per block
#undef LAUNCH_FWARP
        if (done) {
            SGPU_LAUNCH_CHECK();
            count_launch();
            f->hist_written = true;
            return SGPU_OK;
        }
    }
    DwarpGeom dg;
    if (dwarp_geometry(f, dg)) {
        // warp-private tiles (fir_walk.cuh)
        const int PS = dg.PS;
        const bool one = dg.one;
        a.RS = dg.RS;
        const size_t stage_b = dg.stage_b, taps_b = dg.taps_b;
        constexpr int TPW = 8;
        int st = SGPU_OK;
        bool done = false;
#define LAUNCH_DWARP_X(MV, PSV, NWV, MB, ONEV, MIXV)                                          \
    do {                                                                                      \
        const size_t smem = (size_t)NWV * stage_b + taps_b + (MIXV ? kNcoLutBytes : 0);       \
        if (smem <= (size_t)kMaxSmem) {                                                       \
            auto kern = fir_decim_warp_kernel<kR, MV, PSV, NWV, MB, TPW, ONEV, MIXV>;         \
            st = set_smem(kern, smem);                                                        \
            if (st) return st;                                                                \
            const long long per_block = (long long)(32 / PSV) * kR * TPW * NWV;               \
            dim3 grid((unsigned)((n_out + per_block - 1) / per_block), (unsigned)f->C);       \
            kern<<<grid, NWV * 32, smem, s>>>(a);                                             \
            done = true;                                                                      \
        }                                                                                     \
    } while (0)
#define LAUNCH_DWARP(MV, PSV, NWV, MB, ONEV) LAUNCH_DWARP_X(MV, PSV, NWV, MB, ONEV, false)
    // M = 2, 4, 8 also exist with the NCO mix-down fused into the tile loader (DDC)
#define LAUNCH_DWARP_MIX(MV, PSV, NWV, MB, ONEV)                                              \
    do {                                                                                      \
        if (a.lut) LAUNCH_DWARP_X(MV, PSV, NWV, MB, ONEV, true);                              \
        else LAUNCH_DWARP_X(MV, PSV, NWV, MB, ONEV, false);                                   \
    } while (0)
#define LAUNCH_DWARP_M(MV)                                                                    \
    do {                                                                                      \
        if (one) LAUNCH_DWARP_MIX(MV, (MV >= 4 ? 4 : 2), 4, 4, true);                         \
        else if (PS == 1) LAUNCH_DWARP_MIX(MV, 1, 1, 4, false);                               \
        else if (PS == 2) LAUNCH_DWARP_MIX(MV, 2, 2, 4, false);                               \
        else LAUNCH_DWARP_MIX(MV, (MV >= 4 ? 4 : 2), 4, 4, false);                            \
    } while (0)
        if (f->M == 8) LAUNCH_DWARP_M(8);
        else if (f->M == 4) LAUNCH_DWARP_M(4);
        else if (f->M == 2) LAUNCH_DWARP_M(2);
        else if (a.lut) return fail(SGPU_ERR_UNSUPPORTED, "NCO mix requested on a kernel without a fused variant");
        else if (f->M == 16) {  // 16 / 32 phase planes per stage: fewer warps per block as the tile grows
            if (one) LAUNCH_DWARP(16, 4, 4, 2, true);
            else {
                LAUNCH_DWARP(16, 4, 4, 2, false);
                if (!done) LAUNCH_DWARP(16, 4, 2, 2, false);
                if (!done) LAUNCH_DWARP(16, 4, 1, 2, false);
            }
        } else {
            if (one) LAUNCH_DWARP(32, 4, 2, 2, true);
            else {
                LAUNCH_DWARP(32, 4, 2, 2, false);
                if (!done) LAUNCH_DWARP(32, 4, 1, 2, false);
            }
        }
#undef LAUNCH_DWARP_M
#undef LAUNCH_DWARP_MIX
#undef LAUNCH_DWARP
#undef LAUNCH_DWARP_X
        if (done) {
            SGPU_LAUNCH_CHECK();
            count_launch();
            f->hist_written = true;
            return SGPU_OK;
        }
    }
    if (a.lut) return fail(SGPU_ERR_UNSUPPORTED, "NCO mix requested on a kernel without a fused variant");
    {
        const bool m1 = f->M == 1;
        const int R = fir_R(f);
        const int twc = f->complex_taps ? 2 : 1;
        // phase split: as many lanes per output run as there are phases to share, up to 4
        int PS = m1 ? 1 : (f->M >= 4 ? 4 : (f->M >= 2 ? 2 : 1));  // measured at M=8: PS=4 338, PS=2 311, PS=1 192 G in-samp/s
        const int want = env_int("SGPU_DEC_PS", 0);
        if (!m1 && (want == 1 || want == 2 || want == 4) && want <= (int)f->M) PS = want;
        const int OT = kNT / PS;
        const int rows = f->Qpad / R + OT;
        a.RS = rows | 1;
        const size_t plane_f4 = (size_t)(R / 2) * a.RS + 1;
        const size_t smem = f->M * plane_f4 * sizeof(float4) + (size_t)f->M * (f->Qpad * twc + kTapSkew) * sizeof(float);
        int st;
        if (smem > (size_t)kMaxSmem || f->Qpad % (2 * R) != 0) {
            // the tile (all M phase planes + the taps) does not fit one SM's shared memory, e.g. 256 taps at M = 64:
            // direct form, one thread per output.  Any factor and any length the reference takes is served.
            dim3 grid((unsigned)ceil_div((size_t)n_out, 128), (unsigned)f->C);
            if (f->complex_taps) fir_direct_kernel<false, true><<<grid, 128, 0, s>>>(a);
            else fir_direct_kernel<false, false><<<grid, 128, 0, s>>>(a);
            SGPU_LAUNCH_CHECK();
            count_launch();
            f->hist_written = true;
            return SGPU_OK;
        }
        const long long tiles = (n_out + (long long)OT * R - 1) / ((long long)OT * R);
        dim3 grid((unsigned)tiles, (unsigned)f->C);
#define LAUNCH_FIR(RV, PK, M1, MINB, PSV, CTV)                                              \
    do {                                                                                    \
        auto kern = fir_decim_kernel<RV, PK, M1, kNT, MINB, PSV, CTV>;                      \
        st = set_smem(kern, smem);                                                          \
        if (st) return st;                                                                  \
        kern<<<grid, kNT, smem, s>>>(a);                                                    \
    } while (0)
        if (f->complex_taps) {
            if (m1) LAUNCH_FIR(8, true, true, 4, 1, true);
            else if (PS == 4) LAUNCH_FIR(8, true, false, 4, 4, true);
            else if (PS == 2) LAUNCH_FIR(8, true, false, 3, 2, true);
            else LAUNCH_FIR(8, true, false, 1, 1, true);
        } else if (m1) {
            if (f->packed) LAUNCH_FIR(kR, true, true, 4, 1, false);
            else LAUNCH_FIR(kR, false, true, 4, 1, false);
        } else if (PS == 4) {
            if (f->packed) LAUNCH_FIR(kR, true, false, 4, 4, false);
            else LAUNCH_FIR(kR, false, false, 4, 4, false);
        } else if (PS == 2) {
            if (f->packed) LAUNCH_FIR(kR, true, false, 3, 2, false);
            else LAUNCH_FIR(kR, false, false, 3, 2, false);
        } else {
            if (f->packed) LAUNCH_FIR(kR, true, false, 1, 1, false);
            else LAUNCH_FIR(kR, false, false, 1, 1, false);
        }
#undef LAUNCH_FIR
        SGPU_LAUNCH_CHECK();
        count_launch();
        f->hist_written = true;
    }
    return SGPU_OK;
}

}  // namespace

// One pass over device-resident samples: ONE kernel computes the outputs and writes the new history tail (the tensor
// path: two launches, fir_tc.cuh); only the phase counter lives on the host.  With an NCO attached (sgpu_ddc) the
// mix-down runs inside the decimator's tile loader where a fused variant exists, else as a kernel of its own in front.
static int fir_run_device(sgpu_fir *f, const float2 *d_in, size_t nc, long long istr, float2 *d_out, long long ostr,
                          size_t nout, cudaStream_t st_) {
    const unsigned *tab = f->nco_tab;
    if (tab) {
        const bool fused = nout > 0 && fir_mix_fusable(f);
        f->last_mix_fused = fused ? 1 : 0;
        if (!fused) {
            const size_t need = f->C * nc;
            if (need > f->mix_cap) {
                if (f->d_mix) {
                    SGPU_CUDA(cudaStreamSynchronize(st_));
                    cudaFree(f->d_mix);
                }
                f->d_mix = nullptr;
                f->mix_cap = 0;
                if (cudaMalloc(&f->d_mix, need * sizeof(float2)) != cudaSuccess)
                    return fail(SGPU_ERR_ALLOC, "cudaMalloc(mixed samples, %zu bytes) failed", need * sizeof(float2));
                f->mix_cap = need;
            }
            int st = nco_mix_launch(false, d_in, istr, f->d_mix, (long long)nc, (long long)nc, f->C, tab, f->nco_pos, f->nco_lut, st_);
            if (st) return st;
            d_in = f->d_mix;
            istr = (long long)nc;
            f->nco_tab = nullptr;  // the kernels below see mixed samples
        }
    }
    f->hist_written = false;
    int st = fir_launch(f, d_in, (long long)nc, istr, d_out, ostr, (long long)nout, st_);
    if (st == SGPU_OK) {
        if (f->hist_written) f->cur ^= 1;
        else if (f->nco_tab) st = fail(SGPU_ERR_UNSUPPORTED, "fused mix without a history update");
        else st = enqueue_hist_update(d_in, istr, (long long)nc, f->d_hist, f->cur, f->C, f->T - 1, st_);
    }
    f->nco_tab = tab;
    if (st) return st;
    if (tab) f->nco_pos += (unsigned)nc;                                // one NCO::step per sample (nco/mod.rs:93-96)
    if (f->is_decim) f->current_item = (f->current_item + nc) % f->M;  // decim.rs:116
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_execute_block(sgpu_fir *f, const float *in, size_t n_in, size_t in_stride,
                                       float *out, size_t out_stride, size_t *n_out_p, sgpu_mem mem,
                                       void *stream) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    const size_t n_out = sgpu_fir_out_len(f, n_in);
    if (n_out_p) *n_out_p = n_out;
    if (n_in == 0) return SGPU_OK;
    if (!in || (n_out && !out)) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (f->C > 1 && in_stride < n_in) return fail(SGPU_ERR_INVALID_ARGUMENT, "in_stride < n_in");
    if (out_stride < n_out) return fail(SGPU_ERR_CAPACITY, "out capacity %zu < %zu outputs", out_stride, n_out);
    // the kernels read a tile's halo while other blocks already write their outputs: in and out must be disjoint
    if (n_out && ranges_overlap(in, ((f->C - 1) * in_stride + n_in) * 8, out, ((f->C - 1) * out_stride + n_out) * 8))
        return fail(SGPU_ERR_INVALID_ARGUMENT, "in and out overlap: execute_block is not an in-place operation");
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    auto run = [f](const float2 *d_in, size_t nc, long long istr, float2 *d_out, long long ostr, size_t nout,
                   cudaStream_t st_) -> int { return fir_run_device(f, d_in, nc, istr, d_out, ostr, nout, st_); };
    if (mem == SGPU_DEVICE)
        return run(reinterpret_cast<const float2 *>(in), n_in, (long long)in_stride, reinterpret_cast<float2 *>(out),
                   (long long)out_stride, n_out, s);
    return host_pipeline(f->pipe, f->C, in, n_in, in_stride, out, out_stride, 1,
                         [f](size_t nc) { return sgpu_fir_out_len(f, nc); }, run, s);
}

SGPU_EXPORT int sgpu_fir_write(sgpu_fir *f, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
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
    int st = enqueue_hist_update(d_in, istr, (long long)n_in, f->d_hist, f->cur, f->C, f->T - 1, s);
    if (st) return st;
    if (f->is_decim) f->current_item = (f->current_item + n_in) % f->M;  // decim.rs:137
    if (mem == SGPU_HOST) SGPU_CUDA(cudaStreamSynchronize(s));
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_get_state(sgpu_fir *f, float *history, uint64_t *current_item) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    if (history && f->T > 1) {
        SGPU_CUDA(cudaDeviceSynchronize());
        SGPU_CUDA(cudaMemcpy(history, f->d_hist[f->cur], f->C * (f->T - 1) * sizeof(float2),
                             cudaMemcpyDeviceToHost));
    }
    if (current_item) *current_item = f->current_item;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_set_state(sgpu_fir *f, const float *history, uint64_t current_item) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    if (history && f->T > 1) {
        SGPU_CUDA(cudaDeviceSynchronize());
        SGPU_CUDA(cudaMemcpy(f->d_hist[f->cur], history, f->C * (f->T - 1) * sizeof(float2),
                             cudaMemcpyHostToDevice));
    }
    f->current_item = f->is_decim ? current_item % f->M : 0;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_reset(sgpu_fir *f) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_hist[f->cur], 0, f->C * (f->T > 1 ? f->T - 1 : 1) * sizeof(float2)));
    f->current_item = 0;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_fir_clone(const sgpu_fir *f, sgpu_fir **out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    std::vector<double> taps(f->taps_f32.size());
    for (size_t i = 0; i < taps.size(); ++i) taps[i] = (double)f->taps_f32[i];
    sgpu_fir *c = nullptr;
    int st = fir_create_impl(taps.data(), f->T, f->complex_taps ? SGPU_TAPS_COMPLEX : SGPU_TAPS_REAL, f->C, f->per_channel,
                             f->scale_re, f->scale_im, f->is_decim, f->M, &c);
    if (st) return st;
    if (f->T > 1) {
        SGPU_CUDA(cudaDeviceSynchronize());
        SGPU_CUDA(cudaMemcpy(c->d_hist[c->cur], f->d_hist[f->cur], f->C * (f->T - 1) * sizeof(float2),
                             cudaMemcpyDeviceToDevice));
    }
    c->current_item = f->current_item;
    *out = c;
    return SGPU_OK;
}

// =============================================================================================
// Digital down-converter: NCO::mix_down + step per sample (nco/mod.rs:93-96,147-151) feeding a DecimatingFIRFilter
// (fir/decim.rs:221-256) -- SURVEY 8f rank 3.  One NCO and one decimator per channel, the mixed stream never reaches HBM
// where the decimator kernel has a fused variant (fir_walk.cuh, MIX).
// =============================================================================================
struct sgpu_ddc {
    sgpu_fir *fir = nullptr;
    sgpu_nco *nco = nullptr;
};

SGPU_EXPORT int sgpu_ddc_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels, double scale_re,
                                double scale_im, size_t decimation, sgpu_ddc **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "ddc_create: out is NULL");
    *out = nullptr;
    sgpu_ddc *d = new (std::nothrow) sgpu_ddc();
    if (!d) return fail(SGPU_ERR_ALLOC, "out of host memory");
    int st = sgpu_fir_create(taps, n_taps, kind, n_channels, scale_re, scale_im, 1, decimation, &d->fir);
    if (st == SGPU_OK) st = sgpu_nco_create(n_channels, &d->nco);
    if (st) {
        sgpu_ddc_destroy(d);
        return st;
    }
    *out = d;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_ddc_destroy(sgpu_ddc *d) {
    if (!d) return SGPU_OK;
    sgpu_fir_destroy(d->fir);
    sgpu_nco_destroy(d->nco);
    delete d;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_ddc_clone(const sgpu_ddc *d, sgpu_ddc **out) {
    if (!d || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    sgpu_ddc *c = new (std::nothrow) sgpu_ddc();
    if (!c) return fail(SGPU_ERR_ALLOC, "out of host memory");
    int st = sgpu_fir_clone(d->fir, &c->fir);
    if (st == SGPU_OK) st = sgpu_nco_clone(d->nco, &c->nco);
    if (st) {
        sgpu_ddc_destroy(c);
        return st;
    }
    *out = c;
    return SGPU_OK;
}

SGPU_EXPORT sgpu_fir *sgpu_ddc_filter(sgpu_ddc *d) { return d ? d->fir : nullptr; }
SGPU_EXPORT sgpu_nco *sgpu_ddc_nco(sgpu_ddc *d) { return d ? d->nco : nullptr; }
SGPU_EXPORT size_t sgpu_ddc_out_len(const sgpu_ddc *d, size_t n_in) { return d ? sgpu_fir_out_len(d->fir, n_in) : 0; }
SGPU_EXPORT int sgpu_ddc_last_fused(const sgpu_ddc *d) { return d && d->fir ? d->fir->last_mix_fused : 0; }

SGPU_EXPORT int sgpu_ddc_reset(sgpu_ddc *d) {  // NCO::reset (nco/mod.rs:53-56) + the decimator's window and counter
    if (!d) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    int st = sgpu_nco_reset(d->nco);
    return st ? st : sgpu_fir_reset(d->fir);
}

namespace {
// the decimator sees the NCO for the duration of one call
struct NcoAttach {
    sgpu_fir *f;
    sgpu_nco *n;
    NcoAttach(sgpu_fir *f_, sgpu_nco *n_) : f(f_), n(n_) {
        f->nco_tab = n->d_tab;
        f->nco_lut = n->d_lut;
        f->nco_pos = n->pos;
    }
    ~NcoAttach() {
        n->pos = f->nco_pos;
        f->nco_tab = nullptr;
    }
};
}  // namespace

SGPU_EXPORT int sgpu_ddc_execute_block(sgpu_ddc *d, const float *in, size_t n_in, size_t in_stride, float *out,
                                       size_t out_stride, size_t *n_out_p, sgpu_mem mem, void *stream) {
    if (!d) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (n_in) {
        DeviceGuard g(d->fir->device);
        int st = nco_sync_table(d->nco, (cudaStream_t)stream);
        if (st) return st;
    }
    NcoAttach attach(d->fir, d->nco);
    return sgpu_fir_execute_block(d->fir, in, n_in, in_stride, out, out_stride, n_out_p, mem, stream);
}

// DecimatingFIRFilter::write behind the mixer: the samples are mixed and pushed, no output (fir/decim.rs:136-139)
SGPU_EXPORT int sgpu_ddc_write(sgpu_ddc *d, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem, void *stream) {
    if (!d) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (n_in == 0) return SGPU_OK;
    if (!in) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    sgpu_fir *f = d->fir;
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    int st = nco_sync_table(d->nco, s);
    if (st) return st;
    const float2 *d_in = reinterpret_cast<const float2 *>(in);
    long long istr = (long long)in_stride;
    if (mem == SGPU_HOST) {
        st = f->stage.ensure(f->C * n_in * sizeof(float2), 0);
        if (st) return st;
        SGPU_CUDA(cudaMemcpy2DAsync(f->stage.in, n_in * sizeof(float2), in, in_stride * sizeof(float2), n_in * sizeof(float2), f->C,
                                    cudaMemcpyHostToDevice, s));
        d_in = (const float2 *)f->stage.in;
        istr = (long long)n_in;
    }
    {
        NcoAttach attach(f, d->nco);
        st = fir_run_device(f, d_in, n_in, istr, nullptr, 1, 0, s);  // no outputs: mix into scratch + history update
    }
    if (st) return st;
    if (mem == SGPU_HOST) SGPU_CUDA(cudaStreamSynchronize(s));
    return SGPU_OK;
}

// =============================================================================================
// InterpolatingFIRFilter / PolyPhaseFilterBank
// =============================================================================================
struct sgpu_interp {
    int device = 0, sm_count = 0;
    size_t T = 0, C = 0, L = 1, S = 0;  // S = sub-filter length
    bool packed = true, complex_taps = false;
    double scale_re = 1.0, scale_im = 0.0;  // stored, never applied (pfb.rs:85-90)
    bool per_channel = false;                // every channel owns its taps
    std::vector<float> phase_taps;           // [L][S][tw]: hp[p][j] = hpad[p + (S-1-j)*L] (newest first); [C][L][S][tw] when per_channel
    size_t ch_off = 0;                       // first channel of the block being launched
    int Qpad = 0;
    size_t img_floats = 0;                   // floats of one tap image
    float *d_taps = nullptr;
    float2 *d_hist[2] = {nullptr, nullptr};  // S samples per channel: the PFB window (oldest first)
    int cur = 0;
    bool hist_written = false;               // the launch of this call wrote the new window into d_hist[cur ^ 1]
    Staging stage;
    HostPipe pipe;
    FirTcState *tc = nullptr;  // tensor-core path (fir_tc.cu): real taps, L in {2, 4}, sub-filters of more than 32 taps
    bool tc_tried = false;
    int last_path = 0;
};

static int interp_build(sgpu_interp *f, const double *taps_eff /* [C when per_channel][eff_stride] values, L*S (complex: x2) used */,
                        size_t eff_stride) {
    const int L = (int)f->L, S = (int)f->S, tw = f->complex_taps ? 2 : 1;
    const size_t nimg = f->per_channel ? f->C : 1, per = (size_t)L * S * tw;
    f->phase_taps.assign(nimg * per, 0.f);
    f->Qpad = (int)round_up((size_t)S, 2 * (f->complex_taps ? 8 : kR));
    if (S <= kR && !f->complex_taps && f->packed && (L == 2 || L == 4 || L == 8 || L == 16 || L == 32))
        f->Qpad = kR;  // single-chunk walking kernel
    std::vector<float> all, img, one(per);
    for (size_t ch = 0; ch < nimg; ++ch) {
        const double *te = taps_eff + ch * eff_stride;
        for (int p = 0; p < L; ++p)
            for (int j = 0; j < S; ++j)
                for (int c = 0; c < tw; ++c)
                    one[((size_t)p * S + j) * tw + c] = (float)te[(p + (size_t)(S - 1 - j) * L) * tw + c];
        std::copy(one.begin(), one.end(), f->phase_taps.begin() + ch * per);
        build_tap_image(one, L, S, f->Qpad, tw, img);
        f->img_floats = img.size();
        all.insert(all.end(), img.begin(), img.end());
    }
    SGPU_CUDA(cudaMalloc(&f->d_taps, all.size() * sizeof(float)));
    SGPU_CUDA(cudaMemcpy(f->d_taps, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice));
    // window of S samples; kernels treat the last S-1 as "history" (T-1 with T = S)
    const size_t hbytes = f->C * (size_t)S * sizeof(float2);
    for (int i = 0; i < 2; ++i) {
        SGPU_CUDA(cudaMalloc(&f->d_hist[i], hbytes));
        SGPU_CUDA(cudaMemset(f->d_hist[i], 0, hbytes));
    }
    return SGPU_OK;
}

static int interp_create_common(const double *taps_eff, size_t eff_stride, size_t n_taps, size_t n_channels,
                                size_t L, size_t S, double sre, double sim, bool complex_taps, bool per_channel,
                                sgpu_interp **out) {
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    if (L > (1u << 20) || S > (1u << 20)) return fail(SGPU_ERR_UNSUPPORTED, "interpolation/sub-filter too large");
    int dev = 0, sms = 0;
    int st = require_device(&dev, &sms);
    if (st) return st;
    sgpu_interp *f = new (std::nothrow) sgpu_interp();
    if (!f) return fail(SGPU_ERR_ALLOC, "out of host memory");
    f->device = dev;
    f->sm_count = sms;
    f->T = n_taps;
    f->C = n_channels;
    f->L = L;
    f->S = S;
    f->scale_re = sre;
    f->scale_im = sim;
    f->complex_taps = complex_taps;
    f->per_channel = per_channel;
    f->packed = complex_taps ? true : packed_default();
    st = interp_build(f, taps_eff, eff_stride);
    if (st) { sgpu_interp_destroy(f); return st; }
    *out = f;
    return SGPU_OK;
}

static int interp_create_impl(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels, bool per_channel,
                              size_t interpolation, sgpu_interp **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (n_taps == 0 || !taps)  // interp.rs:28-29
        return fail(SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO, "FIR Filter Error CoefficientsLengthZero");
    if (interpolation < 1)  // interp.rs:30-31
        return fail(SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE, "FIR Filter Error InterpolationLessThanOne");
    const size_t tw = kind == SGPU_TAPS_COMPLEX ? 2 : 1;
    // interp.rs:35-40: sub-filter length through an f32 quotient
    const float q = (float)n_taps / (float)interpolation;
    const size_t S = (q == floorf(q)) ? (size_t)q : (size_t)ceilf(q);
    const size_t eff = S * interpolation;  // interp.rs:43
    const size_t stride = (eff > n_taps ? eff : n_taps) * tw, nimg = per_channel ? n_channels : 1;
    std::vector<double> padded(stride * nimg, 0.0);
    for (size_t ch = 0; ch < nimg; ++ch)
        for (size_t i = 0; i < n_taps * tw; ++i) padded[ch * stride + i] = taps[ch * n_taps * tw + i];
    return interp_create_common(padded.data(), stride, n_taps, n_channels, interpolation, S, 1.0, 0.0, tw == 2, per_channel, out);
}

SGPU_EXPORT int sgpu_interp_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                   size_t interpolation, sgpu_interp **out) {
    return interp_create_impl(taps, n_taps, kind, n_channels, false, interpolation, out);
}
SGPU_EXPORT int sgpu_interp_create_per_channel(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                               size_t interpolation, sgpu_interp **out) {
    return interp_create_impl(taps, n_taps, kind, n_channels, true, interpolation, out);
}

SGPU_EXPORT int sgpu_pfb_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                                size_t filters, double scale_re, double scale_im, sgpu_interp **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (filters == 0)  // pfb.rs:25-26
        return fail(SGPU_ERR_FIR_NOT_ENOUGH_FILTERS, "FIR Filter Error NotEnoughFilters");
    if (n_taps == 0 || !taps)  // pfb.rs:27-28
        return fail(SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO, "FIR Filter Error CoefficientsLengthZero");
    const size_t S = n_taps / filters;  // pfb.rs:32 (truncating)
    if (S == 0)  // reference: Window::new(0) assertion panic (window/mod.rs:18)
        return fail(SGPU_ERR_FIR_NOT_ENOUGH_FILTERS, "FIR Filter Error NotEnoughFilters (filters > taps)");
    return interp_create_common(taps, 0, n_taps, n_channels, filters, S, scale_re, scale_im,
                                kind == SGPU_TAPS_COMPLEX, false, out);
}

SGPU_EXPORT int sgpu_interp_destroy(sgpu_interp *f) {
    if (!f) return SGPU_OK;
    DeviceGuard g(f->device);
    if (f->d_taps) cudaFree(f->d_taps);
    for (int i = 0; i < 2; ++i)
        if (f->d_hist[i]) cudaFree(f->d_hist[i]);
    fir_tc_destroy(f->tc);
    f->stage.release();
    f->pipe.release();
    delete f;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_interp_last_path(const sgpu_interp *f) { return f ? f->last_path : 0; }
SGPU_EXPORT size_t sgpu_interp_interpolation(const sgpu_interp *f) { return f ? f->L : 0; }
SGPU_EXPORT size_t sgpu_interp_sub_len(const sgpu_interp *f) { return f ? f->S : 0; }
SGPU_EXPORT size_t sgpu_interp_channels(const sgpu_interp *f) { return f ? f->C : 0; }
SGPU_EXPORT int sgpu_interp_set_scale(sgpu_interp *f, double re, double im) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    f->scale_re = re;
    f->scale_im = im;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_interp_get_scale(const sgpu_interp *f, double *re, double *im) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    if (re) *re = f->scale_re;
    if (im) *im = f->scale_im;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_interp_coefficients(const sgpu_interp *f, double *out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    // pfb.rs:71-73: each DotProduct's stored order = rev_sub_coefs; rev_sub[S-1-idx] = h[p + idx*L]
    // so stored[i] = h[p + (S-1-i)*L] = phase_taps[p][i]
    const size_t per = f->L * f->S * (f->complex_taps ? 2 : 1);  // channel 0 when every channel owns its taps
    for (size_t i = 0; i < per; ++i) out[i] = (double)f->phase_taps[i];
    return SGPU_OK;
}

namespace {
int interp_launch_block(sgpu_interp *f, const float2 *d_in, long long n_in, long long istr, float2 *d_out, long long ostr,
                        long long n_out, cudaStream_t s);

int interp_launch(sgpu_interp *f, const float2 *d_in, long long n_in, long long istr, float2 *d_out, long long ostr,
                  long long n_out, cudaStream_t s) {
    const size_t Ctot = f->C;
    int st = SGPU_OK;
    for (size_t c0 = 0; c0 < Ctot && st == SGPU_OK; c0 += kMaxGridY) {  // channels ride in grid.y (<= 65535)
        f->C = std::min<size_t>(kMaxGridY, Ctot - c0);
        f->ch_off = c0;
        st = interp_launch_block(f, d_in + (long long)c0 * istr, n_in, istr, d_out + (long long)c0 * ostr, ostr, n_out, s);
    }
    f->C = Ctot;
    f->ch_off = 0;
    return st;
}

int interp_launch_block(sgpu_interp *f, const float2 *d_in, long long n_in, long long istr, float2 *d_out, long long ostr,
                        long long n_out, cudaStream_t s) {
    FirArgs a{};
    a.in = d_in;
    a.out = d_out;
    // the kernel's history convention is "T-1 samples before x[0]" with T = S+1 here: the PFB
    // window keeps S samples, of which the interpolator only ever reads the newest S-1 as past.
    a.hist = f->d_hist[f->cur] + f->ch_off * f->S;
    a.hist_new = f->d_hist[f->cur ^ 1] + f->ch_off * f->S;  // every kernel below writes the new window itself
    a.taps = f->d_taps + (f->per_channel ? f->ch_off * f->img_floats : 0);
    a.tap_stride = f->per_channel ? (long long)f->img_floats : 0;
    a.in_stride = istr;
    a.out_stride = ostr;
    a.n_in = n_in;
    a.n_out = n_out;
    a.T = (int)f->S + 1;
    a.M = (int)f->L;
    a.c0 = 0;
    a.Qpad = f->Qpad;
    a.vec_in = ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) && (istr % 2 == 0);
    a.vec_out = 0;
    a.scale_re = 1.f;
    const int tw = f->complex_taps ? 2 : 1;
    f->last_path = 0;
    if (n_out > 0 && !f->complex_taps && !f->per_channel && (f->L == 2 || f->L == 4) &&
        (long long)f->S >= env_int("SGPU_INTERP_TC_MIN_SUB", f->L == 2 ? 17 : 33) && f->S <= 16384 &&
        n_in >= 128 * (128 / (long long)f->L) &&
        n_out * (long long)f->C >= (long long)env_int("SGPU_INTERP_TC_MIN_OUT", 1 << 23) && env_int("SGPU_FIR_TC", 1)) {
        // polyphase interpolator as a banded product on the tcgen05 tensor cores (fir_tc.cu): 128 outputs per block
        // row = 128 / L inputs.  Measured (tools/tc_interp_probe.py, 256 ch x 2^20, G out-samp/s): L = 4: sub-filters <= 32
        // taps stay on the walking kernel (430 vs 422), 48 taps 373 vs 190, 256 taps 163 vs 58; L = 2: 24-32 taps 288 vs 263
        if (!f->tc_tried) {
            f->tc_tried = true;
            if (fir_tc_create_pfb(&f->tc, f->phase_taps.data(), (int)f->L, (int)f->S, false) != SGPU_OK) f->tc = nullptr;
        }
        if (f->tc) {
            int st = fir_tc_run(f->tc, d_in, n_in, istr, a.hist, (int)f->S, a.hist_new, d_out, ostr, f->C, 1.f, 0.f,
                                f->sm_count, s);
            if (st) return st;
            f->last_path = 1;
            f->hist_written = true;
            return SGPU_OK;
        }
    }
    if (!f->complex_taps && f->packed && (f->Qpad == 2 * kR || f->Qpad == kR) &&
        (f->L == 2 || f->L == 4 || f->L == 8 || f->L == 16 || f->L == 32) && env_int("SGPU_WALK", 1)) {
        // walking kernel (fir_walk.cuh): sub-filters of <= 16 / <= 32 taps, one lane per phase, warp-private tiles
        const int NC = f->Qpad / kR;
        const int K = (f->L < 16 && NC == 2 && env_int("SGPU_WALK_K", 5) == 7) ? 7 : 5;  // measured (L=4, 1024 ch): K=1 408, 3 412, 5 434, 7 421, 9 404 G out-samp/s
        const int G = 32 / (int)f->L;
        const int rows = NC + G * K;
        a.RS = rows | 1;
        const size_t smem = (size_t)(kNT / 32) * 2 * ((size_t)(kR / 2) * a.RS + 1) * sizeof(float4) +
                            f->L * (size_t)(f->Qpad + kTapSkew) * sizeof(float);
        const long long tile = (long long)G * K * kR;  // per warp
        constexpr int TPW = 8;
        const long long per_block = tile * TPW * (kNT / 32);
        dim3 grid((unsigned)((n_in + per_block - 1) / per_block), (unsigned)f->C);
        int st;
#define LAUNCH_IWALK(LV, KV, MB, NCV)                                     \
    do {                                                                  \
        auto kern = fir_interp_walk_kernel<kR, LV, KV, kNT, MB, TPW, NCV>; \
        st = set_smem(kern, smem);                                        \
        if (st) return st;                                                \
        kern<<<grid, kNT, smem, s>>>(a);                                  \
    } while (0)
#define LAUNCH_IWALK_T(LV)                                                \
    do {                                                                  \
        if (NC == 1) LAUNCH_IWALK(LV, 5, 3, 1);                           \
        else if (K == 7) LAUNCH_IWALK(LV, 7, 3, 2);                       \
        else LAUNCH_IWALK(LV, 5, 3, 2);                                   \
    } while (0)
        if (f->L == 2) LAUNCH_IWALK_T(2);
        else if (f->L == 4) LAUNCH_IWALK_T(4);
        else if (f->L == 8) LAUNCH_IWALK_T(8);
        else if (f->L == 16) { if (NC == 1) LAUNCH_IWALK(16, 5, 3, 1); else LAUNCH_IWALK(16, 5, 3, 2); }
        else { if (NC == 1) LAUNCH_IWALK(32, 5, 3, 1); else LAUNCH_IWALK(32, 5, 3, 2); }
#undef LAUNCH_IWALK_T
#undef LAUNCH_IWALK
        SGPU_LAUNCH_CHECK();
        count_launch();
        f->hist_written = true;
        return SGPU_OK;
    }
    int PS = f->L >= 2 ? 2 : 1;
    const int want = env_int("SGPU_INT_PS", 0);
    if ((want == 1 || want == 2 || want == 4) && want <= (int)f->L) PS = want;
    const int R = f->complex_taps ? 8 : kR;
    const int OT = kNT / PS;
    const int rows = f->Qpad / R + OT;
    a.RS = rows | 1;
    const size_t plane_f4 = (size_t)(R / 2) * a.RS + 1;
    const size_t smem = plane_f4 * sizeof(float4) + f->L * (size_t)(f->Qpad * tw + kTapSkew) * sizeof(float) +
                        (size_t)OT * (R * f->L + 1) * sizeof(float2);
    if (smem > (size_t)kMaxSmem || f->Qpad % (2 * R) != 0) {
        // L tap sets + the staged outputs of a tile do not fit one SM's shared memory (e.g. complex taps at L >= 28, any
        // L beyond ~100): direct form, one thread per output.  Every factor the reference takes is served.
        dim3 grid((unsigned)ceil_div((size_t)n_out, 128), (unsigned)f->C);
        if (f->complex_taps) fir_direct_kernel<true, true><<<grid, 128, 0, s>>>(a);
        else fir_direct_kernel<true, false><<<grid, 128, 0, s>>>(a);
        SGPU_LAUNCH_CHECK();
        count_launch();
        f->hist_written = true;
        return SGPU_OK;
    }
    const long long tiles = ((long long)n_in + (long long)OT * R - 1) / ((long long)OT * R);
    dim3 grid((unsigned)tiles, (unsigned)f->C);
    int st;
#define LAUNCH_INT(RV, PK, PSV, CTV)                                    \
    do {                                                                \
        auto kern = fir_interp_kernel<RV, PK, kNT, 4, PSV, CTV>;        \
        st = set_smem(kern, smem);                                      \
        if (st) return st;                                              \
        kern<<<grid, kNT, smem, s>>>(a);                                \
    } while (0)
    if (f->complex_taps) {
        if (PS == 4) LAUNCH_INT(8, true, 4, true);
        else if (PS == 2) LAUNCH_INT(8, true, 2, true);
        else LAUNCH_INT(8, true, 1, true);
    } else if (PS == 4) {
        if (f->packed) LAUNCH_INT(kR, true, 4, false); else LAUNCH_INT(kR, false, 4, false);
    } else if (PS == 2) {
        if (f->packed) LAUNCH_INT(kR, true, 2, false); else LAUNCH_INT(kR, false, 2, false);
    } else {
        if (f->packed) LAUNCH_INT(kR, true, 1, false); else LAUNCH_INT(kR, false, 1, false);
    }
#undef LAUNCH_INT
    SGPU_LAUNCH_CHECK();
    count_launch();
    f->hist_written = true;
    return SGPU_OK;
}
}  // namespace

SGPU_EXPORT int sgpu_interp_execute_block(sgpu_interp *f, const float *in, size_t n_in, size_t in_stride,
                                          float *out, size_t out_stride, size_t *n_out_p, sgpu_mem mem,
                                          void *stream) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    const size_t n_out = n_in * f->L;
    if (n_out_p) *n_out_p = n_out;
    if (n_in == 0) return SGPU_OK;
    if (!in || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (f->C > 1 && in_stride < n_in) return fail(SGPU_ERR_INVALID_ARGUMENT, "in_stride < n_in");
    if (out_stride < n_out) return fail(SGPU_ERR_CAPACITY, "out capacity %zu < %zu outputs", out_stride, n_out);
    if (ranges_overlap(in, ((f->C - 1) * in_stride + n_in) * 8, out, ((f->C - 1) * out_stride + n_out) * 8))
        return fail(SGPU_ERR_INVALID_ARGUMENT, "in and out overlap: execute_block is not an in-place operation");
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    auto run = [f](const float2 *d_in, size_t nc, long long istr, float2 *d_out, long long ostr, size_t nout,
                   cudaStream_t st_) -> int {
        f->hist_written = false;
        int st = interp_launch(f, d_in, (long long)nc, istr, d_out, ostr, (long long)nout, st_);
        if (st) return st;
        if (f->hist_written) {
            f->cur ^= 1;
            return SGPU_OK;
        }
        return enqueue_hist_update(d_in, istr, (long long)nc, f->d_hist, f->cur, f->C, f->S, st_);
    };
    if (mem == SGPU_DEVICE)
        return run(reinterpret_cast<const float2 *>(in), n_in, (long long)in_stride, reinterpret_cast<float2 *>(out),
                   (long long)out_stride, n_out, s);
    const size_t L = f->L;
    return host_pipeline(f->pipe, f->C, in, n_in, in_stride, out, out_stride, L,
                         [L](size_t nc) { return nc * L; }, run, s);
}

SGPU_EXPORT int sgpu_interp_push(sgpu_interp *f, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
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
    int st = enqueue_hist_update(d_in, istr, (long long)n_in, f->d_hist, f->cur, f->C, f->S, s);
    if (st) return st;
    if (mem == SGPU_HOST) SGPU_CUDA(cudaStreamSynchronize(s));
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_interp_execute_phase(sgpu_interp *f, size_t index, float *out, sgpu_mem mem, void *stream) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    if (index >= f->L) return fail(SGPU_ERR_INVALID_ARGUMENT, "phase index %zu >= %zu filters", index, f->L);
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    float2 *d_out = reinterpret_cast<float2 *>(out);
    if (mem == SGPU_HOST) {
        int st = f->stage.ensure(0, f->C * sizeof(float2));
        if (st) return st;
        d_out = (float2 *)f->stage.out;
    }
    const int tw = f->complex_taps ? 2 : 1;
    pfb_phase_kernel<<<(unsigned)ceil_div(f->C, 128), 128, 0, s>>>(f->d_hist[f->cur], (int)f->S, f->d_taps,
                                                                     f->per_channel ? (long long)f->img_floats : 0, f->Qpad, tw,
                                                                     (int)index, d_out, (int)f->C);
    SGPU_LAUNCH_CHECK();
    count_launch();
    if (mem == SGPU_HOST) {
        SGPU_CUDA(cudaMemcpyAsync(out, d_out, f->C * sizeof(float2), cudaMemcpyDeviceToHost, s));
        SGPU_CUDA(cudaStreamSynchronize(s));
    }
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_interp_get_state(sgpu_interp *f, float *history) {
    if (!f || !history) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    // expose the S-1 samples that influence future outputs (drop the oldest of the S kept)
    if (f->S > 1)
        SGPU_CUDA(cudaMemcpy2D(history, (f->S - 1) * sizeof(float2), f->d_hist[f->cur] + 1, f->S * sizeof(float2),
                               (f->S - 1) * sizeof(float2), f->C, cudaMemcpyDeviceToHost));
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_interp_set_state(sgpu_interp *f, const float *history) {
    if (!f || !history) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_hist[f->cur], 0, f->C * f->S * sizeof(float2)));
    if (f->S > 1)
        SGPU_CUDA(cudaMemcpy2D(f->d_hist[f->cur] + 1, f->S * sizeof(float2), history, (f->S - 1) * sizeof(float2),
                               (f->S - 1) * sizeof(float2), f->C, cudaMemcpyHostToDevice));
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_interp_reset(sgpu_interp *f) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_hist[f->cur], 0, f->C * f->S * sizeof(float2)));
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_interp_clone(const sgpu_interp *f, sgpu_interp **out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    // rebuild the effective tap vector hpad[p + (S-1-j)*L] = phase_taps[p][j]
    const size_t tw = f->complex_taps ? 2 : 1, per = f->L * f->S * tw, nimg = f->per_channel ? f->C : 1;
    std::vector<double> eff(per * nimg);
    for (size_t ch = 0; ch < nimg; ++ch)
        for (size_t p = 0; p < f->L; ++p)
            for (size_t j = 0; j < f->S; ++j)
                for (size_t c2 = 0; c2 < tw; ++c2)
                    eff[ch * per + (p + (f->S - 1 - j) * f->L) * tw + c2] = (double)f->phase_taps[ch * per + (p * f->S + j) * tw + c2];
    sgpu_interp *c = nullptr;
    int st = interp_create_common(eff.data(), per, f->T, f->C, f->L, f->S, f->scale_re, f->scale_im,
                                  f->complex_taps, f->per_channel, &c);
    if (st) return st;
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemcpy(c->d_hist[c->cur], f->d_hist[f->cur], f->C * f->S * sizeof(float2),
                         cudaMemcpyDeviceToDevice));
    *out = c;
    return SGPU_OK;
}
