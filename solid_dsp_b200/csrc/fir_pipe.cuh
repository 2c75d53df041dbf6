// Persistent, multi-stage versions of the decimator and interpolator kernels.
//
// ncu on the one-tile-per-block kernels (profiles/r1b_decim_interp_kernels.md) showed the FMA pipe
// only 56 % / 49 % active with the tile-load latency exposed: three single-buffered blocks per SM
// cannot keep ~100 KB of loads in flight while another block computes.  Here a block stays resident
// and walks its share of the tiles through an NSTAGE-deep ring of shared-memory stages filled by
// cp.async (LDGSTS): while tile i is being computed, tiles i+1 .. i+NSTAGE-1 are in flight.  One
// __syncthreads per tile; outputs go straight from registers to global memory (the PS lanes of a
// group own adjacent 16-byte pieces, so every store instruction writes whole 32-byte sectors).
#pragma once

#include <cstdint>

#include "fir_core.cuh"

namespace sgpu {

struct FirPipeArgs {
    const float2 *in;
    float2 *out;
    const float2 *hist;
    const float *taps;
    long long in_stride, out_stride;
    long long n_in, n_out;  // per channel
    long long total_tiles;
    int tiles_per_ch;
    int T;      // history convention: hist holds T-1 samples per channel
    int M;      // decimation M or interpolation L
    int c0;     // decimator phase on entry
    int Qpad;   // taps per phase, multiple of R
    int RS;     // plane row stride in float4 (odd)
    int vec_out;
    float scale_re;
};

__device__ __forceinline__ void pipe_cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pipe_cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pipe_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pipe_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ float2 pipe_fetch(const float2 *__restrict__ x, const float2 *__restrict__ hist,
                                             const long long i, const long long n_in, const int T) {
    if (i >= 0) return i < n_in ? x[i] : make_float2(0.f, 0.f);
    const long long h = (long long)(T - 1) + i;
    return h >= 0 ? hist[h] : make_float2(0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// Decimator (M >= 2, NT % (R*M) == 0).  See fir.cu / DESIGN.md 4.3 for the phase decomposition.
// MP = phases per lane (M == PS * MP), a template parameter so the phase loop is fully unrolled and
// ptxas can hoist the next phase's first window loads under the current phase's FFMA2 stream.
template <int R, bool PACKED, int NT, int PS, int MP, int NSTAGE, int MINB>
__global__ void __launch_bounds__(NT, MINB) fir_decim_pipe_kernel(const FirPipeArgs a) {
    extern __shared__ float4 smem[];
    constexpr int OT = NT / PS;
    constexpr int TW = 1;
    const int tid = threadIdx.x;
    constexpr int M = PS * MP;
    const int Qpad = a.Qpad, HR = Qpad / R, rows = HR + OT, RS = a.RS;
    const int plane_f4 = (R / 2) * RS + 1;
    const int stage_f4 = M * plane_f4;
    float *taps_s = reinterpret_cast<float *>(smem + (size_t)NSTAGE * stage_f4);
    {
        const int n4 = M * (Qpad + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }
    // loader role: fixed (phase, position in row), walks down the rows
    const int rm = R * M;
    const int e_in = tid % rm, rem = e_in % M, j = e_in / M, lp = M - 1 - rem;
    const int row0_ld = tid / rm, row_step = NT / rm;
    const int dst_off = ((lp * plane_f4 + (j >> 1) * RS + row0_ld) << 1) + (j & 1);  // float2 units
    const long long tile_in = (long long)rows * rm;                                    // samples per tile load

    auto issue = [&](const long long tile, const int stg) {
        const unsigned ch = (unsigned)tile / (unsigned)a.tiles_per_ch;
        const long long ti = (unsigned)tile - ch * (unsigned)a.tiles_per_ch;
        const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
        const long long i_lo = (ti * (OT * R) - Qpad) * M - a.c0;
        float2 *dst = reinterpret_cast<float2 *>(smem + (size_t)stg * stage_f4) + dst_off;
        if (i_lo >= 0 && i_lo + tile_in <= a.n_in) {
            const float2 *src = x + i_lo + tid;
            for (int rho = row0_ld; rho < rows; rho += row_step) {
                pipe_cp_async8(dst, src);
                dst += 2 * row_step;
                src += NT;
            }
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            long long i = i_lo + tid;
            for (int rho = row0_ld; rho < rows; rho += row_step) {
                if (i >= 0 && i < a.n_in) pipe_cp_async8(dst, x + i);
                else *dst = pipe_fetch(x, hist, i, a.n_in, a.T);
                dst += 2 * row_step;
                i += NT;
            }
        }
        pipe_commit();
    };

    const long long stride = gridDim.x;
    long long tile = blockIdx.x;
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (tile + s * stride < a.total_tiles) issue(tile + s * stride, s);
        else pipe_commit();
    }
    const int ot = tid / PS, part = tid % PS;
    const int row0 = HR + ot;
    const int nchunks = Qpad / R;
    constexpr int RP = R / PS;  // outputs this lane stores
    int stg = 0;
    for (; tile < a.total_tiles; tile += stride) {
        pipe_wait<NSTAGE - 2>();
        __syncthreads();  // tile landed for everyone; everyone is done with the stage refilled next
        {
            const long long nxt = tile + (NSTAGE - 1) * stride;
            const int ns = stg == 0 ? NSTAGE - 1 : stg - 1;
            if (nxt < a.total_tiles) issue(nxt, ns);
            else pipe_commit();
        }
        const float4 *st = smem + (size_t)stg * stage_f4;
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int sidx = 0; sidx < MP; ++sidx) {
            const int p = part * MP + sidx;
            fir_core<R, PACKED>(acc, st + (size_t)p * plane_f4, RS, row0, taps_s + (size_t)p * (Qpad + kTapSkew), nchunks);
        }
        if constexpr (PS > 1) {
#pragma unroll
            for (int o = 1; o < PS; o <<= 1) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, o);
                    acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, o);
                }
            }
        }
        // lane `part` of the group stores outputs [part*RP, (part+1)*RP) of the run: 8*RP contiguous bytes
        const unsigned ch = (unsigned)tile / (unsigned)a.tiles_per_ch;
        const long long ti = (unsigned)tile - ch * (unsigned)a.tiles_per_ch;
        const long long o0 = ti * (OT * R) + (long long)ot * R + part * RP;
        float2 *__restrict__ y = a.out + (long long)ch * a.out_stride + o0;
        const float s = a.scale_re;
#pragma unroll
        for (int pp = 0; pp < PS; ++pp) {
            if (pp == part) {  // static register indices per branch
#pragma unroll
                for (int q = 0; q < RP; q += 2) {
                    const float2 u = acc[pp * RP + q], v = acc[pp * RP + q + 1];
                    if (a.vec_out && o0 + q + 1 < a.n_out) {
                        *reinterpret_cast<float4 *>(y + q) = make_float4(u.x * s, u.y * s, v.x * s, v.y * s);
                    } else {
                        if (o0 + q < a.n_out) y[q] = make_float2(u.x * s, u.y * s);
                        if (o0 + q + 1 < a.n_out) y[q + 1] = make_float2(v.x * s, v.y * s);
                    }
                }
            }
        }
        stg = stg + 1 == NSTAGE ? 0 : stg + 1;
    }
    pipe_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Interpolator: one input plane per stage; the PS lanes of a group split the L output phases and
// store their outputs directly (phase-adjacent lanes write adjacent 8-byte samples).
template <int R, bool PACKED, int NT, int PS, int NSTAGE, int MINB>
__global__ void __launch_bounds__(NT, MINB) fir_interp_pipe_kernel(const FirPipeArgs a) {
    extern __shared__ float4 smem[];
    constexpr int OT = NT / PS;
    constexpr int TW = 1;
    const int tid = threadIdx.x;
    const int L = a.M, Qpad = a.Qpad, HR = Qpad / R, rows = HR + OT, RS = a.RS;
    const int stage_f4 = (R / 2) * RS + 1;
    float *taps_s = reinterpret_cast<float *>(smem + (size_t)NSTAGE * stage_f4);
    {
        const int n4 = L * (Qpad + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }
    const int total_pairs = rows * R / 2;
    const long long tile_in = (long long)rows * R;

    auto issue = [&](const long long tile, const int stg) {
        const unsigned ch = (unsigned)tile / (unsigned)a.tiles_per_ch;
        const long long ti = (unsigned)tile - ch * (unsigned)a.tiles_per_ch;
        const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
        const long long i_lo = ti * (OT * R) - Qpad;
        float4 *plane = smem + (size_t)stg * stage_f4;
        const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        if (i_lo >= 0 && i_lo + tile_in <= a.n_in) {
            for (int pe = tid; pe < total_pairs; pe += NT) {
                float4 *dst = plane + (pe % (R / 2)) * RS + pe / (R / 2);
                const float2 *src = x + i_lo + 2 * pe;
                if (vec) pipe_cp_async16(dst, src);
                else {
                    pipe_cp_async8(dst, src);
                    pipe_cp_async8(reinterpret_cast<float2 *>(dst) + 1, src + 1);
                }
            }
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            for (int pe = tid; pe < total_pairs; pe += NT) {
                const long long i = i_lo + 2 * pe;
                const float2 s0 = pipe_fetch(x, hist, i, a.n_in, a.T);
                const float2 s1 = pipe_fetch(x, hist, i + 1, a.n_in, a.T);
                plane[(pe % (R / 2)) * RS + pe / (R / 2)] = make_float4(s0.x, s0.y, s1.x, s1.y);
            }
        }
        pipe_commit();
    };

    const long long stride = gridDim.x;
    long long tile = blockIdx.x;
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (tile + s * stride < a.total_tiles) issue(tile + s * stride, s);
        else pipe_commit();
    }
    const int ot = tid / PS, part = tid % PS;
    const int row0 = HR + ot;
    const int nchunks = Qpad / R;
    const int Lp = (L + PS - 1) / PS;
    int stg = 0;
    for (; tile < a.total_tiles; tile += stride) {
        pipe_wait<NSTAGE - 2>();
        __syncthreads();
        {
            const long long nxt = tile + (NSTAGE - 1) * stride;
            const int ns = stg == 0 ? NSTAGE - 1 : stg - 1;
            if (nxt < a.total_tiles) issue(nxt, ns);
            else pipe_commit();
        }
        const float4 *st = smem + (size_t)stg * stage_f4;
        const unsigned ch = (unsigned)tile / (unsigned)a.tiles_per_ch;
        const long long ti = (unsigned)tile - ch * (unsigned)a.tiles_per_ch;
        const long long n0 = ti * (OT * R) + (long long)ot * R;  // first input position of this run
        float2 *__restrict__ y = a.out + (long long)ch * a.out_stride + n0 * L;
        const long long room = a.n_out - n0 * L;  // outputs of this channel from y on
        // phases interleaved over the lanes of a group (p = part, part+PS, ...): lanes store adjacent samples
        for (int sidx = 0; sidx < Lp; ++sidx) {
            const int p = sidx * PS + part;
            if (p >= L) break;
            float2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
            fir_core<R, PACKED>(acc, st, RS, row0, taps_s + (size_t)p * (Qpad + kTapSkew), nchunks);
            float2 *yp = y + p;
            if ((long long)R * L <= room) {
#pragma unroll
                for (int r = 0; r < R; ++r) yp[(long long)r * L] = acc[r];  // no scale: pfb.rs:85-90
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((long long)r * L + p < room) yp[(long long)r * L] = acc[r];
            }
        }
        stg = stg + 1 == NSTAGE ? 0 : stg + 1;
    }
    pipe_wait<0>();
}

}  // namespace sgpu
