// Register-window FIR core shared by the FIR, decimator and interpolator kernels.
//
// A thread owns R consecutive outputs of one (phase) sequence and walks the taps in chunks of
// R.  The 2R-slot circular register window W holds the samples the current chunk needs:
// before chunk c the thread loads ONE row of R samples (R/2 LDS.128) into the half of W
// that the previous chunk no longer reads, so the steady state is R samples loaded per R*R
// complex MACs.  Slots are addressed statically (the chunk pair is fully unrolled), so W and
// the accumulators live in registers.
//
// Shared-memory plane layout ("transposed rows"): row rho holds R consecutive samples; the
// pair (2jj, 2jj+1) of row rho is the float4 plane[jj * RS + rho].  Consecutive lanes own
// consecutive rows, so every LDS.128 of a warp touches 32 consecutive float4 -- conflict-free
// -- and RS is odd so the tile loader's STS (8 lanes = one row's 8 pairs) is conflict-free too.
//
// Arithmetic: real taps on complex samples.  PACKED = true issues one fma.rn.f32x2 (SASS
// FFMA2, new on sm_100) per complex MAC with the tap as its scalar-broadcast operand;
// PACKED = false issues two scalar FFMA.  Accumulation order is newest sample first, as
// dot_product/mod.rs:159-170 does.
#pragma once

#include <cuda_runtime.h>

namespace sgpu {

// Tap image rows (one per phase) are Qpad + kTapSkew floats apart: the 16-byte skew puts the PS
// different tap rows that the lanes of a phase-split warp read into different banks.
constexpr int kTapSkew = 4;

template <bool PACKED>
__device__ __forceinline__ void cmac_real(float2 &acc, const float2 w, const float gx, const float gy) {
    if constexpr (PACKED) {
        acc = __ffma2_rn(w, make_float2(gx, gy), acc);
    } else {
        acc.x = fmaf(gx, w.x, acc.x);
        acc.y = fmaf(gx, w.y, acc.y);
    }
}

template <int R, int BASE>
__device__ __forceinline__ void load_row(float2 (&W)[2 * R], const float4 *__restrict__ plane,
                                         const int RS, const int row) {
#pragma unroll
    for (int jj = 0; jj < R / 2; ++jj) {
        const float4 v = plane[jj * RS + row];
        W[BASE + 2 * jj] = make_float2(v.x, v.y);
        W[BASE + 2 * jj + 1] = make_float2(v.z, v.w);
    }
}

// One chunk of R taps.  OFF = 0 for even chunks, R for odd ones (see header comment).
// taps: R floats (real taps) or R (re, im) pairs (CT, complex taps) in shared memory, 16-byte
// aligned.  The packed path feeds the tap as FFMA2's scalar-broadcast operand (SASS
// `FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2`), so no duplicated (g,g) pairs are needed.
// Complex taps (a+bi)(c+di): accA += (c,d)*a, accB += (c,d)*b; the caller combines
// re = accA.x - accB.y, im = accA.y + accB.x once at the end.  acc holds accA in [0,R), accB in [R,2R).
template <int R, bool PACKED, int OFF, bool CT>
__device__ __forceinline__ void fir_chunk(float2 (&acc)[CT ? 2 * R : R], const float2 (&W)[2 * R],
                                          const float *__restrict__ taps) {
    const float4 *t4 = reinterpret_cast<const float4 *>(taps);
    if constexpr (CT) {
#pragma unroll
        for (int u2 = 0; u2 < R / 2; ++u2) {
            const float4 g = t4[u2];  // (a_u, b_u, a_{u+1}, b_{u+1})
            const float ga[2] = {g.x, g.z}, gb[2] = {g.y, g.w};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 w = W[(r - (2 * u2 + k) + OFF + 4 * R) & (2 * R - 1)];
                    cmac_real<PACKED>(acc[r], w, ga[k], ga[k]);
                    cmac_real<PACKED>(acc[R + r], w, gb[k], gb[k]);
                }
            }
        }
    } else {
#pragma unroll
        for (int u4 = 0; u4 < R / 4; ++u4) {
            const float4 g = t4[u4];
            const float gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    cmac_real<PACKED>(acc[r], W[(r - (4 * u4 + k) + OFF + 4 * R) & (2 * R - 1)], gs[k], gs[k]);
            }
        }
    }
}

// acc[r] += sum_{k < R*nchunks} g[k] * seq[R*row0 + r - k]   (seq = the plane's sample sequence)
// nchunks = taps per phase / R.  Chunks run in pairs (the two halves of the register window swap
// roles), so nchunks is even -- except in the ONE instantiations, which are the single-chunk
// kernels for sub-filters of <= R taps (no loop at all).
template <int R, bool PACKED, bool CT = false, bool ONE = false>
__device__ __forceinline__ void fir_core(float2 (&acc)[CT ? 2 * R : R], const float4 *__restrict__ plane,
                                         const int RS, const int row0,
                                         const float *__restrict__ taps, const int nchunks) {
    constexpr int TW = CT ? 2 : 1;  // floats per tap in shared memory
    float2 W[2 * R];
    load_row<R, 0>(W, plane, RS, row0);
    if constexpr (ONE) {
        load_row<R, R>(W, plane, RS, row0 - 1);
        fir_chunk<R, PACKED, 0, CT>(acc, W, taps);
    } else {
        int row = row0;
        for (int cp = 0; cp < (nchunks >> 1); ++cp) {
            load_row<R, R>(W, plane, RS, row - 1);
            fir_chunk<R, PACKED, 0, CT>(acc, W, taps);
            load_row<R, 0>(W, plane, RS, row - 2);
            fir_chunk<R, PACKED, R, CT>(acc, W, taps + R * TW);
            row -= 2;
            taps += 2 * R * TW;
        }
    }
}

}  // namespace sgpu
