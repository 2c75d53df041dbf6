// "Walking" polyphase kernels for short sub-filters (taps per phase <= 2R = 32).
//
// ncu on the tile kernels (profiles/r1e_kernels.md) showed the FMA pipe 57-65 % active with ~0.6
// non-FMA instructions per FFMA2: every run of R outputs re-loaded its whole register window (three
// rows of R samples for a 32-tap sub-filter) and paid the tile's fixed costs (index math, barriers,
// staging) for only 512-1024 FFMA2 per thread.  Here a thread owns K consecutive runs and walks them
// backwards in time: run s needs rows (rho-s, rho-s-1, rho-s-2) of which the last two are the next
// run's first two, so the 2R-slot register window slides by ONE row load (R/2 LDS.128) per 2R*R
// complex MACs, the tile's fixed costs are spread over K runs, and the outputs go from registers
// straight to global memory in whole 32-byte sectors.
//
// Shared-memory plane layout and fir_chunk are those of fir_core.cuh.  Groups of a warp sit K rows
// apart, so K is odd to keep the lanes' LDS.128 conflict free.
//
// Included by fir.cu inside namespace sgpu::<anonymous>, after FirArgs / fetch_sample / cp_async*.
#pragma once

// one run of R outputs over a 2R-tap sub-filter.  PAR = 0: the run's newest row sits in W[0, R) and
// the row before it in W[R, 2R); PAR = 1: the other way round.  `next_row` (two rows older than
// the run's newest) replaces the newest row between the two tap chunks.
template <int R, int PAR>
__device__ __forceinline__ void walk_run(float2 (&acc)[R], float2 (&W)[2 * R], const float4 *__restrict__ plane,
                                         const int RS, const int next_row, const float *__restrict__ taps) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
    fir_chunk<R, true, PAR ? R : 0, false>(acc, W, taps);
    load_row<R, PAR ? R : 0>(W, plane, RS, next_row);
    fir_chunk<R, true, PAR ? 0 : R, false>(acc, W, taps + R);
}

// --------------------------------------------------------------------------------------------
// Interpolator, L phases = L lanes per group (L in {2, 4, 8}), sub-filter length S <= 2R.
//   y[n*L + p] = sum_{j<S} hp[p][j] * x[n-j]   (fir/pfb.rs:85-90, fir/interp.rs:102-111; no scale)
// Block = NT threads = NT/L groups; group g owns input positions [g*K*R, (g+1)*K*R) of the tile and
// its lane p produces phase p of their outputs.
template <int R, int L, int K, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) fir_interp_walk_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    static_assert(K % 2 == 1, "K must be odd (bank conflicts, parity of the last run)");
    constexpr int G = NT / L, HR = 2, ROWS = HR + G * K, QP = 2 * R;
    const int tid = threadIdx.x;
    const int RS = a.RS;
    float *taps_s = reinterpret_cast<float *>(smem + (R / 2) * RS + 1);
    {
        const int n4 = L * (QP + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }
    const int ch = blockIdx.y;
    const long long n_base = (long long)blockIdx.x * (G * K * R);  // first input position of the tile
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    {   // tile load: ROWS * R consecutive samples from n_base - QP on, pairs -> transposed rows
        const long long i_lo = n_base - QP;
        constexpr int total_pairs = ROWS * R / 2;
        if (i_lo >= 0 && i_lo + (long long)ROWS * R <= a.n_in && a.vec_in) {
            const float2 *src = x + i_lo;
            for (int pe = tid; pe < total_pairs; pe += NT)
                cp_async16(smem + (pe % (R / 2)) * RS + pe / (R / 2), src + 2 * pe);
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            for (int pe = tid; pe < total_pairs; pe += NT) {
                const long long i = i_lo + 2 * pe;
                const float2 s0 = fetch_sample(x, hist, i, a.n_in, a.T);
                const float2 s1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
                smem[(pe % (R / 2)) * RS + pe / (R / 2)] = make_float4(s0.x, s0.y, s1.x, s1.y);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int g = tid / L, p = tid % L;
    const float *tp = taps_s + p * (QP + kTapSkew);
    int row = HR + g * K + (K - 1);                          // newest row of the group's newest run
    long long n0 = n_base + (long long)(g * K + (K - 1)) * R;  // its first input position
    float2 *__restrict__ yp = a.out + (long long)ch * a.out_stride + n0 * L + p;
    float2 W[2 * R], acc[R];
    load_row<R, 0>(W, smem, RS, row);
    load_row<R, R>(W, smem, RS, row - 1);
    auto store = [&]() {
        if (n0 + R <= a.n_in) {
#pragma unroll
            for (int r = 0; r < R; ++r) yp[r * L] = acc[r];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (n0 + r < a.n_in) yp[r * L] = acc[r];
        }
        n0 -= R;
        yp -= R * L;
    };
#pragma unroll 1
    for (int it = 0; it < (K - 1) / 2; ++it) {
        walk_run<R, 0>(acc, W, smem, RS, row - 2, tp);
        store();
        walk_run<R, 1>(acc, W, smem, RS, row - 3, tp);
        store();
        row -= 2;
    }
    walk_run<R, 0>(acc, W, smem, RS, row - 2, tp);  // rows 0/1 of the plane are the halo: row - 2 >= 0
    store();
}
