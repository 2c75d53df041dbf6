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
// Included by fir.cu inside namespace sgpu::<anonymous>, after FirArgs / fetch_sample / cp_async* / hist_tail_update.
#pragma once

// one run of R outputs over a 2R-tap sub-filter.  PAR = 0: the run's newest row sits in W[0, R) and
// the row before it in W[R, 2R); PAR = 1: the other way round.  `next_row` (two rows older than
// the run's newest) replaces the newest row between the two tap chunks.
template <int R, int PAR, int NC = 2>
__device__ __forceinline__ void walk_run(float2 (&acc)[R], float2 (&W)[2 * R], const float4 *__restrict__ plane,
                                         const int RS, const int next_row, const float *__restrict__ taps) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
    fir_chunk<R, true, PAR ? R : 0, false>(acc, W, taps);
    load_row<R, PAR ? R : 0>(W, plane, RS, next_row);
    // NC = 1 (sub-filters of <= R taps): the run ends here and the row just loaded is the next run's
    // older row; NC = 2: second tap chunk over (newest-1, newest-2)
    if constexpr (NC == 2) fir_chunk<R, true, PAR ? 0 : R, false>(acc, W, taps + R);
}

// --------------------------------------------------------------------------------------------
// Interpolator, L phases = L lanes per group (L in {2, 4, 8}), sub-filter length S <= 2R.
//   y[n*L + p] = sum_{j<S} hp[p][j] * x[n-j]   (fir/pfb.rs:85-90, fir/interp.rs:102-111; no scale)
// Every WARP runs its own software pipeline over TPW consecutive tiles of its channel: two private
// shared-memory stages filled with cp.async, __syncwarp only -- no block barrier after the taps are
// in, so the warps of an SM drift apart and the FMA pipe always finds one in its arithmetic phase.
// A warp tile = 32/L groups x K runs x R input positions; group g owns positions [g*K*R, (g+1)*K*R)
// and its lane p produces phase p of their outputs.
template <int R, int L, int K, int NT, int MINB, int TPW, int NC = 2>
__global__ void __launch_bounds__(NT, MINB) fir_interp_walk_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    static_assert(K % 2 == 1, "K must be odd (bank conflicts, parity of the last run)");
    static_assert(NC == 1 || NC == 2, "sub-filters of NC * R taps");
    constexpr int G = 32 / L, HR = NC, ROWS = HR + G * K, QP = NC * R, TILE = G * K * R, NW = NT / 32;
    constexpr int RS = ROWS | 1, STAGE_F4 = (R / 2) * RS + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *taps_s = reinterpret_cast<float *>(smem + NW * 2 * STAGE_F4);
    const int ch = blockIdx.y;
    hist_tail_update(a, ch, tid, NT);
    {
        const int n4 = L * (QP + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps + (long long)ch * a.tap_stride);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    float4 *stage0 = smem + warp * 2 * STAGE_F4;
    auto issue = [&](const long long n_base, float4 *plane) {
        // ROWS * R consecutive samples from n_base - QP on, pairs -> transposed rows
        const long long i_lo = n_base - QP;
        constexpr int total_pairs = ROWS * R / 2;
        if (i_lo >= 0 && i_lo + (long long)ROWS * R <= a.n_in && a.vec_in) {
            const float2 *src = x + i_lo;
#pragma unroll 4
            for (int pe = lane; pe < total_pairs; pe += 32)
                cp_async16(plane + (pe % (R / 2)) * RS + pe / (R / 2), src + 2 * pe);
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            for (int pe = lane; pe < total_pairs; pe += 32) {
                const long long i = i_lo + 2 * pe;
                const float2 s0 = fetch_sample(x, hist, i, a.n_in, a.T);
                const float2 s1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
                plane[(pe % (R / 2)) * RS + pe / (R / 2)] = make_float4(s0.x, s0.y, s1.x, s1.y);
            }
        }
    };
    const int g = lane / L, p = lane % L;
    const float *tp = taps_s + p * (QP + kTapSkew);
    // first input position of the warp's current tile
    // the block covers NW * TPW consecutive tiles; its warps take them round-robin, so neighbouring tiles
    // (which share the halo) are in flight at the same time and the halo is an L2 hit
    constexpr long long STEP = (long long)NW * TILE;
    long long n_base = ((long long)blockIdx.x * NW * TPW + warp) * (long long)TILE;
    if (n_base >= a.n_in) return;
    issue(n_base, stage0);
#pragma unroll 1
    for (int t = 0; t < TPW && n_base < a.n_in; ++t, n_base += STEP) {
        cp_async_wait_all();
        __syncwarp();  // tile t landed; every lane is done with the other stage
        const float4 *plane = stage0 + (t & 1) * STAGE_F4;
        if (t + 1 < TPW && n_base + STEP < a.n_in) issue(n_base + STEP, stage0 + ((t + 1) & 1) * STAGE_F4);
        int row = HR + g * K + (K - 1);                          // newest row of the group's newest run
        long long n0 = n_base + (long long)(g * K + (K - 1)) * R;  // its first input position
        float2 *__restrict__ yp = a.out + (long long)ch * a.out_stride + n0 * L + p;
        float2 W[2 * R], acc[R];
        load_row<R, 0>(W, plane, RS, row);
        load_row<R, R>(W, plane, RS, row - 1);
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
            walk_run<R, 0, NC>(acc, W, plane, RS, row - 2, tp);
            store();
            walk_run<R, 1, NC>(acc, W, plane, RS, row - 3, tp);
            store();
            row -= 2;
        }
        // the halo rows keep row - 2 >= 0 for NC = 2; for NC = 1 the last row load is not needed: clamp
        walk_run<R, 0, NC>(acc, W, plane, RS, max(row - 2, 0), tp);
        store();
    }
}

// --------------------------------------------------------------------------------------------
// Decimator with warp-private tiles (M in {2, 4, 8}, any sub-filter length that fits).
//   y[m] = scale * sum_p sum_q g[q*M+p] * x[(m-q)*M + (M-1-c0) - p]          (fir/decim.rs:221-228)
// Same phase decomposition as fir_decim_kernel, but every warp owns its tiles: a private stage of
// M phase planes filled by 8-byte cp.async (the de-interleave), __syncwarp only, so the warps of an
// SM drift apart and tile loads hide behind other warps' arithmetic.  PS adjacent lanes share a run
// of R outputs and each walks M/PS planes.  The partial sums meet in shared memory (the stage is
// free by then): a lane writes its R sums as R/2 float4, then reads and adds the PS partials of
// the pieces it stores -- every PS-th 16-byte piece of the run, so each store instruction writes
// whole 32-byte sectors.  8 STS + 8 LDS + 12 FADD2 instead of the 128 SHFL/FADD of a butterfly.
// MIX: the NCO mix-down of a digital down-converter (nco/mod.rs:141-172) is applied to the tile IN shared memory, by
// the lane that loaded the element, right after the asynchronous copies have landed: the mixed stream never exists in
// HBM and the loads stay asynchronous.  History samples (in front of the call) are already mixed.
template <int R, int M, int PS, int NW, int MINB, int TPW, bool ONE = false, bool MIX = false>
__global__ void __launch_bounds__(NW * 32, MINB) fir_decim_warp_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    constexpr int G = 32 / PS, MP = M / PS, RM = R * M, LPRW = RM / 32;  // LPRW: loader steps per row of all planes
    constexpr int NPC = R / 2 / PS;                                     // 16-byte pieces a lane stores
    static_assert(RM % 32 == 0 && M % PS == 0 && (R / 2) % PS == 0, "shape");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Qpad = a.Qpad, HR = Qpad / R, rows = HR + G, RS = a.RS;
    const int plane_f4 = (R / 2) * RS + 1, stage_f4 = max(M * plane_f4, 32 * (R / 2 + 1));  // room for the reduction
    float *taps_s = reinterpret_cast<float *>(smem + (size_t)NW * stage_f4);
    const int ch = blockIdx.y;
    hist_tail_update(a, ch, tid, NW * 32);
    const int n4 = M * (Qpad + kTapSkew) / 4;
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.taps + (long long)ch * a.tap_stride);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NW * 32) dst[i] = src[i];
    }
    const float2 *lut_s = reinterpret_cast<const float2 *>(reinterpret_cast<float4 *>(taps_s) + n4);  // MIX: 1024 x (cos, sin)
    if constexpr (MIX) {
        float4 *dst = reinterpret_cast<float4 *>(taps_s) + n4;
        const float4 *src = reinterpret_cast<const float4 *>(a.lut);
        for (int i = tid; i < 512; i += NW * 32) dst[i] = src[i];
    }
    __syncthreads();
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    float4 *stage = smem + (size_t)warp * stage_f4;
    // loader role: element e = k*32 + lane of a row-of-all-planes (RM consecutive input samples)
    // goes to plane M-1-(e%M), position e/M of the row; the next row is 16 bytes further on
    unsigned sdst[LPRW];  // shared-space byte addresses, row 0
    int eoff[LPRW];       // float2 offsets, for the guarded path
#pragma unroll
    for (int k = 0; k < LPRW; ++k) {
        const int e = k * 32 + lane, rem = e % M, j = e / M, p = M - 1 - rem;
        eoff[k] = ((p * plane_f4 + (j >> 1) * RS) << 1) + (j & 1);
        sdst[k] = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float2 *>(stage) + eoff[k]);
    }
    auto issue = [&](const long long m_base) {
        const long long i_lo = (m_base - Qpad) * M - a.c0;
        if (i_lo >= 0 && i_lo + (long long)rows * RM <= a.n_in) {
            const float2 *src = x + i_lo + lane;
#pragma unroll 2
            for (int rho = 0; rho < rows; ++rho) {
#pragma unroll
                for (int k = 0; k < LPRW; ++k)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst[k] + 16u * rho), "l"(src + k * 32)
                                 : "memory");
                src += RM;
            }
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            float2 *base = reinterpret_cast<float2 *>(stage);
            long long i = i_lo + lane;
            for (int rho = 0; rho < rows; ++rho) {
#pragma unroll
                for (int k = 0; k < LPRW; ++k) {
                    const long long ii = i + k * 32;
                    if (ii >= 0 && ii < a.n_in) cp_async8(base + eoff[k] + 2 * rho, x + ii);
                    else base[eoff[k] + 2 * rho] = fetch_sample(x, hist, ii, a.n_in, a.T);
                }
                i += RM;
            }
        }
    };
    // MIX: every sample of the tile that belongs to this call's input is multiplied by conj(phasor) IN PLACE, one
    // float4 (positions j, j + 1 of one phase plane = input samples i, i + M) per step: lane -> (plane, position pair),
    // rows in batches of five so that the shared-memory latencies overlap.
    unsigned nco_theta = 0, nco_delta = 0;
    if constexpr (MIX) nco_channel(a, ch, nco_theta, nco_delta);
    auto mix_tile = [&](const long long m_base) {
        const long long i_lo = (m_base - Qpad) * M - a.c0;
        const bool interior = i_lo >= 0 && i_lo + (long long)rows * RM <= a.n_in;
        const unsigned d_row = (unsigned)RM * nco_delta, d_m = (unsigned)M * nco_delta;
        constexpr int COMBOS = M * (R / 2);
#pragma unroll
        for (int c0 = 0; c0 < COMBOS; c0 += 32) {
            const int c = c0 + lane;
            if (COMBOS % 32 != 0 && c >= COMBOS) break;
            const int p = c % M, jp = c / M;
            float4 *cell = stage + p * plane_f4 + jp * RS;
            const int e0 = 2 * jp * M + (M - 1 - p);  // offset of the pair's first sample inside a row of all planes
            unsigned th = nco_theta + (unsigned)(i_lo + e0) * nco_delta + (1u << 21);  // rounding folded in
            long long i0 = i_lo + e0;
            auto mix1 = [&](float &xr, float &xi, const float2 cs) {  // conj(c + j s) x  (nco/mod.rs:147-151)
                const float r = fmaf(cs.x, xr, cs.y * xi), i = fmaf(cs.x, xi, -cs.y * xr);
                xr = r;
                xi = i;
            };
            constexpr int NB = 5;  // rows per batch: 5 cells and 10 table entries in flight per lane
            int rho = 0;
            if (interior) {
                for (; rho + NB <= rows; rho += NB) {
                    float4 v[NB];
                    float2 t0[NB], t1[NB];
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        v[b] = cell[rho + b];
                        t0[b] = lut_s[(th + (unsigned)b * d_row) >> 22];
                        t1[b] = lut_s[(th + (unsigned)b * d_row + d_m) >> 22];
                    }
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        mix1(v[b].x, v[b].y, t0[b]);
                        mix1(v[b].z, v[b].w, t1[b]);
                        cell[rho + b] = v[b];
                    }
                    th += NB * d_row;
                }
                i0 += (long long)rho * RM;
            }
            for (; rho < rows; ++rho) {
                float4 v = cell[rho];
                if (interior || (i0 >= 0 && i0 < a.n_in)) mix1(v.x, v.y, lut_s[th >> 22]);
                if (interior || (i0 + M >= 0 && i0 + M < a.n_in)) mix1(v.z, v.w, lut_s[(th + d_m) >> 22]);
                cell[rho] = v;
                th += d_row;
                i0 += RM;
            }
        }
    };
    const int g = lane / PS, part = lane % PS;
    const int nchunks = Qpad / R;
    constexpr int TILE = G * R;  // outputs per warp tile
    constexpr int RED = R / 2 + 1;  // float4 pitch of a lane's partial sums
    constexpr long long STEP = (long long)NW * TILE;  // warps take the block's tiles round-robin (halo = L2 hit)
    long long m_base = ((long long)blockIdx.x * NW * TPW + warp) * (long long)TILE;
    if (m_base >= a.n_out) return;
    issue(m_base);
#pragma unroll 1
    for (int t = 0; t < TPW && m_base < a.n_out; ++t, m_base += STEP) {
        cp_async_wait_all();
        if constexpr (MIX) {
            __syncwarp();  // the tile has landed for every lane
            mix_tile(m_base);
        }
        __syncwarp();
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int sidx = 0; sidx < MP; ++sidx) {
            const int p = part * MP + sidx;
            fir_core<R, true, false, ONE>(acc, stage + (size_t)p * plane_f4, RS, HR + g, taps_s + (size_t)p * (Qpad + kTapSkew),
                                          nchunks);
        }
        float4 out[NPC];
        if constexpr (PS > 1) {
            __syncwarp();  // every lane is done reading the planes
#pragma unroll
            for (int q = 0; q < R / 2; ++q)
                stage[lane * RED + q] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NPC; ++k) {
                float2 lo = make_float2(0.f, 0.f), hi = lo;
#pragma unroll
                for (int sl = 0; sl < PS; ++sl) {
                    const float4 v = stage[(g * PS + sl) * RED + k * PS + part];
                    lo = __fadd2_rn(lo, make_float2(v.x, v.y));
                    hi = __fadd2_rn(hi, make_float2(v.z, v.w));
                }
                out[k] = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NPC; ++k) out[k] = make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
        }
        __syncwarp();  // the stage is free again
        if (t + 1 < TPW && m_base + STEP < a.n_out) issue(m_base + STEP);
        // lane `part` stores the pieces k*PS + part, k < NPC: sector-complete per instruction
        const long long o0 = m_base + (long long)g * R;
        float2 *__restrict__ y = a.out + (long long)ch * a.out_stride + o0;
        const float s = a.scale_re;
        if (a.vec_out && o0 + R <= a.n_out) {
#pragma unroll
            for (int k = 0; k < NPC; ++k) {
                const float4 v = out[k];
                *reinterpret_cast<float4 *>(y + 2 * (k * PS + part)) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NPC; ++k) {
                const float4 v = out[k];
                const int q = 2 * (k * PS + part);
                if (o0 + q < a.n_out) y[q] = make_float2(v.x * s, v.y * s);
                if (o0 + q + 1 < a.n_out) y[q + 1] = make_float2(v.z * s, v.w * s);
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// Plain FIR (M = 1, real taps) with warp-private tiles.
//   y[n] = scale * sum_i h[T-1-i] * x[n-i]                                   (fir/mod.rs:209-212)
// A warp tile is 32 runs of R outputs (one per lane) plus the Qpad-sample halo in front of them; NS
// private stages filled by 16-byte cp.async, __syncwarp only.  The R outputs of a lane are turned
// into coalesced 16-byte stores through the stage the warp has just finished reading.
template <int R, int NW, int MINB, int TPW, int NS, bool ONE = false>
__global__ void __launch_bounds__(NW * 32, MINB) fir_warp_kernel(const FirArgs a) {
    extern __shared__ float4 smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Qpad = a.Qpad, HR = Qpad / R, rows = HR + 32, RS = a.RS;
    const int stage_f4 = max((R / 2) * RS + 1, 32 * (R / 2 + 1));  // room for the output transpose
    float *taps_s = reinterpret_cast<float *>(smem + (size_t)NW * NS * stage_f4);
    const int ch = blockIdx.y;
    // Programmatic dependent launch (short calls, csrc/fir.cu): the next call on the stream may be scheduled while this
    // one still runs; everything up to griddepcontrol.wait touches only the handle's constant tap image, so a call's
    // launch latency and tap staging hide behind its predecessor.  Both instructions are no-ops on an ordinary launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {
        const int n4 = (Qpad + kTapSkew) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(a.taps + (long long)ch * a.tap_stride);
        float4 *dst = reinterpret_cast<float4 *>(taps_s);
        for (int i = tid; i < n4; i += NW * 32) dst[i] = src[i];
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the predecessor has completed: its history and outputs are visible
    hist_tail_update(a, ch, tid, NW * 32);
    __syncthreads();
    const float2 *__restrict__ x = a.in + (long long)ch * a.in_stride;
    float4 *stage0 = smem + (size_t)warp * NS * stage_f4;
    const int total_pairs = rows * (R / 2);
    // loader role: pair pe = it*32 + lane -> row pe / (R/2), piece pe % (R/2)
    const int role = (lane % (R / 2)) * RS + lane / (R / 2);
    constexpr int ROWS_PER_IT = 32 / (R / 2);
    auto issue = [&](const long long m_base, float4 *plane) {
        const long long i_lo = m_base - Qpad;
        if (i_lo >= 0 && i_lo + (long long)rows * R <= a.n_in && a.vec_in) {
            const float2 *src = x + i_lo + 2 * lane;
            float4 *dst = plane + role;
#pragma unroll 4
            for (int pe = lane; pe < total_pairs; pe += 32) {
                cp_async16(dst, src);
                dst += ROWS_PER_IT;
                src += 64;
            }
        } else {
            const float2 *__restrict__ hist = a.hist + (long long)ch * (a.T - 1);
            for (int pe = lane; pe < total_pairs; pe += 32) {
                const long long i = i_lo + 2 * pe;
                const float2 s0 = fetch_sample(x, hist, i, a.n_in, a.T);
                const float2 s1 = fetch_sample(x, hist, i + 1, a.n_in, a.T);
                plane[(pe % (R / 2)) * RS + pe / (R / 2)] = make_float4(s0.x, s0.y, s1.x, s1.y);
            }
        }
    };
    const int nchunks = Qpad / R;
    constexpr int TILE = 32 * R;
    constexpr int RED = R / 2 + 1;
    constexpr long long STEP = (long long)NW * TILE;  // warps take the block's tiles round-robin (halo = L2 hit)
    long long m_base = ((long long)blockIdx.x * NW * TPW + warp) * (long long)TILE;
    if (m_base >= a.n_out) return;
    issue(m_base, stage0);
#pragma unroll 1
    for (int t = 0; t < TPW && m_base < a.n_out; ++t, m_base += STEP) {
        cp_async_wait_all();
        __syncwarp();
        float4 *st = stage0 + (NS == 2 ? (t & 1) * stage_f4 : 0);
        if (NS == 2 && t + 1 < TPW && m_base + STEP < a.n_out) issue(m_base + STEP, stage0 + ((t + 1) & 1) * stage_f4);
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
        fir_core<R, true, false, ONE>(acc, st, RS, HR + lane, taps_s, nchunks);
        // transpose: lane's R outputs -> rows of the stage -> 16-byte pieces, 4 runs (512 bytes) per instruction
        __syncwarp();
        const float s = a.scale_re;
#pragma unroll
        for (int q = 0; q < R / 2; ++q)
            st[lane * RED + q] = make_float4(acc[2 * q].x * s, acc[2 * q].y * s, acc[2 * q + 1].x * s, acc[2 * q + 1].y * s);
        __syncwarp();
        float2 *__restrict__ y = a.out + (long long)ch * a.out_stride + m_base;
        const int prow = lane / (R / 2), pq = lane % (R / 2);
        if (a.vec_out && m_base + TILE <= a.n_out) {
#pragma unroll
            for (int i = 0; i < R / 2; ++i) {
                const int run = i * ROWS_PER_IT + prow;
                *reinterpret_cast<float4 *>(y + run * R + 2 * pq) = st[run * RED + pq];
            }
        } else {
#pragma unroll
            for (int i = 0; i < R / 2; ++i) {
                const int run = i * ROWS_PER_IT + prow;
                const float4 v = st[run * RED + pq];
                const long long o = m_base + run * R + 2 * pq;
                if (o < a.n_out) y[run * R + 2 * pq] = make_float2(v.x, v.y);
                if (o + 1 < a.n_out) y[run * R + 2 * pq + 1] = make_float2(v.z, v.w);
            }
        }
        if (NS == 1) {
            __syncwarp();  // the stage is free again
            if (t + 1 < TPW && m_base + STEP < a.n_out) issue(m_base + STEP, stage0);
        }
    }
}
