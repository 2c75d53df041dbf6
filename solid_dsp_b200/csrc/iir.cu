// IIRFilter (SecondOrder cascade + Normal), DecimatingIIRFilter, InterpolatingIIRFilter on sm_100a.
//
// Reference loops replaced (relative to the reference's src/):
//   IIRFilter::execute_block            filter/iir/mod.rs:310-316 -> :270-289
//   SecondOrderFilter::execute          filter/iir/sos.rs:92-114
//   DecimatingIIRFilter::execute_block  filter/iir/decim.rs:222-233
//   InterpolatingIIRFilter::execute_block filter/iir/interp.rs:215-221 -> :184-190
//
// Two strategies for the SOS cascade:
//   batch : one (virtual) channel per thread.  A warp owns 32 channels and streams them in tiles
//           of 16 samples: cp.async (LDGSTS) global -> shared, double buffered, 128-byte coalesced
//           per channel; lanes then walk their own row from registers/shared and write the
//           outputs in place; the tile goes back to global coalesced.
//   scan  : one long stream is cut into P chunks that become "virtual channels" of the batch
//           kernel.  Pass A runs every chunk from zero state and keeps only the end state z_p;
//           the carry kernel solves s_{p+1} = A^Lc s_p + z_p (A^Lc built on the host in f64,
//           recurrence evaluated in f64 on the device, hierarchically); pass C re-runs every
//           chunk from its true start state and writes the outputs.
#include <algorithm>
#include <cmath>

#include "sgpu_common.cuh"

using namespace sgpu;

namespace {

constexpr int kMaxSec = 16;
// Shared-memory tile: 32 rows (one per lane) of TILE samples = TILE/2 float4 + 1 float4 of padding.
// With the odd pitch the lanes' LDS.128/STS.128 at a fixed column and the loader's rows of TILE/2
// consecutive float4 are both bank-conflict free, and every access is base + immediate.

struct SosCoefs {  // normalised by a0 (sos.rs:62-68); na* are the negated feedback taps
    float b0[kMaxSec], b1[kMaxSec], b2[kMaxSec], na1[kMaxSec], na2[kMaxSec];
};

struct IirArgs {
    const float2 *in;
    float2 *out;
    const float2 *state_in;  // [VC][NSEC][2] (v1, v2) start states, nullptr = zero
    float2 *state_out;       // [VC][NSEC][2] end states, nullptr = discard
    long long in_stride, out_stride;
    long long n_in;   // samples per real channel (WRAP 2: outputs per channel = n_in * factor)
    long long Lc;     // chunk length; P == 1 -> whole stream
    int C, P;         // real channels, chunks per channel (multiple of 32 when > 1)
    int write_out;
    int factor, idx0;  // decimation / interpolation factor, decimator counter on entry
    int vec;           // 16-byte aligned pointers and even strides -> cp.async path
    // Row layouts (a warp owns 32 rows with affine addressing; lane = row):
    //   0 plain     : row = channel.
    //   1 three-pass: per channel P slots; chunk p (Lc samples) in slot p < NP, the ragged tail chunk
    //                 alone in the warp that starts at slot `tail_slot`.  Start states come from
    //                 state_in[slot], end states go to state_out[slot].
    //   2 fused scan, chunks of one channel side by side (C < 32): warp 0 = chunk 0 of every channel,
    //                 then per channel Pm/32 warps with chunks 1 .. NPt-2, last warp = chunk NPt-1 of
    //                 every channel.
    //   3 fused scan, channels side by side (C >= 32): warp (p, cb) = chunk p of channels 32cb .. 32cb+31.
    //   Fused: chunk 0 starts from the channel's state; chunk p >= 1 starts from zero state `warm`
    //   samples early and discards those outputs (the filter's memory has decayed below 1e-10 by
    //   then); the last chunk (NPt-1, `last_len` samples) leaves the channel's new state in state_out[channel].
    int layout;
    long long n_warps;  // warps that have rows (the grid is rounded up to whole blocks)
    int NP, tail_slot;
    long long tail_off, tail_len;
    int NPt, Pm, CW;
    long long last_len;
    long long warm;    // multiple of the tile length
    SosCoefs k;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// one biquad on one complex sample, direct form II, real coefficients.  UNIT_B0: the section's b0
// has been folded into the cascade gain (b0 == 1), which saves the multiply: 4 FMA-pipe operations
// per real component instead of 5.
template <bool UNIT_B0>
__device__ __forceinline__ float2 biquad(float2 x, float2 &v1, float2 &v2, const float b0, const float b1,
                                         const float b2, const float na1, const float na2) {
    // packed: one fma.rn.f32x2 (SASS FFMA2, coefficient as the scalar-broadcast operand) per complex FMA
    float2 v0 = __ffma2_rn(v1, make_float2(na1, na1), x), y;
    v0 = __ffma2_rn(v2, make_float2(na2, na2), v0);
    if constexpr (UNIT_B0) {
        y = __ffma2_rn(v1, make_float2(b1, b1), v0);
        y = __ffma2_rn(v2, make_float2(b2, b2), y);
    } else {
        y = __fmul2_rn(v2, make_float2(b2, b2));
        y = __ffma2_rn(v1, make_float2(b1, b1), y);
        y = __ffma2_rn(v0, make_float2(b0, b0), y);
    }
    v2 = v1;
    v1 = v0;
    return y;
}

// the whole cascade on one sample.  FOLD: sections 0 .. NSEC-2 have b0 == 1 and the last section
// carries the product of all b0 (see fold_sections on the host side).  st[2s] = v1, st[2s+1] = v2.
template <int NSEC, bool FOLD>
__device__ __forceinline__ float2 cascade(float2 y, float2 (&st)[2 * NSEC], const SosCoefs &k) {
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        if (FOLD && s < NSEC - 1)
            y = biquad<true>(y, st[2 * s], st[2 * s + 1], k.b0[s], k.b1[s], k.b2[s], k.na1[s], k.na2[s]);
        else
            y = biquad<false>(y, st[2 * s], st[2 * s + 1], k.b0[s], k.b1[s], k.b2[s], k.na1[s], k.na2[s]);
    }
    return y;
}

// Normal mode (iir/mod.rs:98-130,272-280): one direct form II of order W-1 with W state values
// st[i] = v[n-1-i].  k.b0[i] = b_i / a0, k.na1[i] = -a_{i+1} / a0 (zero padded to W).
template <int W>
__device__ __forceinline__ float2 normal_step(float2 x, float2 (&st)[W], const SosCoefs &k) {
    float2 v0 = x;
#pragma unroll
    for (int i = 0; i + 1 < W; ++i) v0 = __ffma2_rn(st[i], make_float2(k.na1[i], k.na1[i]), v0);
    float2 y = __fmul2_rn(v0, make_float2(k.b0[0], k.b0[0]));
#pragma unroll
    for (int i = 0; i + 1 < W; ++i) y = __ffma2_rn(st[i], make_float2(k.b0[i + 1], k.b0[i + 1]), y);
#pragma unroll
    for (int i = W - 1; i > 0; --i) st[i] = st[i - 1];
    st[0] = v0;
    return y;
}

// one sample through the filter the kernel was instantiated for (NORD > 0: Normal mode of window NORD)
template <int NSEC, bool FOLD, int NORD>
__device__ __forceinline__ float2 filter_step(float2 x, float2 (&st)[NORD > 0 ? NORD : 2 * NSEC], const SosCoefs &k) {
    if constexpr (NORD > 0) return normal_step<NORD>(x, st, k);
    else return cascade<NSEC, FOLD>(x, st, k);
}

// WRAP: 0 plain, 1 decimating (keep every M-th output), 2 interpolating (L-1 zeros after each input)
//
// Full tiles (every row of the warp has 16 more samples) take the fast path: cp.async prefetch of
// the next tile, unpredicated fully unrolled cascade, 16-byte coalesced stores.  The ragged tail
// (and every tile of the interpolating wrapper) takes the guarded path.
template <int NSEC, int WRAP, bool FOLD, int TILE, int NW, int MINB, int NORD = 0>
__global__ void __launch_bounds__(NW * 32, MINB) iir_sos_kernel(const IirArgs a) {
    constexpr int NST = NORD > 0 ? NORD : 2 * NSEC;  // complex state values per row
    extern __shared__ float4 smem[];
    constexpr int kTile = TILE;          // samples per channel per tile (TILE * 8 bytes)
    constexpr int LPR = TILE / 2;        // loader lanes per row (16 bytes each)
    constexpr int RPI = 32 / LPR;        // rows per loader instruction
    static_assert(LPR * RPI == 32, "TILE must be 16, 32 or 64");
    constexpr int kRowF4 = LPR + 1;
    constexpr int kTileF4 = 32 * kRowF4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wi = (long long)blockIdx.x * (blockDim.x >> 5) + warp;  // global warp index
    const long long vc0 = wi * 32;
    float4 *buf = smem + warp * (2 * kTileF4);

    // affine row addressing for this warp; len0 = length of a valid row, nrows = valid rows (a prefix)
    long long in_base, out_base, rstr_in, rstr_out, warm = 0, len0 = 0;
    long long st_in = -1, st_out = -1;  // this lane's row in state_in / state_out, -1 = zero / discard
    int nrows = 0;
    const long long n_loop = WRAP == 2 ? a.n_in * a.factor : a.n_in;
    bool is_tail = false;  // layout 1
    long long p0 = 0;      // layout 1
    if (a.layout == 0) {
        if (vc0 >= a.C) return;
        in_base = vc0 * a.in_stride;
        out_base = vc0 * a.out_stride;
        rstr_in = a.in_stride;
        rstr_out = a.out_stride;
        nrows = (int)min(32LL, (long long)a.C - vc0);
        len0 = n_loop;
        st_in = st_out = vc0 + lane;
    } else if (a.layout == 1) {
        if (vc0 >= (long long)a.C * a.P) return;
        const long long chan = vc0 / a.P;
        p0 = vc0 - chan * a.P;
        is_tail = p0 == a.tail_slot;
        const long long off = is_tail ? a.tail_off : p0 * a.Lc;
        in_base = chan * a.in_stride + off;
        out_base = chan * a.out_stride + off;
        rstr_in = rstr_out = a.Lc;
        nrows = is_tail ? 1 : (int)max(0LL, min(32LL, (long long)a.NP - p0));
        len0 = is_tail ? a.tail_len : a.Lc;
        st_in = st_out = vc0 + lane;
    } else {
        // fused scan: (first chunk index, first channel, row stride) of the warp
        long long chunk, chan;
        bool by_channel;  // rows = consecutive channels (true) or consecutive chunks (false)
        if (a.layout == 3) {
            if (wi >= (long long)a.NPt * a.CW) return;
            chunk = wi / a.CW;
            chan = (wi - chunk * a.CW) * 32;
            by_channel = true;
        } else {
            const long long mid = (long long)a.C * (a.Pm / 32);
            if (wi > mid + 1) return;
            by_channel = wi == 0 || wi == mid + 1;
            if (wi == 0) { chunk = 0; chan = 0; }
            else if (wi == mid + 1) { chunk = a.NPt - 1; chan = 0; }
            else {
                chan = (wi - 1) / (a.Pm / 32);
                chunk = 1 + ((wi - 1) - chan * (a.Pm / 32)) * 32;
            }
        }
        warm = chunk > 0 ? a.warm : 0;
        const long long off = chunk * a.Lc - warm;
        in_base = chan * a.in_stride + off;
        out_base = chan * a.out_stride + off;  // the first `warm` outputs of a row are never stored
        if (by_channel) {
            rstr_in = a.in_stride;
            rstr_out = a.out_stride;
            nrows = (int)min(32LL, (long long)a.C - chan);
            len0 = (chunk == a.NPt - 1 ? a.last_len : a.Lc) + warm;
            if (chunk == 0) st_in = chan + lane;
            if (chunk == a.NPt - 1) st_out = chan + lane;
        } else {
            rstr_in = rstr_out = a.Lc;
            nrows = (int)max(0LL, min(32LL, (long long)(a.NPt - 1) - chunk));  // chunks 1 .. NPt-2
            len0 = a.Lc + warm;
        }
    }
    if (nrows <= 0) return;
    auto row_len = [&](int r) -> long long { return r < nrows ? len0 : 0; };
    const long long my_len = row_len(lane);
    const int nvalid = nrows;
    const long long max_len = len0, min_len = len0;
    const long long ntiles = (max_len + kTile - 1) / kTile;
    long long full_tiles = WRAP == 2 ? 0 : min_len / kTile;  // WRAP 2: set by the interpolating fast path
    const long long warm_tiles = warm / kTile;  // <= full_tiles: every valid row is longer than `warm`

    // state
    float2 st[NST];
#pragma unroll
    for (int i = 0; i < NST; ++i) st[i] = make_float2(0.f, 0.f);
    if (a.state_in && lane < nvalid && st_in >= 0) {
        const float2 *sp = a.state_in + st_in * NST;
        if constexpr (NST % 2 == 0) {
#pragma unroll
            for (int i = 0; i < NST / 2; ++i) {
                const float4 t = reinterpret_cast<const float4 *>(sp)[i];
                st[2 * i] = make_float2(t.x, t.y);
                st[2 * i + 1] = make_float2(t.z, t.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NST; ++i) st[i] = sp[i];
        }
    }

    // loader role: a tile is 32 * LPR 16-byte chunks; the lane moves chunks i*32 + lane, i < LPR
    const int lrow = lane / LPR, lj = lane % LPR;  // rows lrow + RPI * i, chunk lj
    auto role_r = [&](int i) { return lrow + RPI * i; };
    auto role_c = [&](int) { return lj; };
    int dec_cnt = a.idx0;                          // decimator phase (WRAP 1)
    int kout = 0;                                  // outputs kept so far in the current tile (WRAP 1)
    long long dec_pos = 0;                         // outputs stored so far per row (WRAP 1)

    // loader / storer role: the lane's i-th 16-byte chunk of a tile is chunk role_c(i) of row
    // role_r(i).  `whole` (warp-uniform): all 32 rows valid and 16-byte accesses allowed -> no
    // per-row predicates or clamps.
    const bool whole = a.vec && nvalid == 32;
    float2 *st_ptr = a.out + out_base;
    const long long st_step = RPI * rstr_out, st_role = (long long)lrow * rstr_out + 2 * lj;
    const int role_off = lrow * kRowF4 + lj;
    auto store_tile = [&](const float4 *cur) {  // one tile of outputs, 16-byte coalesced
        if (whole) {
            float2 *dst = st_ptr + st_role;
            const float4 *src = cur + role_off;
#pragma unroll
            for (int i = 0; i < LPR; ++i) {
                *reinterpret_cast<float4 *>(dst) = src[i * RPI * kRowF4];
                dst += st_step;
            }
        } else {
#pragma unroll
            for (int i = 0; i < LPR; ++i) {
                const int r = role_r(i), c = role_c(i);
                const float4 v = cur[r * kRowF4 + c];
                float2 *dst = st_ptr + (long long)r * rstr_out + 2 * c;
                if (r < nvalid) {
                    if (a.vec) {
                        *reinterpret_cast<float4 *>(dst) = v;
                    } else {
                        dst[0] = make_float2(v.x, v.y);
                        dst[1] = make_float2(v.z, v.w);
                    }
                }
            }
        }
    };

    // ------------------------------------------------------------------ interpolating wrapper, fast path
    // L a power of two <= the tile: a tile of 32 outputs per row consumes 32/L inputs per row, which
    // are staged (8-byte cp.async, double buffered in the second tile buffer) while the outputs are
    // produced in the first one: input, then L-1 zeros (iir/interp.rs:184-190).
    long long ip_done = 0;  // inputs per row consumed by this path
    int ip_phase = 0;       // outputs since the last input when the guarded path takes over (0: the next output carries one)
    if constexpr (WRAP == 2) {
        const int L = a.factor;
        const bool pow2 = (L & (L - 1)) == 0 && L >= 2 && L <= kTile;
        const long long tiles2 = pow2 ? len0 / kTile : 0;
        if (tiles2 > 0) {
            static_assert(kTile == 32, "interpolating fast path assumes 32-sample tiles");
            const int sh = __ffs(L) - 1, ish = 5 - sh, IN_T = kTile >> sh;
            constexpr int IPITCH = kTile / 2 + 1;  // float2 units: 136-byte rows, conflict-free lane reads
            constexpr int ISTAGE = 32 * IPITCH;    // two stages = one tile buffer
            float2 *ibuf = reinterpret_cast<float2 *>(buf + kTileF4);
            const float2 *ld = a.in + in_base;
            auto prefetch_in = [&](float2 *dst) {
                for (int i = 0; i < IN_T; ++i) {
                    const int e = i * 32 + lane, r = e >> ish, c = e & (IN_T - 1);
                    const int rl = r < nvalid ? r : nvalid - 1;
                    cp_async8(dst + r * IPITCH + c, ld + (long long)rl * rstr_in + c);
                }
                cp_async_commit();
                ld += IN_T;
            };
            prefetch_in(ibuf);
            for (long long t = 0; t < tiles2; ++t) {
                const float2 *irow = ibuf + (t & 1) * ISTAGE + lane * IPITCH;
                if (t + 1 < tiles2) {
                    prefetch_in(ibuf + ((t + 1) & 1) * ISTAGE);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                float4 *orow = buf + lane * kRowF4;
#pragma unroll
                for (int j = 0; j < LPR; ++j) {
                    float2 y0 = make_float2(0.f, 0.f), y1 = y0;
                    if (((2 * j) & (L - 1)) == 0) y0 = irow[(2 * j) >> sh];
                    y0 = filter_step<NSEC, FOLD, NORD>(y0, st, a.k);
                    if (((2 * j + 1) & (L - 1)) == 0) y1 = irow[(2 * j + 1) >> sh];
                    y1 = filter_step<NSEC, FOLD, NORD>(y1, st, a.k);
                    orow[j] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
                __syncwarp();
                if (a.write_out) store_tile(buf);
                st_ptr += kTile;
                __syncwarp();
            }
            ip_done = tiles2 * IN_T;
            full_tiles = tiles2;
        }
        // Any other L <= the tile (iir/interp.rs:184-190 takes every factor): the same structure with the input
        // positions tracked instead of shifted.  Output o of a row is input o / L when L | o, else a zero; a tile of 32
        // outputs consumes the inputs ceil(32 t / L) .. ceil(32 (t + 1) / L) - 1 (IN_MAX = ceil(32 / L) at most), staged
        // one tile ahead with 8-byte cp.async; every row of the warp is in the same phase.
        const long long tiles3 = (!pow2 && L >= 3 && L <= kTile) ? len0 / kTile : 0;
        if (tiles3 > 0) {
            const int IN_MAX = (kTile + L - 1) / L;
            constexpr int IPITCH = kTile / 2 + 1;
            constexpr int ISTAGE = 32 * IPITCH;
            float2 *ibuf = reinterpret_cast<float2 *>(buf + kTileF4);
            const int dr = 32 / IN_MAX, dc = 32 - dr * IN_MAX;  // (row, column) step of element e -> e + 32
            const int r0 = lane / IN_MAX, c0 = lane - r0 * IN_MAX;
            long long in_next = 0;  // first input of the next tile to stage
            auto stage_in = [&](float2 *dst, const long long t) {
                // inputs of tile t: [in_next, in_hi)
                const long long in_hi = ((t + 1) * kTile + L - 1) / L;
                const int nt = (int)(in_hi - in_next);
                const float2 *ld = a.in + in_base + in_next;
                int r = r0, c = c0;
                for (int i = 0; i < IN_MAX; ++i) {
                    if (c < nt && r < 32) {
                        const int rl = r < nvalid ? r : nvalid - 1;
                        cp_async8(dst + r * IPITCH + c, ld + (long long)rl * rstr_in + c);
                    }
                    r += dr;
                    c += dc;
                    if (c >= IN_MAX) { c -= IN_MAX; ++r; }
                }
                cp_async_commit();
                in_next = in_hi;
            };
            stage_in(ibuf, 0);
            int nextpos = 0;  // position inside the current tile of the next output that carries an input
            for (long long t = 0; t < tiles3; ++t) {
                const float2 *irow = ibuf + (t & 1) * ISTAGE + lane * IPITCH;
                if (t + 1 < tiles3) {
                    stage_in(ibuf + ((t + 1) & 1) * ISTAGE, t + 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                float4 *orow = buf + lane * kRowF4;
                int idx = 0;
#pragma unroll
                for (int j = 0; j < LPR; ++j) {
                    float2 y0 = make_float2(0.f, 0.f), y1 = y0;
                    if (2 * j == nextpos) { y0 = irow[idx++]; nextpos += L; }
                    y0 = filter_step<NSEC, FOLD, NORD>(y0, st, a.k);
                    if (2 * j + 1 == nextpos) { y1 = irow[idx++]; nextpos += L; }
                    y1 = filter_step<NSEC, FOLD, NORD>(y1, st, a.k);
                    orow[j] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
                nextpos -= kTile;
                __syncwarp();
                if (a.write_out) store_tile(buf);
                st_ptr += kTile;
                __syncwarp();
            }
            ip_done = (tiles3 * kTile + L - 1) / L;
            ip_phase = (int)((tiles3 * kTile) % L);
            full_tiles = tiles3;
        }
    }

    // ------------------------------------------------------------------ fast path: full tiles
    if (WRAP != 2 && full_tiles > 0) {
        const float2 *ld_ptr = a.in + in_base;
        const long long ld_step = RPI * rstr_in;
        const long long ld_role = (long long)lrow * rstr_in + 2 * lj;
        auto prefetch = [&](float4 *dst) {
            if (whole) {
                const float2 *src = ld_ptr + ld_role;
                float4 *d = dst + role_off;
#pragma unroll
                for (int i = 0; i < LPR; ++i) {
                    cp_async16(d + i * RPI * kRowF4, src);
                    src += ld_step;
                }
            } else {
                // rows past the valid prefix re-read the last valid row (never out of bounds, never stored)
#pragma unroll
                for (int i = 0; i < LPR; ++i) {
                    const int r = role_r(i), c = role_c(i);
                    const int rl = r < nvalid ? r : nvalid - 1;
                    const float2 *src = ld_ptr + (long long)rl * rstr_in + 2 * c;
                    float4 *d = dst + r * kRowF4 + c;
                    if (a.vec) {
                        cp_async16(d, src);
                    } else {
                        cp_async8(d, src);
                        cp_async8(reinterpret_cast<float2 *>(d) + 1, src + 1);
                    }
                }
            }
            cp_async_commit();
            ld_ptr += kTile;
        };
        prefetch(buf);
        for (long long t = 0; t < full_tiles; ++t) {
            float4 *cur = buf + (t & 1) * kTileF4;
            if (t + 1 < full_tiles) {
                prefetch(buf + ((t + 1) & 1) * kTileF4);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            float4 *myrow = cur + lane * kRowF4;
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const float4 xin = myrow[j];
                float2 y0 = make_float2(xin.x, xin.y), y1 = make_float2(xin.z, xin.w);
                y0 = filter_step<NSEC, FOLD, NORD>(y0, st, a.k);
                if constexpr (WRAP == 1) {  // kept outputs are compacted at the head of the row (warp-uniform)
                    if (++dec_cnt == a.factor) { dec_cnt = 0; reinterpret_cast<float2 *>(myrow)[kout++] = y0; }
                }
                y1 = filter_step<NSEC, FOLD, NORD>(y1, st, a.k);
                if constexpr (WRAP == 1) {
                    if (++dec_cnt == a.factor) { dec_cnt = 0; reinterpret_cast<float2 *>(myrow)[kout++] = y1; }
                }
                if constexpr (WRAP == 0) myrow[j] = make_float4(y0.x, y0.y, y1.x, y1.y);
            }
            if constexpr (WRAP == 1) {
                // the tile kept kout <= kTile/2 outputs per row (factor >= 2): half-warps store rows of 8-byte samples
                __syncwarp();
                if (a.write_out) {
                    const int hrow = lane >> 4, col = lane & 15;
                    static_assert(kTile == 32, "decimating store assumes 32-sample tiles");
                    const float2 *src = reinterpret_cast<const float2 *>(cur) + hrow * (2 * kRowF4) + col;
                    float2 *dst = a.out + out_base + (long long)hrow * rstr_out + dec_pos + col;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (col < kout && 2 * i + hrow < nvalid) *dst = src[i * (4 * kRowF4)];
                        dst += 2 * rstr_out;
                    }
                }
                dec_pos += kout;
                kout = 0;
            }
            if constexpr (WRAP == 0) {
                __syncwarp();
                if (a.write_out && t >= warm_tiles) store_tile(cur);
                st_ptr += kTile;
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ guarded path: ragged tail
    float2 *dec_ptr = a.out + out_base + (long long)lane * rstr_out + dec_pos;  // next decimated output (WRAP 1)
    if (full_tiles < ntiles) {
        float4 *cur = buf;
        const float2 *my_in = a.in + in_base + (long long)lane * rstr_in;
        int ip_cnt = ip_phase;  // interpolator phase (WRAP 2)
        long long ip_in = ip_done;  // next input index (WRAP 2)
        for (long long t = full_tiles; t < ntiles; ++t) {
            const long long s0 = t * kTile;
            if constexpr (WRAP != 2) {
                for (int i = 0; i < LPR; ++i) {
                    const int r = role_r(i), lj = role_c(i);
                    const long long rl = row_len(r);
                    const long long sidx = s0 + 2 * lj;
                    const float2 *src = a.in + in_base + (long long)r * rstr_in + sidx;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (sidx < rl) { const float2 q = src[0]; v.x = q.x; v.y = q.y; }
                    if (sidx + 1 < rl) { const float2 q = src[1]; v.z = q.x; v.w = q.y; }
                    cur[r * kRowF4 + lj] = v;
                }
                __syncwarp();
            }
            float2 *myrow2 = reinterpret_cast<float2 *>(cur + lane * kRowF4);
            for (int n = 0; n < kTile; ++n) {
                if (s0 + n >= my_len) break;
                float2 *cell = myrow2 + n;
                float2 y;
                if constexpr (WRAP == 2) {
                    y = make_float2(0.f, 0.f);
                    if (ip_cnt == 0) y = my_in[ip_in++];
                    if (++ip_cnt == a.factor) ip_cnt = 0;
                } else {
                    y = *cell;
                }
                y = filter_step<NSEC, FOLD, NORD>(y, st, a.k);
                if constexpr (WRAP == 1) {
                    if (++dec_cnt == a.factor) { dec_cnt = 0; if (a.write_out) *dec_ptr = y; ++dec_ptr; }
                } else {
                    *cell = y;
                }
            }
            if constexpr (WRAP != 1) {
                __syncwarp();
                if (a.write_out) {
                    for (int i = 0; i < LPR; ++i) {
                        const int r = role_r(i), lj = role_c(i);
                        const long long rl = row_len(r);
                        const long long sidx = s0 + 2 * lj;
                        const float4 v = cur[r * kRowF4 + lj];
                        float2 *dst = a.out + out_base + (long long)r * rstr_out + sidx;
                        if (sidx < rl) dst[0] = make_float2(v.x, v.y);
                        if (sidx + 1 < rl) dst[1] = make_float2(v.z, v.w);
                    }
                }
            }
            __syncwarp();
        }
    }

    if (a.state_out && lane < nvalid && st_out >= 0) {
        float2 *sp = a.state_out + st_out * NST;
        if constexpr (NST % 2 == 0) {
#pragma unroll
            for (int i = 0; i < NST / 2; ++i)
                reinterpret_cast<float4 *>(sp)[i] = make_float4(st[2 * i].x, st[2 * i].y, st[2 * i + 1].x, st[2 * i + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < NST; ++i) sp[i] = st[i];
        }
    }
}

// ---- carry recurrence s_{p+1} = Mat * s_p + z_p in f64, one warp per (channel, group) -------
// mode 0: zero start over the group's chunks           -> gagg[c*G+g]
// mode 1: one warp per channel over groups (Mat = A^(Lc*CH)): Sg[c*G+g] = start state of group g
// mode 2: from Sg[c*G+g] (or state0[c] when Sg == nullptr): sbuf[c*P+p] = start state of chunk p
// Lane i owns row i of Mat (registers); the state vector lives in shared memory.  The additive term
// of step p+1 is fetched from global memory while step p is being evaluated, and the dot product runs
// as four independent DFMA chains, so a step costs ~one shared-memory round trip.
__global__ void __launch_bounds__(128) carry_kernel(int mode, const double *__restrict__ Mat, int D,
                                                    const float2 *__restrict__ z, double2 *gagg, double2 *Sg,
                                                    float2 *__restrict__ sbuf, const float2 *__restrict__ state0,
                                                    int P, int NP, int tail_slot, int CH, int G, int n_units) {
    __shared__ double2 sh[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int unit = blockIdx.x * 4 + w;
    if (unit >= n_units) return;
    double m[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) m[j] = (lane < D && j < D) ? Mat[lane * D + j] : 0.0;
    double2 *s = sh[w];
    const int DQ = (D + 3) & ~3;
    auto step = [&](double2 add) {
        double2 a0 = add, a1 = make_double2(0., 0.), a2 = a1, a3 = a1;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (j < DQ) {
                const double2 s0 = s[j], s1 = s[j + 1], s2 = s[j + 2], s3 = s[j + 3];
                a0.x = fma(m[j], s0.x, a0.x);         a0.y = fma(m[j], s0.y, a0.y);
                a1.x = fma(m[j + 1], s1.x, a1.x);     a1.y = fma(m[j + 1], s1.y, a1.y);
                a2.x = fma(m[j + 2], s2.x, a2.x);     a2.y = fma(m[j + 2], s2.y, a2.y);
                a3.x = fma(m[j + 3], s3.x, a3.x);     a3.y = fma(m[j + 3], s3.y, a3.y);
            }
        }
        __syncwarp();
        s[lane] = make_double2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
        __syncwarp();
    };
    if (mode == 1) {
        const int c = unit;
        const float2 i0 = lane < D ? state0[(long long)c * D + lane] : make_float2(0.f, 0.f);
        s[lane] = make_double2(i0.x, i0.y);
        __syncwarp();
        const double2 *ga = gagg + (long long)c * G * D + lane;
        double2 nxt = (lane < D && G > 0) ? ga[0] : make_double2(0., 0.);
        for (int g = 0; g < G; ++g) {
            const double2 add = nxt;
            if (lane < D && g + 1 < G) nxt = ga[(long long)(g + 1) * D];
            if (lane < D) Sg[((long long)c * G + g) * D + lane] = s[lane];
            step(add);
        }
        return;
    }
    const int c = unit / G, g = unit - c * G;
    const int p_lo = g * CH, p_hi = min(NP, p_lo + CH);  // chunk slots [0, NP) are scanned; P is the slot stride
    if (mode == 0) {
        s[lane] = make_double2(0., 0.);
    } else {
        double2 st = make_double2(0., 0.);
        if (lane < D) {
            if (Sg) st = Sg[((long long)c * G + g) * D + lane];
            else {
                const float2 i0 = state0[(long long)c * D + lane];
                st = make_double2(i0.x, i0.y);
            }
        }
        s[lane] = st;
    }
    __syncwarp();
    // software prefetch of z, four steps ahead
    constexpr int PF = 4;
    float2 zq[PF];
    const float2 *zp = z + ((long long)c * P + p_lo) * D + lane;
#pragma unroll
    for (int k = 0; k < PF; ++k) zq[k] = (lane < D && p_lo + k < p_hi) ? zp[(long long)k * D] : make_float2(0.f, 0.f);
    for (int p = p_lo; p < p_hi; p += PF) {
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            if (p + k < p_hi) {
                const long long vc = (long long)c * P + p + k;
                if (mode == 2 && lane < D) sbuf[vc * D + lane] = make_float2((float)s[lane].x, (float)s[lane].y);
                const float2 zz = zq[k];
                zq[k] = (lane < D && p + k + PF < p_hi) ? zp[(long long)(p - p_lo + k + PF) * D] : make_float2(0.f, 0.f);
                step(make_double2(zz.x, zz.y));
            }
        }
    }
    if (mode == 0 && lane < D) gagg[((long long)c * G + g) * D + lane] = s[lane];
    // the state after the last full chunk starts the ragged tail chunk
    if (mode == 2 && p_hi == NP && tail_slot >= 0 && lane < D)
        sbuf[((long long)c * P + tail_slot) * D + lane] = make_float2((float)s[lane].x, (float)s[lane].y);
}

// ---- Normal mode: one direct-form II of arbitrary order (iir/mod.rs:98-130,272-280) ---------
constexpr int kMaxOrder = 64;
struct NormalArgs {
    const float2 *in;
    float2 *out;
    float2 *state;  // [C][W] newest first
    long long in_stride, out_stride, n_in;
    int C, W, nb, na1;  // na1 = len(a) - 1 feedback taps
    int wrap, factor, idx0;
    float b[kMaxOrder], a[kMaxOrder];  // b[0..nb), a[i] = fb[i+1]/a0
};

__global__ void __launch_bounds__(128) iir_normal_kernel(const NormalArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    float2 w[kMaxOrder];
    const int W = a.W;
    for (int i = 0; i < W; ++i) w[i] = a.state[(long long)c * W + i];
    int head = 0;  // w[(head + i) % W] = i-th newest
    const float2 *x = a.in + (long long)c * a.in_stride;
    float2 *y = a.out + (long long)c * a.out_stride;
    const int nden = min(W - 1, a.na1);  // iir/mod.rs:274 + dot_product/mod.rs:160
    const int nnum = min(W, a.nb);
    long long cnt = a.idx0, o = 0;
    const long long n_loop = a.wrap == 2 ? a.n_in * a.factor : a.n_in;
    int ip = 0;
    long long in_i = 0;
    for (long long n = 0; n < n_loop; ++n) {
        float2 xin;
        if (a.wrap == 2) {
            xin = ip == 0 ? x[in_i++] : make_float2(0.f, 0.f);
            if (++ip == a.factor) ip = 0;
        } else {
            xin = x[n];
        }
        float2 den = make_float2(0.f, 0.f);
        for (int i = 0; i < nden; ++i) {
            const float2 v = w[(head + i) % W];
            den.x = fmaf(a.a[i], v.x, den.x);
            den.y = fmaf(a.a[i], v.y, den.y);
        }
        const float2 v0 = make_float2(xin.x - den.x, xin.y - den.y);
        head = (head + W - 1) % W;
        w[head] = v0;
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < nnum; ++i) {
            const float2 v = w[(head + i) % W];
            acc.x = fmaf(a.b[i], v.x, acc.x);
            acc.y = fmaf(a.b[i], v.y, acc.y);
        }
        if (a.wrap == 1) {
            if (++cnt == a.factor) {
                cnt = 0;
                y[o++] = acc;
            }
        } else {
            y[n] = acc;
        }
    }
    for (int i = 0; i < W; ++i) a.state[(long long)c * W + i] = w[(head + i) % W];
}

int pad_sections(int nsec) {
    const int opts[] = {1, 2, 4, 8, 16};
    for (int o : opts)
        if (nsec <= o) return o;
    return -1;
}

}  // namespace

struct sgpu_iir {
    int device = 0, sm_count = 0;
    int type = SGPU_IIR_SECOND_ORDER;
    int wrap = SGPU_IIR_PLAIN;
    size_t factor = 1, C = 0;
    int nsec = 0, nsec_pad = 0;       // SOS
    int W = 0, nb = 0, na = 0;        // Normal
    uint64_t index = 0;               // decimator counter (iir/decim.rs:9)
    int mode = -1;
    bool fold = false;           // b0 of every section folded into the last one (see fold_sections)
    std::vector<double> gpre;    // [nsec_pad] product of b0 over the sections in front of s: v_ref = v_dev * gpre[s]
    std::vector<double> ff_raw, fb_raw;  // as given (numerator_coefs()/denominator_coefs() in SOS mode)
    std::vector<double> num_norm, den_norm;  // Normal mode: ff/a0, fb[1..]/a0
    SosCoefs k{};
    std::vector<double> kd;  // the kernel's f32 coefs as doubles [nsec_pad][5] = b0,b1,b2,a1,a2
    float2 *d_state = nullptr;  // SOS: [C][nsec_pad][2]; Normal: [C][W]
    // scan scratch
    float2 *d_z = nullptr, *d_s = nullptr;
    double2 *d_gagg = nullptr, *d_Sg = nullptr;
    double *d_mat = nullptr;  // [2][D*D]: A^Lc, A^(Lc*CH)
    size_t scratch_vc = 0, scratch_units = 0;
    long long mat_Lc = -1, mat_CH = -1;
    long long decay_len = -2;  // see decay_length(); -2 = not computed yet, -1 = does not decay
    Staging stage;
    HostPipe pipe;
};

namespace {

size_t state_len_dev(const sgpu_iir *f) {
    return f->type == SGPU_IIR_SECOND_ORDER ? (size_t)f->nsec_pad * 2 : (size_t)f->W;
}

// one-sample zero-input transition matrix of the cascade, state order [s][v1,v2], in f64
void build_transition(const sgpu_iir *f, std::vector<double> &A) {
    const int D = 2 * f->nsec_pad;
    A.assign((size_t)D * D, 0.0);
    for (int col = 0; col < D; ++col) {
        std::vector<double> st(D, 0.0);
        st[col] = 1.0;
        double y = 0.0;  // zero input
        for (int s = 0; s < f->nsec_pad; ++s) {
            const double b0 = f->kd[s * 5 + 0], b1 = f->kd[s * 5 + 1], b2 = f->kd[s * 5 + 2];
            const double a1 = f->kd[s * 5 + 3], a2 = f->kd[s * 5 + 4];
            const double v1 = st[2 * s], v2 = st[2 * s + 1];
            const double v0 = y - (a1 * v1 + a2 * v2);
            y = b0 * v0 + b1 * v1 + b2 * v2;
            st[2 * s + 1] = v1;
            st[2 * s] = v0;
        }
        for (int r = 0; r < D; ++r) A[(size_t)r * D + col] = st[r];
    }
}

void mat_mul(const std::vector<double> &X, const std::vector<double> &Y, std::vector<double> &Z, int D) {
    std::vector<double> T((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
            const double x = X[(size_t)i * D + k];
            if (x == 0.0) continue;
            for (int j = 0; j < D; ++j) T[(size_t)i * D + j] += x * Y[(size_t)k * D + j];
        }
    Z.swap(T);
}

void mat_pow(std::vector<double> A, long long e, std::vector<double> &R, int D) {
    R.assign((size_t)D * D, 0.0);
    for (int i = 0; i < D; ++i) R[(size_t)i * D + i] = 1.0;
    while (e > 0) {
        if (e & 1) mat_mul(R, A, R, D);
        e >>= 1;
        if (e) mat_mul(A, A, A, D);
    }
}

// Tile shapes (measured on B200, 65536 channels x 2^14, profiles/r1f_iir_tiles.md): 32-sample tiles
// (256-byte row segments, 4 warps x 3 blocks = 12 warps/SM) reach 5.5 TB/s where 16-sample tiles
// (128-byte segments, 24 warps/SM) stop at 4.5 TB/s; 64-sample tiles leave too few warps (6/SM).
constexpr int kScanTile = 32;
constexpr int kWarpsPerSm = 12;  // resident warps per SM of the plain kernel (smem-limited)

template <int NSEC, int WRAP, bool FOLD>
int launch_sos_t(const IirArgs &a, cudaStream_t s) {
    constexpr int TILE = kScanTile, NW = 4, MINB = 3;
    const unsigned blocks = (unsigned)((a.n_warps + NW - 1) / NW);
    const size_t smem = (size_t)NW * 2 * 32 * (TILE / 2 + 1) * sizeof(float4);
    auto kern = iir_sos_kernel<NSEC, WRAP, FOLD, TILE, NW, MINB>;
    SGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, NW * 32, smem, s>>>(a);
    SGPU_LAUNCH_CHECK();
    count_launch();
    return SGPU_OK;
}

template <int NSEC, bool FOLD>
int launch_sos_n(const IirArgs &a, int wrap, cudaStream_t s) {
    if (wrap == 1) return launch_sos_t<NSEC, 1, FOLD>(a, s);
    if (wrap == 2) return launch_sos_t<NSEC, 2, FOLD>(a, s);
    return launch_sos_t<NSEC, 0, FOLD>(a, s);
}

// Normal mode, plain wrapper, window W <= 8: the tile kernel with the direct-form step
template <int W>
int launch_normal_t(const IirArgs &a, cudaStream_t s) {
    constexpr int TILE = kScanTile, NW = 4, MINB = 3;
    const unsigned blocks = (unsigned)((a.n_warps + NW - 1) / NW);
    const size_t smem = (size_t)NW * 2 * 32 * (TILE / 2 + 1) * sizeof(float4);
    auto kern = iir_sos_kernel<1, 0, false, TILE, NW, MINB, W>;
    SGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, NW * 32, smem, s>>>(a);
    SGPU_LAUNCH_CHECK();
    count_launch();
    return SGPU_OK;
}

int launch_normal(int W, const IirArgs &a, cudaStream_t s) {
    switch (W) {
        case 1: return launch_normal_t<1>(a, s);
        case 2: return launch_normal_t<2>(a, s);
        case 3: return launch_normal_t<3>(a, s);
        case 4: return launch_normal_t<4>(a, s);
        case 5: return launch_normal_t<5>(a, s);
        case 6: return launch_normal_t<6>(a, s);
        case 7: return launch_normal_t<7>(a, s);
        case 8: return launch_normal_t<8>(a, s);
    }
    return fail(SGPU_ERR_UNSUPPORTED, "Normal-mode window above 8 in the tile kernel");
}

int launch_sos(const sgpu_iir *f, const IirArgs &a, int wrap, cudaStream_t s) {
    if (wrap != 0 && a.factor == 1) wrap = 0;  // decimation / interpolation by 1 is the plain filter
#define SGPU_SOS_CASE(N) \
    case N: return f->fold ? launch_sos_n<N, true>(a, wrap, s) : launch_sos_n<N, false>(a, wrap, s);
    switch (f->nsec_pad) {
        SGPU_SOS_CASE(1)
        SGPU_SOS_CASE(2)
        SGPU_SOS_CASE(4)
        SGPU_SOS_CASE(8)
        SGPU_SOS_CASE(16)
    }
#undef SGPU_SOS_CASE
    return fail(SGPU_ERR_UNSUPPORTED, "unsupported section count");
}

// Smallest multiple of 32 with ||A^k||_inf < 1e-10 (A = one-sample zero-input transition of the
// cascade), searched up to 2^16 samples; -1 = the filter does not decay (or is unstable).  Past that
// many samples the influence of an older state is 3 orders of magnitude below f32 resolution.
long long decay_length(sgpu_iir *f) {
    if (f->decay_len != -2) return f->decay_len;
    const int D = 2 * f->nsec_pad;
    std::vector<double> A, A32, Pk;
    build_transition(f, A);
    mat_pow(A, 32, A32, D);
    Pk = A32;
    f->decay_len = -1;
    for (long long k = 32; k <= 65536; k += 32) {
        double nrm = 0.0;
        for (int i = 0; i < D; ++i) {
            double rs = 0.0;
            for (int j = 0; j < D; ++j) rs += std::fabs(Pk[(size_t)i * D + j]);
            nrm = rs > nrm ? rs : nrm;
        }
        if (!(nrm == nrm) || nrm > 1e30) break;
        if (nrm < 1e-10) { f->decay_len = k; break; }
        mat_mul(Pk, A32, Pk, D);
    }
    return f->decay_len;
}

int ensure_scan_scratch(sgpu_iir *f, size_t vc, size_t units, long long Lc, long long CH) {
    const int D = 2 * f->nsec_pad;
    if (vc > f->scratch_vc) {
        if (f->d_z) cudaFree(f->d_z);
        if (f->d_s) cudaFree(f->d_s);
        f->d_z = f->d_s = nullptr;
        f->scratch_vc = 0;
        SGPU_CUDA(cudaMalloc(&f->d_z, vc * D * sizeof(float2)));
        SGPU_CUDA(cudaMalloc(&f->d_s, vc * D * sizeof(float2)));
        f->scratch_vc = vc;
    }
    if (units > f->scratch_units && units > 0) {
        if (f->d_gagg) cudaFree(f->d_gagg);
        if (f->d_Sg) cudaFree(f->d_Sg);
        f->d_gagg = f->d_Sg = nullptr;
        f->scratch_units = 0;
        SGPU_CUDA(cudaMalloc(&f->d_gagg, units * D * sizeof(double2)));
        SGPU_CUDA(cudaMalloc(&f->d_Sg, units * D * sizeof(double2)));
        f->scratch_units = units;
    }
    if (Lc < 0) return SGPU_OK;  // fused scan: state scratch only
    if (!f->d_mat) SGPU_CUDA(cudaMalloc(&f->d_mat, 2 * (size_t)D * D * sizeof(double)));
    if (f->mat_Lc != Lc || f->mat_CH != CH) {
        std::vector<double> A, ALc, AG;
        build_transition(f, A);
        mat_pow(A, Lc, ALc, D);
        mat_pow(ALc, CH, AG, D);
        SGPU_CUDA(cudaMemcpy(f->d_mat, ALc.data(), (size_t)D * D * sizeof(double), cudaMemcpyHostToDevice));
        SGPU_CUDA(cudaMemcpy(f->d_mat + (size_t)D * D, AG.data(), (size_t)D * D * sizeof(double),
                             cudaMemcpyHostToDevice));
        f->mat_Lc = Lc;
        f->mat_CH = CH;
    }
    return SGPU_OK;
}

int iir_run(sgpu_iir *f, const float2 *d_in, long long n_in, long long istr, float2 *d_out, long long ostr,
            cudaStream_t s) {
    const bool vec = ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0) && (istr % 2 == 0) &&
                     ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0) && (ostr % 2 == 0);
    if (f->type == SGPU_IIR_NORMAL && f->wrap == SGPU_IIR_PLAIN && f->W <= 8 && !getenv("SGPU_IIR_NORMAL_SLOW")) {
        // tile kernel (coalesced 256-byte row segments, states in registers); coefficients ride in the
        // SosCoefs block: b0[i] = b_i/a0, na1[i] = -a_{i+1}/a0
        IirArgs a{};
        a.in = d_in; a.out = d_out;
        a.in_stride = istr; a.out_stride = ostr; a.n_in = n_in;
        a.C = (int)f->C; a.factor = 1; a.idx0 = 0;
        a.vec = vec ? 1 : 0;
        for (int i = 0; i < kMaxSec; ++i) { a.k.b0[i] = 0.f; a.k.na1[i] = 0.f; }
        for (int i = 0; i < f->nb; ++i) a.k.b0[i] = (float)f->num_norm[i];
        for (int i = 0; i < f->na - 1; ++i) a.k.na1[i] = -(float)f->den_norm[i];
        a.layout = 0; a.P = 1; a.Lc = n_in; a.n_warps = (long long)ceil_div(f->C, 32);
        a.state_in = f->d_state; a.state_out = f->d_state; a.write_out = 1;
        return launch_normal(f->W, a, s);
    }
    if (f->type == SGPU_IIR_NORMAL) {
        NormalArgs a{};
        a.in = d_in; a.out = d_out; a.state = f->d_state;
        a.in_stride = istr; a.out_stride = ostr; a.n_in = n_in;
        a.C = (int)f->C; a.W = f->W; a.nb = f->nb; a.na1 = f->na - 1;
        a.wrap = f->wrap; a.factor = (int)f->factor; a.idx0 = (int)f->index;
        for (int i = 0; i < f->nb; ++i) a.b[i] = (float)f->num_norm[i];
        for (int i = 0; i < f->na - 1; ++i) a.a[i] = (float)f->den_norm[i];
        iir_normal_kernel<<<(unsigned)ceil_div(f->C, 128), 128, 0, s>>>(a);
        SGPU_LAUNCH_CHECK();
        count_launch();
        return SGPU_OK;
    }
    IirArgs a{};
    a.in = d_in; a.out = d_out;
    a.in_stride = istr; a.out_stride = ostr; a.n_in = n_in;
    a.C = (int)f->C; a.factor = (int)f->factor; a.idx0 = (int)f->index;
    a.vec = vec ? 1 : 0;
    a.k = f->k;
    a.layout = 0; a.P = 1; a.Lc = n_in;
    a.NP = 0; a.tail_slot = -1; a.tail_off = 0; a.tail_len = 0;
    a.NPt = 0; a.Pm = 0; a.CW = 0; a.last_len = 0; a.warm = 0;
    auto plain = [&]() {
        a.layout = 0; a.P = 1; a.Lc = n_in; a.n_warps = (long long)ceil_div(f->C, 32);
        a.state_in = f->d_state; a.state_out = f->d_state; a.write_out = 1;
        return launch_sos(f, a, f->wrap, s);
    };
    // ---- strategy.  One lane per channel needs about sm_count * 12 * 32 channels to fill the chip;
    // with fewer, every channel's stream is cut into chunks that run side by side (scan).
    const long long lanes = (long long)f->sm_count * kWarpsPerSm * 32;
    const bool want_scan = f->wrap == SGPU_IIR_PLAIN &&
                           (f->mode >= 1 || (f->mode == -1 && (long long)f->C * 2 <= lanes && n_in >= 4096));
    if (!want_scan) return plain();
    const int D = 2 * f->nsec_pad;
    const long long decay = (f->mode == 2 || getenv("SGPU_IIR_NO_TRUNC")) ? -1 : decay_length(f);
    const long long chunks_wanted = std::max<long long>(2, lanes / (long long)f->C);
    if (decay > 0) {
        // ---- fused scan: one launch, every chunk but the first warms up over the `decay` samples in
        // front of it.  Chunks are at least 2 * decay long (re-read <= 50 % of the input) unless the
        // scan is forced (mode 1: at least `decay`).
        const long long warm = (long long)round_up((size_t)decay, kScanTile);
        const long long lc_ideal = (long long)round_up(ceil_div((size_t)n_in, (size_t)chunks_wanted), kScanTile);
        long long Lc = std::max(lc_ideal, f->mode == 1 ? warm : 2 * warm);
        long long NPt = (long long)ceil_div((size_t)n_in, (size_t)Lc);
        if (f->mode == -1 && NPt * (long long)f->C * 5 < lanes * 2 && Lc + warm > 4096) {
            // a long filter memory leaves too few chunks to fill 40 % of the lanes AND every lane has a
            // long sequential run in front of it (~0.1 us per sample when the chip is that empty): hand
            // over to the three-pass scan, whose chunks are as short as they like.  Measured, 8 sections
            // at pole radius 0.999 (memory 62912 samples), 2^28 samples: fused 21, three-pass 148 Gsamp/s.
            // Short streams stay fused: their few chunks are short (2^22 samples: 0.06 vs 0.27 ms).
            NPt = 0;
        }
        if (NPt == 1) return plain();
        if (NPt >= 2) {
        a.Lc = Lc; a.warm = warm; a.NPt = (int)NPt; a.last_len = n_in - (NPt - 1) * Lc;
        long long warps;
        if (f->C >= 32) {
            a.layout = 3;
            a.CW = (int)ceil_div(f->C, 32);
            warps = NPt * a.CW;
        } else {
            a.layout = 2;
            a.Pm = (int)round_up((size_t)(NPt - 2), 32);
            warps = 2 + (long long)f->C * (a.Pm / 32);
        }
        a.n_warps = warps;
        if (f->scratch_vc < f->C) {
            int st = ensure_scan_scratch(f, f->C, 0, -1, -1);
            if (st) return st;
        }
        a.state_in = f->d_state; a.state_out = f->d_z; a.write_out = 1;
        int st = launch_sos(f, a, 0, s);
        if (st) return st;
        SGPU_CUDA(cudaMemcpyAsync(f->d_state, f->d_z, f->C * (size_t)D * sizeof(float2), cudaMemcpyDeviceToDevice, s));
        return SGPU_OK;
        }
    }
    // ---- three-pass scan: filters whose memory does not decay within 2^16 samples, or is too long
    // for the fused scan to fill the chip.
    long long Lc = (long long)round_up(ceil_div((size_t)n_in, (size_t)chunks_wanted), kScanTile);
    if (Lc < 256) Lc = 256;
    const long long NP = n_in / Lc;          // full chunks
    const long long tail = n_in - NP * Lc;   // ragged tail chunk (its own warp)
    if (NP == 0) return plain();
    const long long Pw = (long long)round_up((size_t)NP, 32);
    const long long P = Pw + (tail ? 32 : 0);
    const int tail_slot = tail ? (int)Pw : -1;
    const long long CH = 256;
    const long long G = (NP + CH - 1) / CH;
    int st = ensure_scan_scratch(f, (size_t)f->C * P, (size_t)f->C * G, Lc, CH);
    if (st) return st;
    a.layout = 1; a.n_warps = (long long)f->C * P / 32;
    a.P = (int)P; a.Lc = Lc; a.NP = (int)NP; a.tail_slot = tail_slot; a.tail_off = NP * Lc;
    // pass A: zero-state end state of every full chunk
    a.state_in = nullptr; a.state_out = f->d_z; a.write_out = 0; a.tail_len = 0;
    st = launch_sos(f, a, 0, s);
    if (st) return st;
    if (G > 1) {
        const int units = (int)(f->C * G);
        carry_kernel<<<(unsigned)ceil_div((size_t)units, 4), 128, 0, s>>>(0, f->d_mat, D, f->d_z, f->d_gagg, nullptr, nullptr,
                                                                         nullptr, (int)P, (int)NP, tail_slot, (int)CH, (int)G, units);
        SGPU_LAUNCH_CHECK();
        carry_kernel<<<(unsigned)ceil_div(f->C, 4), 128, 0, s>>>(1, f->d_mat + (size_t)D * D, D, nullptr, f->d_gagg, f->d_Sg,
                                                                 nullptr, f->d_state, (int)P, (int)NP, tail_slot, (int)CH, (int)G,
                                                                 (int)f->C);
        SGPU_LAUNCH_CHECK();
        carry_kernel<<<(unsigned)ceil_div((size_t)units, 4), 128, 0, s>>>(2, f->d_mat, D, f->d_z, nullptr, f->d_Sg, f->d_s,
                                                                         f->d_state, (int)P, (int)NP, tail_slot, (int)CH, (int)G, units);
        SGPU_LAUNCH_CHECK();
        count_launch(3);
    } else {
        const int units = (int)f->C;
        carry_kernel<<<(unsigned)ceil_div((size_t)units, 4), 128, 0, s>>>(2, f->d_mat, D, f->d_z, nullptr, nullptr, f->d_s,
                                                                         f->d_state, (int)P, (int)NP, tail_slot, (int)CH, 1, units);
        SGPU_LAUNCH_CHECK();
        count_launch();
    }
    // pass C: true start states, outputs, end states
    a.state_in = f->d_s; a.state_out = f->d_z; a.write_out = 1; a.tail_len = tail;
    st = launch_sos(f, a, 0, s);
    if (st) return st;
    // the handle's new state = end state of the last chunk of every channel
    const long long last = tail ? (long long)tail_slot : NP - 1;
    SGPU_CUDA(cudaMemcpy2DAsync(f->d_state, (size_t)D * sizeof(float2), f->d_z + (size_t)last * D,
                                (size_t)P * D * sizeof(float2), (size_t)D * sizeof(float2), f->C,
                                cudaMemcpyDeviceToDevice, s));
    return SGPU_OK;
}

}  // namespace

namespace {

// Kernel coefficients.  Unfolded: the normalised coefficients as they are.  Folded: section s < last
// runs with (1, b1/b0, b2/b0) and the last section is multiplied by the product of the b0 in front
// of it, so the cascade output is unchanged while every section but one saves a multiply per real
// component (the cascade is linear: scaling a section's output scales everything behind it).  The
// states of section s are then kept divided by gpre[s] = prod_{k<s} b0_k; get/set_state convert.
// Folding needs every b0 != 0 and the running product within 2^+-40 (f32 range to spare).
void fold_sections(sgpu_iir *f, const double (&nc)[kMaxSec][5]) {
    const int NP = f->nsec_pad;
    f->gpre.assign((size_t)NP, 1.0);
    bool ok = NP > 1 && !getenv("SGPU_IIR_NO_FOLD");
    double g = 1.0;
    for (int s = 0; s + 1 < NP && ok; ++s) {
        const double b0 = nc[s][0];
        g *= b0;
        if (!(b0 != 0.0) || !std::isfinite(g) || std::fabs(g) > 0x1p40 || std::fabs(g) < 0x1p-40) ok = false;
        f->gpre[s + 1] = g;
    }
    if (ok) {
        const float chk[3] = {(float)(g * nc[NP - 1][0]), (float)(g * nc[NP - 1][1]), (float)(g * nc[NP - 1][2])};
        for (float c : chk)
            if (!std::isfinite(c)) ok = false;
    }
    f->fold = ok;
    if (!ok) f->gpre.assign((size_t)NP, 1.0);
    f->kd.assign((size_t)NP * 5, 0.0);
    for (int s = 0; s < kMaxSec; ++s) {
        float b0 = (float)nc[s][0], b1 = (float)nc[s][1], b2 = (float)nc[s][2];
        if (ok && s < NP) {
            if (s + 1 < NP) {
                b1 = (float)(nc[s][1] / nc[s][0]);
                b2 = (float)(nc[s][2] / nc[s][0]);
                b0 = 1.f;
            } else {
                b0 = (float)(g * nc[s][0]);
                b1 = (float)(g * nc[s][1]);
                b2 = (float)(g * nc[s][2]);
            }
        }
        f->k.b0[s] = b0; f->k.b1[s] = b1; f->k.b2[s] = b2;
        f->k.na1[s] = -(float)nc[s][3]; f->k.na2[s] = -(float)nc[s][4];
        if (s < NP) {
            f->kd[s * 5 + 0] = b0; f->kd[s * 5 + 1] = b1; f->kd[s * 5 + 2] = b2;
            f->kd[s * 5 + 3] = (float)nc[s][3]; f->kd[s * 5 + 4] = (float)nc[s][4];
        }
    }
}

}  // namespace

SGPU_EXPORT int sgpu_iir_create(sgpu_iirtype type, const double *ff, size_t n_ff, const double *fb, size_t n_fb,
                                size_t n_channels, sgpu_iirwrap wrap, size_t factor, sgpu_iir **out) {
    if (!out) return fail(SGPU_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (wrap != SGPU_IIR_PLAIN) {  // iir/decim.rs:31-41, iir/interp.rs:30-40
        if (n_ff == 0) return fail(SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO, "IIR Filter Error NumeratorLengthZero");
        if (n_fb == 0) return fail(SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO, "IIR Filter Error DenominatorLengthZero");
        if (factor < 1)
            return wrap == SGPU_IIR_DECIMATING
                       ? fail(SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE, "IIR Filter Error DecimationLessThanOne")
                       : fail(SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE, "IIR Filter Error InterpolationLessThanOne");
    }
    if (type == SGPU_IIR_NORMAL) {  // iir/mod.rs:99-103
        if (n_ff == 0) return fail(SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO, "IIR Filter Error NumeratorLengthZero");
        if (n_fb == 0) return fail(SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO, "IIR Filter Error DenominatorLengthZero");
    } else {  // iir/mod.rs:132-142
        if (n_ff != n_fb) return fail(SGPU_ERR_IIR_SOS_SIZE_MISMATCH, "IIR Filter Error SecondOrderSectionSizeMismatch");
        if (n_ff == 0) return fail(SGPU_ERR_IIR_SOS_SIZE_ZERO, "IIR Filter Error SecondOrderSectionSizeZero");
        if (n_ff % 3 != 0)
            return fail(SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3, "IIR Filter Error SecondOrderSectionSizeNotMultpleOf3");
    }
    if (!ff || !fb) return fail(SGPU_ERR_INVALID_ARGUMENT, "null coefficients");
    if (n_channels == 0) return fail(SGPU_ERR_INVALID_ARGUMENT, "n_channels == 0");
    if (type == SGPU_IIR_SECOND_ORDER && n_ff / 3 > (size_t)kMaxSec)
        return fail(SGPU_ERR_UNSUPPORTED, "more than %d second-order sections", kMaxSec);
    if (type == SGPU_IIR_NORMAL && (n_ff > (size_t)kMaxOrder || n_fb > (size_t)kMaxOrder))
        return fail(SGPU_ERR_UNSUPPORTED, "Normal-mode order above %d", kMaxOrder);
    int dev = 0, sms = 0;
    int st = require_device(&dev, &sms);
    if (st) return st;
    sgpu_iir *f = new (std::nothrow) sgpu_iir();
    if (!f) return fail(SGPU_ERR_ALLOC, "out of host memory");
    f->device = dev;
    f->sm_count = sms;
    f->type = type;
    f->wrap = wrap;
    f->factor = wrap == SGPU_IIR_PLAIN ? 1 : factor;
    f->C = n_channels;
    f->ff_raw.assign(ff, ff + n_ff);
    f->fb_raw.assign(fb, fb + n_fb);
    if (type == SGPU_IIR_SECOND_ORDER) {
        f->nsec = (int)(n_ff / 3);
        f->nsec_pad = pad_sections(f->nsec);
        double nc[kMaxSec][5];  // normalised by a0 (sos.rs:62-68), rounded to f32: b0, b1, b2, a1, a2
        for (int s = 0; s < kMaxSec; ++s) {  // identity padding: y = x
            nc[s][0] = 1.0; nc[s][1] = nc[s][2] = nc[s][3] = nc[s][4] = 0.0;
        }
        for (int s = 0; s < f->nsec; ++s) {
            const double a0 = fb[3 * s];
            for (int i = 0; i < 3; ++i) nc[s][i] = (double)(float)(ff[3 * s + i] / a0);
            nc[s][3] = (double)(float)(fb[3 * s + 1] / a0);
            nc[s][4] = (double)(float)(fb[3 * s + 2] / a0);
        }
        fold_sections(f, nc);
    } else {
        f->nb = (int)n_ff;
        f->na = (int)n_fb;
        f->W = (int)(n_fb > n_ff ? n_fb : n_ff);  // iir/mod.rs:105-109
        const double a0 = fb[0];
        for (size_t i = 0; i < n_ff; ++i) f->num_norm.push_back(ff[i] / a0);
        for (size_t i = 1; i < n_fb; ++i) f->den_norm.push_back(fb[i] / a0);
    }
    const size_t sb = n_channels * state_len_dev(f) * sizeof(float2);
    if (cudaMalloc(&f->d_state, sb) != cudaSuccess) {
        sgpu_iir_destroy(f);
        return fail(SGPU_ERR_CUDA, "cudaMalloc(iir state %zu bytes) failed", sb);
    }
    cudaMemset(f->d_state, 0, sb);
    *out = f;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_iir_destroy(sgpu_iir *f) {
    if (!f) return SGPU_OK;
    DeviceGuard g(f->device);
    if (f->d_state) cudaFree(f->d_state);
    if (f->d_z) cudaFree(f->d_z);
    if (f->d_s) cudaFree(f->d_s);
    if (f->d_gagg) cudaFree(f->d_gagg);
    if (f->d_Sg) cudaFree(f->d_Sg);
    if (f->d_mat) cudaFree(f->d_mat);
    f->stage.release();
    f->pipe.release();
    delete f;
    return SGPU_OK;
}

SGPU_EXPORT size_t sgpu_iir_out_len(const sgpu_iir *f, size_t n_in) {
    if (!f) return 0;
    if (f->wrap == SGPU_IIR_DECIMATING) return (size_t)((f->index + n_in) / f->factor);
    if (f->wrap == SGPU_IIR_INTERPOLATING) return n_in * f->factor;
    return n_in;
}
SGPU_EXPORT size_t sgpu_iir_sections(const sgpu_iir *f) { return f ? (size_t)f->nsec : 0; }
SGPU_EXPORT size_t sgpu_iir_channels(const sgpu_iir *f) { return f ? f->C : 0; }
SGPU_EXPORT int sgpu_iir_type(const sgpu_iir *f) { return f ? f->type : -1; }
SGPU_EXPORT size_t sgpu_iir_state_len(const sgpu_iir *f) {
    if (!f) return 0;
    return f->type == SGPU_IIR_SECOND_ORDER ? (size_t)f->nsec * 2 : (size_t)f->W;
}
SGPU_EXPORT int sgpu_iir_set_mode(sgpu_iir *f, int mode) {
    if (!f || mode < -1 || mode > 2) return fail(SGPU_ERR_INVALID_ARGUMENT, "bad mode");
    f->mode = mode;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_iir_decay_length(sgpu_iir *f, size_t *n) {
    if (!f || !n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    *n = 0;
    if (f->type != SGPU_IIR_SECOND_ORDER) return SGPU_OK;
    const long long d = decay_length(f);
    if (d > 0) *n = (size_t)d;
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_iir_transition(sgpu_iir *f, uint64_t n, double *A, size_t *dim) {
    if (!f || !dim) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    if (f->type != SGPU_IIR_SECOND_ORDER) return fail(SGPU_ERR_UNSUPPORTED, "transition matrix: second-order cascades only");
    const int D = 2 * f->nsec_pad, Da = 2 * f->nsec;
    *dim = (size_t)Da;
    if (!A) return SGPU_OK;
    std::vector<double> A1, An;
    build_transition(f, A1);  // over the kernel's (folded) coefficients, i.e. on device-scaled states v / gpre[s]
    mat_pow(A1, (long long)n, An, D);
    // get_state / set_state speak the reference's scaling: A_ref = G A_dev G^-1, G = diag(gpre); the identity padding
    // sections sit behind the real ones and the cascade is block lower triangular, so the leading Da x Da block is closed
    for (int i = 0; i < Da; ++i)
        for (int j = 0; j < Da; ++j) A[(size_t)i * Da + j] = An[(size_t)i * D + j] * f->gpre[i / 2] / f->gpre[j / 2];
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_iir_numerator_coefs(const sgpu_iir *f, double *out, size_t *n) {
    if (!f || !n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    const std::vector<double> &v = f->type == SGPU_IIR_SECOND_ORDER ? f->ff_raw : f->num_norm;
    if (out)
        for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    *n = v.size();
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_iir_denominator_coefs(const sgpu_iir *f, double *out, size_t *n) {
    if (!f || !n) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    const std::vector<double> &v = f->type == SGPU_IIR_SECOND_ORDER ? f->fb_raw : f->den_norm;
    if (out)
        for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    *n = v.size();
    return SGPU_OK;
}

SGPU_EXPORT int sgpu_iir_execute_block(sgpu_iir *f, const float *in, size_t n_in, size_t in_stride, float *out,
                                       size_t out_stride, size_t *n_out_p, sgpu_mem mem, void *stream) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    const size_t n_out = sgpu_iir_out_len(f, n_in);
    if (n_out_p) *n_out_p = n_out;
    if (n_in == 0) return SGPU_OK;
    if (!in || (n_out && !out)) return fail(SGPU_ERR_INVALID_ARGUMENT, "null buffer");
    if (f->C > 1 && in_stride < n_in) return fail(SGPU_ERR_INVALID_ARGUMENT, "in_stride < n_in");
    if (out_stride < n_out) return fail(SGPU_ERR_CAPACITY, "out capacity %zu < %zu outputs", out_stride, n_out);
    DeviceGuard g(f->device);
    cudaStream_t s = (cudaStream_t)stream;
    auto run = [f](const float2 *d_in, size_t nc, long long istr, float2 *d_out, long long ostr, size_t /*nout*/,
                   cudaStream_t st_) -> int {
        int st = iir_run(f, d_in, (long long)nc, istr, d_out, ostr, st_);
        if (st) return st;
        if (f->wrap == SGPU_IIR_DECIMATING) f->index = (f->index + nc) % f->factor;  // iir/decim.rs:225
        return SGPU_OK;
    };
    if (mem == SGPU_DEVICE)
        return run(reinterpret_cast<const float2 *>(in), n_in, (long long)in_stride, reinterpret_cast<float2 *>(out),
                   (long long)out_stride, n_out, s);
    return host_pipeline(f->pipe, f->C, in, n_in, in_stride, out, out_stride,
                         f->wrap == SGPU_IIR_INTERPOLATING ? f->factor : 1,
                         [f](size_t nc) { return sgpu_iir_out_len(f, nc); }, run, s);
}

SGPU_EXPORT int sgpu_iir_get_state(sgpu_iir *f, float *state, uint64_t *index) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    if (state) {
        SGPU_CUDA(cudaDeviceSynchronize());
        const size_t api = sgpu_iir_state_len(f), dev = state_len_dev(f);
        SGPU_CUDA(cudaMemcpy2D(state, api * sizeof(float2), f->d_state, dev * sizeof(float2), api * sizeof(float2),
                               f->C, cudaMemcpyDeviceToHost));
        if (f->fold)  // device states of section s are v / gpre[s]
            for (size_t c = 0; c < f->C; ++c)
                for (size_t i = 0; i < api; ++i) {
                    float *v = state + 2 * (c * api + i);
                    v[0] = (float)(v[0] * f->gpre[i / 2]);
                    v[1] = (float)(v[1] * f->gpre[i / 2]);
                }
    }
    if (index) *index = f->index;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_iir_set_state(sgpu_iir *f, const float *state, uint64_t index) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    if (state) {
        SGPU_CUDA(cudaDeviceSynchronize());
        const size_t api = sgpu_iir_state_len(f), dev = state_len_dev(f);
        SGPU_CUDA(cudaMemset(f->d_state, 0, f->C * dev * sizeof(float2)));
        std::vector<float> scaled;
        if (f->fold) {
            scaled.assign(state, state + 2 * f->C * api);
            for (size_t c = 0; c < f->C; ++c)
                for (size_t i = 0; i < api; ++i) {
                    float *v = scaled.data() + 2 * (c * api + i);
                    v[0] = (float)(v[0] / f->gpre[i / 2]);
                    v[1] = (float)(v[1] / f->gpre[i / 2]);
                }
            state = scaled.data();
        }
        SGPU_CUDA(cudaMemcpy2D(f->d_state, dev * sizeof(float2), state, api * sizeof(float2), api * sizeof(float2),
                               f->C, cudaMemcpyHostToDevice));
    }
    f->index = f->wrap == SGPU_IIR_DECIMATING ? index % f->factor : 0;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_iir_reset(sgpu_iir *f) {
    if (!f) return fail(SGPU_ERR_INVALID_ARGUMENT, "null handle");
    DeviceGuard g(f->device);
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemset(f->d_state, 0, f->C * state_len_dev(f) * sizeof(float2)));
    f->index = 0;
    return SGPU_OK;
}
SGPU_EXPORT int sgpu_iir_clone(const sgpu_iir *f, sgpu_iir **out) {
    if (!f || !out) return fail(SGPU_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard g(f->device);
    sgpu_iir *c = nullptr;
    int st = sgpu_iir_create((sgpu_iirtype)f->type, f->ff_raw.data(), f->ff_raw.size(), f->fb_raw.data(),
                             f->fb_raw.size(), f->C, (sgpu_iirwrap)f->wrap, f->factor, &c);
    if (st) return st;
    SGPU_CUDA(cudaDeviceSynchronize());
    SGPU_CUDA(cudaMemcpy(c->d_state, f->d_state, f->C * state_len_dev(f) * sizeof(float2), cudaMemcpyDeviceToDevice));
    c->index = f->index;
    c->mode = f->mode;
    *out = c;
    return SGPU_OK;
}
