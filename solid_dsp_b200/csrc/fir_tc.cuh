// Long real-tap FIR on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
// Interface between fir.cu (handle, dispatch) and fir_tc.cu (kernels, scratch, tensor maps).
#pragma once

#include "sgpu_common.cuh"

namespace sgpu {

struct FirTcState;  // opaque: banded tap matrix, split-plane scratch, tensor maps

// taps in caller order h[0..T) (interleaved re, im when complex_taps), already rounded to f32 (fir/mod.rs:79-88 keeps them reversed; we index
// g[i] = h[T-1-i] ourselves).  Returns SGPU_OK and *out = nullptr when the driver cannot encode tensor maps.
int fir_tc_create(FirTcState **out, const float *taps, int T, bool complex_taps);
// Polyphase bank (InterpolatingFIRFilter, pfb.rs:85-90): tp[p][j] multiplies x[n-j] for output L n + p; L in {1, 2, 4}
// (other L: *out stays nullptr and the caller keeps its FP32 kernels).
int fir_tc_create_pfb(FirTcState **out, const float *tp, int L, int S, bool complex_taps);
void fir_tc_destroy(FirTcState *st);

// C channels, one call: out[c][L n + p] = scale * sum_j tp[p][j] * x[c][n-j], x[c][<0] from hist[c] (the last H
// inputs, oldest first).  `in` and `out` must not overlap.  Two launches on `s`: the persistent tcgen05 kernel and
// fir_tc_post_kernel, which recomputes tiles that saw an Inf / NaN sample in the reference's order and, when hist_new
// is not NULL, writes the new history (the last H of old history ++ input, [C][H]) there.
int fir_tc_run(FirTcState *st, const float2 *in, long long n_in, long long in_stride, const float2 *hist, int H,
               float2 *hist_new, float2 *out, long long out_stride, size_t C, float scale, float scale_im, int sm_count,
               cudaStream_t s);

}  // namespace sgpu
