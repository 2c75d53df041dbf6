// Long real-tap FIR on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
// Interface between fir.cu (handle, dispatch) and fir_tc.cu (kernels, scratch, tensor maps).
#pragma once

#include "sgpu_common.cuh"

namespace sgpu {

struct FirTcState;  // opaque: banded tap matrix, split-plane scratch, tensor maps

// taps in caller order h[0..T), already rounded to f32 (fir/mod.rs:79-88 keeps them reversed; we index
// g[i] = h[T-1-i] ourselves).  Returns SGPU_OK and *out = nullptr when the driver cannot encode tensor maps.
int fir_tc_create(FirTcState **out, const float *taps, int T);
void fir_tc_destroy(FirTcState *st);

// One channel, one call: out[n] = scale * sum_i h[T-1-i] * x[n-i], x[<0] from hist (last T-1 inputs, oldest
// first).  `in` and `out` must not overlap.  Launches the split pre-pass and the tcgen05 kernel on `s`.
int fir_tc_run(FirTcState *st, const float2 *in, long long n_in, const float2 *hist, float2 *out, float scale,
               int sm_count, cudaStream_t s);

}  // namespace sgpu
