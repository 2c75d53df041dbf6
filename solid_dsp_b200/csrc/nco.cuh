// NCO (nco/mod.rs) on the device: a bank of 32-bit phase accumulators, one per channel, and the 1024-entry sine table.
// The decimator kernels (fir_walk.cuh, MIX = true) read the same table and the same [C][2] phase words when an sgpu_ddc
// fuses the mix-down into their tile loader.
#pragma once

#include "sgpu_common.cuh"

struct sgpu_nco {
    int device = 0;
    size_t C = 0;
    // phase of channel c at stream position i (samples consumed so far = pos): theta0[c] + i * delta[c], mod 2^32
    // (NCO::step, nco/mod.rs:93-96).  Host copy: [C][2] = (theta0, delta_theta); `dirty` = the device copy is stale.
    std::vector<uint32_t> tab;
    uint32_t pos = 0;
    unsigned *d_tab = nullptr;
    bool dirty = true;
    const float2 *d_lut = nullptr;  // per-device table, shared by all handles, never freed
    sgpu::Staging stage;
};

namespace sgpu {

// (cos, sin) pairs of the reference's table, rounded to f32: lut[i] = (table[(i + 256) & 1023], table[i]) with
// table[i] = sin(2 pi i / 1024) in f64 (nco/mod.rs:36-41,103-111).  One copy per device.
int nco_lut(int device, const float2 **out);
// upload the phase words if they changed since the last launch (on stream s)
int nco_sync_table(sgpu_nco *n, cudaStream_t s);
// out[c][i] = phasor(c, i) * in[c][i] (up) or conj(phasor(c, i)) * in[c][i] (down) for i < n (nco/mod.rs:141-151)
int nco_mix_launch(bool up, const float2 *in, long long in_stride, float2 *out, long long out_stride, long long n,
                   size_t C, const unsigned *d_tab, unsigned pos, const float2 *d_lut, cudaStream_t s);

}  // namespace sgpu
