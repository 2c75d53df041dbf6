//! Raw bindings to `include/solid_gpu.h` (ABI version 1).  One declaration per C prototype; see the
//! header for semantics and for the reference item (file:line) each entry point replaces.
#![allow(non_camel_case_types)]

use libc::{c_char, c_double, c_float, c_int, c_void, size_t};

pub const SGPU_OK: c_int = 0;
pub const SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO: c_int = -1;
pub const SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE: c_int = -2;
pub const SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE: c_int = -3;
pub const SGPU_ERR_FIR_NOT_ENOUGH_FILTERS: c_int = -4;
pub const SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO: c_int = -10;
pub const SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO: c_int = -11;
pub const SGPU_ERR_IIR_SOS_SIZE_ZERO: c_int = -12;
pub const SGPU_ERR_IIR_SOS_SIZE_MISMATCH: c_int = -13;
pub const SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3: c_int = -14;
pub const SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE: c_int = -15;
pub const SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE: c_int = -16;
pub const SGPU_ERR_SOS_COEFFICIENTS_NOT_IN_RANGE: c_int = -17;
pub const SGPU_ERR_FIRDES_BANDWIDTH: c_int = -20;
pub const SGPU_ERR_FIRDES_STOP_BAND_LEVEL: c_int = -21;
pub const SGPU_ERR_FIRDES_MU: c_int = -22;
pub const SGPU_ERR_INVALID_ARGUMENT: c_int = -30;
pub const SGPU_ERR_CAPACITY: c_int = -31;
pub const SGPU_ERR_CUDA: c_int = -32;
pub const SGPU_ERR_UNSUPPORTED: c_int = -33;
pub const SGPU_ERR_NO_DEVICE: c_int = -34;
pub const SGPU_ERR_ALLOC: c_int = -35;

pub const SGPU_HOST: c_int = 0;
pub const SGPU_DEVICE: c_int = 1;
pub const SGPU_TAPS_REAL: c_int = 0;
pub const SGPU_TAPS_COMPLEX: c_int = 1;
pub const SGPU_FORWARD: c_int = 0;
pub const SGPU_REVERSE: c_int = 1;
pub const SGPU_IIR_NORMAL: c_int = 0;
pub const SGPU_IIR_SECOND_ORDER: c_int = 1;
pub const SGPU_IIR_PLAIN: c_int = 0;
pub const SGPU_IIR_DECIMATING: c_int = 1;
pub const SGPU_IIR_INTERPOLATING: c_int = 2;

#[repr(C)] pub struct sgpu_fir { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_interp { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_iir { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_dot { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_autocorr { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_nco { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_ctx { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_sharded { _private: [u8; 0] }
#[repr(C)] pub struct sgpu_ddc { _private: [u8; 0] }
pub const SGPU_ALL_CHANNELS: size_t = usize::MAX;

extern "C" {
    pub fn sgpu_abi_version() -> c_int;
    pub fn sgpu_last_error() -> *const c_char;
    pub fn sgpu_status_name(status: c_int) -> *const c_char;
    pub fn sgpu_device_info(device: *mut c_int, sm_count: *mut c_int, cc_major: *mut c_int,
                            cc_minor: *mut c_int, total_mem: *mut size_t) -> c_int;
    pub fn sgpu_launch_count() -> u64;

    pub fn sgpu_fir_create(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                           scale_re: c_double, scale_im: c_double, is_decimator: c_int,
                           decimation: size_t, out: *mut *mut sgpu_fir) -> c_int;
    pub fn sgpu_fir_create_per_channel(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                                       scale_re: c_double, scale_im: c_double, is_decimator: c_int,
                                       decimation: size_t, out: *mut *mut sgpu_fir) -> c_int;
    pub fn sgpu_fir_taps_per_channel(f: *const sgpu_fir) -> c_int;
    pub fn sgpu_fir_channel_coefficients(f: *const sgpu_fir, channel: size_t, out: *mut c_double) -> c_int;
    pub fn sgpu_fir_destroy(f: *mut sgpu_fir) -> c_int;
    pub fn sgpu_fir_clone(f: *const sgpu_fir, out: *mut *mut sgpu_fir) -> c_int;
    pub fn sgpu_fir_execute_block(f: *mut sgpu_fir, input: *const c_float, n_in: size_t, in_stride: size_t,
                                  out: *mut c_float, out_stride: size_t, n_out: *mut size_t,
                                  mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_fir_write(f: *mut sgpu_fir, input: *const c_float, n_in: size_t, in_stride: size_t,
                          mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_fir_out_len(f: *const sgpu_fir, n_in: size_t) -> size_t;
    pub fn sgpu_fir_set_scale(f: *mut sgpu_fir, re: c_double, im: c_double) -> c_int;
    pub fn sgpu_fir_get_scale(f: *const sgpu_fir, re: *mut c_double, im: *mut c_double) -> c_int;
    pub fn sgpu_fir_len(f: *const sgpu_fir) -> size_t;
    pub fn sgpu_fir_decimation(f: *const sgpu_fir) -> size_t;
    pub fn sgpu_fir_channels(f: *const sgpu_fir) -> size_t;
    pub fn sgpu_fir_last_path(f: *const sgpu_fir) -> c_int;
    pub fn sgpu_fir_coefficients(f: *const sgpu_fir, out: *mut c_double) -> c_int;
    pub fn sgpu_fir_get_state(f: *mut sgpu_fir, history: *mut c_float, current_item: *mut u64) -> c_int;
    pub fn sgpu_fir_set_state(f: *mut sgpu_fir, history: *const c_float, current_item: u64) -> c_int;
    pub fn sgpu_fir_reset(f: *mut sgpu_fir) -> c_int;

    pub fn sgpu_interp_create(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                              interpolation: size_t, out: *mut *mut sgpu_interp) -> c_int;
    pub fn sgpu_pfb_create(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                           filters: size_t, scale_re: c_double, scale_im: c_double,
                           out: *mut *mut sgpu_interp) -> c_int;
    pub fn sgpu_interp_create_per_channel(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                                          interpolation: size_t, out: *mut *mut sgpu_interp) -> c_int;
    pub fn sgpu_interp_destroy(f: *mut sgpu_interp) -> c_int;
    pub fn sgpu_interp_clone(f: *const sgpu_interp, out: *mut *mut sgpu_interp) -> c_int;
    pub fn sgpu_interp_execute_block(f: *mut sgpu_interp, input: *const c_float, n_in: size_t,
                                     in_stride: size_t, out: *mut c_float, out_stride: size_t,
                                     n_out: *mut size_t, mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_interp_push(f: *mut sgpu_interp, input: *const c_float, n_in: size_t, in_stride: size_t,
                            mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_interp_execute_phase(f: *mut sgpu_interp, index: size_t, out: *mut c_float, mem: c_int,
                                     stream: *mut c_void) -> c_int;
    pub fn sgpu_interp_set_scale(f: *mut sgpu_interp, re: c_double, im: c_double) -> c_int;
    pub fn sgpu_interp_get_scale(f: *const sgpu_interp, re: *mut c_double, im: *mut c_double) -> c_int;
    pub fn sgpu_interp_interpolation(f: *const sgpu_interp) -> size_t;
    pub fn sgpu_interp_sub_len(f: *const sgpu_interp) -> size_t;
    pub fn sgpu_interp_channels(f: *const sgpu_interp) -> size_t;
    pub fn sgpu_interp_last_path(f: *const sgpu_interp) -> c_int;
    pub fn sgpu_interp_coefficients(f: *const sgpu_interp, out: *mut c_double) -> c_int;
    pub fn sgpu_interp_get_state(f: *mut sgpu_interp, history: *mut c_float) -> c_int;
    pub fn sgpu_interp_set_state(f: *mut sgpu_interp, history: *const c_float) -> c_int;
    pub fn sgpu_interp_reset(f: *mut sgpu_interp) -> c_int;

    pub fn sgpu_iir_create(kind: c_int, ff: *const c_double, n_ff: size_t, fb: *const c_double, n_fb: size_t,
                           n_channels: size_t, wrap: c_int, factor: size_t, out: *mut *mut sgpu_iir) -> c_int;
    pub fn sgpu_iir_destroy(f: *mut sgpu_iir) -> c_int;
    pub fn sgpu_iir_clone(f: *const sgpu_iir, out: *mut *mut sgpu_iir) -> c_int;
    pub fn sgpu_iir_execute_block(f: *mut sgpu_iir, input: *const c_float, n_in: size_t, in_stride: size_t,
                                  out: *mut c_float, out_stride: size_t, n_out: *mut size_t,
                                  mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_iir_out_len(f: *const sgpu_iir, n_in: size_t) -> size_t;
    pub fn sgpu_iir_sections(f: *const sgpu_iir) -> size_t;
    pub fn sgpu_iir_channels(f: *const sgpu_iir) -> size_t;
    pub fn sgpu_iir_type(f: *const sgpu_iir) -> c_int;
    pub fn sgpu_iir_numerator_coefs(f: *const sgpu_iir, out: *mut c_double, n: *mut size_t) -> c_int;
    pub fn sgpu_iir_denominator_coefs(f: *const sgpu_iir, out: *mut c_double, n: *mut size_t) -> c_int;
    pub fn sgpu_iir_get_state(f: *mut sgpu_iir, state: *mut c_float, index: *mut u64) -> c_int;
    pub fn sgpu_iir_set_state(f: *mut sgpu_iir, state: *const c_float, index: u64) -> c_int;
    pub fn sgpu_iir_reset(f: *mut sgpu_iir) -> c_int;
    pub fn sgpu_iir_state_len(f: *const sgpu_iir) -> size_t;
    pub fn sgpu_iir_set_mode(f: *mut sgpu_iir, mode: c_int) -> c_int;
    pub fn sgpu_iir_decay_length(f: *mut sgpu_iir, n: *mut usize) -> c_int;
    pub fn sgpu_iir_transition(f: *mut sgpu_iir, n: u64, a: *mut c_double, dim: *mut size_t) -> c_int;

    pub fn sgpu_autocorr_create(window_size: size_t, delay: size_t, n_channels: size_t,
                                out: *mut *mut sgpu_autocorr) -> c_int;
    pub fn sgpu_autocorr_destroy(f: *mut sgpu_autocorr) -> c_int;
    pub fn sgpu_autocorr_clone(f: *const sgpu_autocorr, out: *mut *mut sgpu_autocorr) -> c_int;
    pub fn sgpu_autocorr_window_size(f: *const sgpu_autocorr) -> size_t;
    pub fn sgpu_autocorr_delay(f: *const sgpu_autocorr) -> size_t;
    pub fn sgpu_autocorr_channels(f: *const sgpu_autocorr) -> size_t;
    pub fn sgpu_autocorr_execute_block(f: *mut sgpu_autocorr, input: *const c_float, n_in: size_t, in_stride: size_t,
                                       out: *mut c_float, out_stride: size_t, n_out: *mut size_t, mem: c_int,
                                       stream: *mut c_void) -> c_int;
    pub fn sgpu_autocorr_write(f: *mut sgpu_autocorr, input: *const c_float, n_in: size_t, in_stride: size_t,
                               mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_autocorr_execute(f: *mut sgpu_autocorr, out: *mut c_double) -> c_int;
    pub fn sgpu_autocorr_get_energy(f: *mut sgpu_autocorr, out: *mut c_double) -> c_int;
    pub fn sgpu_autocorr_reset(f: *mut sgpu_autocorr) -> c_int;
    pub fn sgpu_autocorr_get_state(f: *mut sgpu_autocorr, state: *mut c_float) -> c_int;
    pub fn sgpu_autocorr_set_state(f: *mut sgpu_autocorr, state: *const c_float) -> c_int;

    pub fn sgpu_dot_create(coefs: *const c_double, n: size_t, kind: c_int, dir: c_int,
                           out: *mut *mut sgpu_dot) -> c_int;
    pub fn sgpu_dot_destroy(d: *mut sgpu_dot) -> c_int;
    pub fn sgpu_dot_len(d: *const sgpu_dot) -> size_t;
    pub fn sgpu_dot_coefficients(d: *const sgpu_dot, out: *mut c_double) -> c_int;
    pub fn sgpu_dot_execute(d: *mut sgpu_dot, x: *const c_float, n_x: size_t, x_stride: size_t, n_vec: size_t,
                            result: *mut c_float, mem: c_int, stream: *mut c_void) -> c_int;

    pub fn sgpu_host_alloc(bytes: size_t, device: c_int, out: *mut *mut c_void) -> c_int;
    pub fn sgpu_host_free(p: *mut c_void) -> c_int;

    pub fn sgpu_firdes_kaiser(filter_length: size_t, cutoff_frequency: *const f64, stop_band_attenuation: *const f64,
                              fractional_sample_offset: *const f64, n_designs: size_t, out: *mut f64, mem: c_int,
                              stream: *mut c_void) -> c_int;
    pub fn sgpu_nco_create(n_channels: size_t, out: *mut *mut sgpu_nco) -> c_int;
    pub fn sgpu_nco_destroy(n: *mut sgpu_nco) -> c_int;
    pub fn sgpu_nco_clone(n: *const sgpu_nco, out: *mut *mut sgpu_nco) -> c_int;
    pub fn sgpu_nco_channels(n: *const sgpu_nco) -> size_t;
    pub fn sgpu_nco_reset(n: *mut sgpu_nco) -> c_int;
    pub fn sgpu_nco_set_frequency(n: *mut sgpu_nco, channel: size_t, delta_theta: c_double) -> c_int;
    pub fn sgpu_nco_adjust_frequency(n: *mut sgpu_nco, channel: size_t, dt: c_double) -> c_int;
    pub fn sgpu_nco_set_phase(n: *mut sgpu_nco, channel: size_t, phi: c_double) -> c_int;
    pub fn sgpu_nco_adjust_phase(n: *mut sgpu_nco, channel: size_t, delta_phi: c_double) -> c_int;
    pub fn sgpu_nco_step(n: *mut sgpu_nco, count: u64) -> c_int;
    pub fn sgpu_nco_get(n: *const sgpu_nco, channel: size_t, theta: *mut u32, delta_theta: *mut u32) -> c_int;
    pub fn sgpu_nco_set(n: *mut sgpu_nco, channel: size_t, theta: u32, delta_theta: u32) -> c_int;
    pub fn sgpu_nco_constrain(theta: c_double) -> u32;
    pub fn sgpu_nco_mix_block(n: *mut sgpu_nco, up: c_int, input: *const c_float, n_in: size_t, in_stride: size_t,
                              out: *mut c_float, out_stride: size_t, mem: c_int, stream: *mut c_void) -> c_int;

    pub fn sgpu_ddc_create(taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                           scale_re: c_double, scale_im: c_double, decimation: size_t,
                           out: *mut *mut sgpu_ddc) -> c_int;
    pub fn sgpu_ddc_destroy(d: *mut sgpu_ddc) -> c_int;
    pub fn sgpu_ddc_clone(d: *const sgpu_ddc, out: *mut *mut sgpu_ddc) -> c_int;
    pub fn sgpu_ddc_filter(d: *mut sgpu_ddc) -> *mut sgpu_fir;
    pub fn sgpu_ddc_nco(d: *mut sgpu_ddc) -> *mut sgpu_nco;
    pub fn sgpu_ddc_out_len(d: *const sgpu_ddc, n_in: size_t) -> size_t;
    pub fn sgpu_ddc_execute_block(d: *mut sgpu_ddc, input: *const c_float, n_in: size_t, in_stride: size_t,
                                  out: *mut c_float, out_stride: size_t, n_out: *mut size_t,
                                  mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_ddc_write(d: *mut sgpu_ddc, input: *const c_float, n_in: size_t, in_stride: size_t,
                          mem: c_int, stream: *mut c_void) -> c_int;
    pub fn sgpu_ddc_reset(d: *mut sgpu_ddc) -> c_int;
    pub fn sgpu_ddc_last_fused(d: *const sgpu_ddc) -> c_int;

    pub fn sgpu_ctx_create(n_gpus: c_int, out: *mut *mut sgpu_ctx) -> c_int;
    pub fn sgpu_ctx_create_devices(devices: *const c_int, n: c_int, out: *mut *mut sgpu_ctx) -> c_int;
    pub fn sgpu_ctx_destroy(ctx: *mut sgpu_ctx) -> c_int;
    pub fn sgpu_ctx_devices(ctx: *const sgpu_ctx) -> c_int;
    pub fn sgpu_ctx_fir_create(ctx: *mut sgpu_ctx, taps: *const c_double, n_taps: size_t, kind: c_int, n_channels: size_t,
                               scale_re: c_double, scale_im: c_double, is_decimator: c_int, decimation: size_t,
                               out: *mut *mut sgpu_sharded) -> c_int;
    pub fn sgpu_ctx_interp_create(ctx: *mut sgpu_ctx, taps: *const c_double, n_taps: size_t, kind: c_int,
                                  n_channels: size_t, interpolation: size_t, out: *mut *mut sgpu_sharded) -> c_int;
    pub fn sgpu_ctx_iir_create(ctx: *mut sgpu_ctx, kind: c_int, ff: *const c_double, n_ff: size_t, fb: *const c_double,
                               n_fb: size_t, n_channels: size_t, wrap: c_int, factor: size_t,
                               out: *mut *mut sgpu_sharded) -> c_int;
    pub fn sgpu_sharded_destroy(f: *mut sgpu_sharded) -> c_int;
    pub fn sgpu_sharded_shards(f: *const sgpu_sharded) -> c_int;
    pub fn sgpu_sharded_shard_info(f: *const sgpu_sharded, index: c_int, device: *mut c_int, first_channel: *mut size_t,
                                   n_channels: *mut size_t) -> c_int;
    pub fn sgpu_sharded_last_segments(f: *const sgpu_sharded) -> c_int;
    pub fn sgpu_sharded_out_len(f: *const sgpu_sharded, n_in: size_t) -> size_t;
    pub fn sgpu_sharded_reset(f: *mut sgpu_sharded) -> c_int;
    pub fn sgpu_sharded_execute_block(f: *mut sgpu_sharded, input: *const c_float, n_in: size_t, in_stride: size_t,
                                      out: *mut c_float, out_stride: size_t, n_out: *mut size_t) -> c_int;

    pub fn sgpu_shard_channels(n_channels: size_t, world: c_int, rank: c_int, first: *mut size_t,
                               count: *mut size_t) -> c_int;
    pub fn sgpu_shard_stream(n_samples: size_t, align: size_t, world: c_int, rank: c_int,
                             first: *mut size_t, count: *mut size_t) -> c_int;
}
