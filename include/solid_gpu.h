/*
 * solid_gpu.h -- C ABI of libsolid_gpu.so: the B200 (sm_100a) implementation of
 * juliantos/solid-dsp's filtering hot path.
 *
 * The reference is a pure-Rust crate with NO native boundary of its own (no extern "C", no
 * build.rs -- SURVEY.md "Ground facts"), so these entry points are what a `solid-gpu-sys`
 * crate binds (INTEGRATION.md shows the Rust side).  Every function cites the reference
 * item it replaces as <file>:<line> relative to the reference's src/.
 *
 * Conventions
 *   - Every function returns an int status: SGPU_OK (0) or a negative sgpu_status.
 *     Construction errors map one-to-one onto the reference's error enums; execute never
 *     fails in the reference, here it can only fail with SGPU_ERR_CUDA / _CAPACITY /
 *     _INVALID_ARGUMENT (the latter includes `in` and `out` ranges that overlap: execute_block
 *     is not an in-place operation, the kernels read a tile's halo while other blocks already
 *     write outputs).  sgpu_last_error() returns a thread-local message.
 *   - Samples are cf32: interleaved (re, im) floats, 8 bytes per sample.  Buffers are
 *     channel-major: channel c starts at base + c * stride (stride in SAMPLES).  One
 *     "channel" is one reference filter object; the channels of a handle advance in
 *     lock-step and share the taps unless the handle was made by a *_create_per_channel
 *     constructor (every reference object owns its coefficients, fir/mod.rs:79-88).
 *   - Coefficients cross the boundary as doubles (the reference's Coef = f64 /
 *     Complex<f64>) and are rounded once to f32 inside; arithmetic is f32 FMA.
 *   - `mem` says where `in`/`out` live.  SGPU_DEVICE: pointers are device pointers on the
 *     handle's device, work is enqueued on `stream` (a cudaStream_t, NULL = default stream)
 *     and the call returns without synchronising.  SGPU_HOST: pageable or pinned host
 *     memory; the call stages through device buffers owned by the handle and returns after
 *     the result is in `out`.  Pinned buffers (sgpu_host_alloc, cudaHostAlloc, cudaHostRegister)
 *     are copied directly at the PCIe rate; pageable buffers of 4 MiB or more go through pinned
 *     staging buffers of the handle, filled and drained by host threads while the neighbouring
 *     chunk is on the bus (SGPU_HOST_STAGING=0: the driver's own staged copies).  `in` and
 *     `out` must not overlap.
 *   - A handle is one logical stream of calls (like `&mut self`): not thread-safe per
 *     handle; distinct handles may be used from distinct threads.
 *   - No CPU fallback exists: without a usable sm_100 device every create returns
 *     SGPU_ERR_NO_DEVICE.
 */
#ifndef SOLID_GPU_H
#define SOLID_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGPU_ABI_VERSION 1

typedef enum sgpu_status {
    SGPU_OK = 0,
    /* fir/mod.rs:39-45  FIRErrorCode */
    SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO = -1,
    SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE = -2,
    SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE = -3,
    SGPU_ERR_FIR_NOT_ENOUGH_FILTERS = -4,
    /* iir/mod.rs:40-49  IIRErrorCode */
    SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO = -10,
    SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO = -11,
    SGPU_ERR_IIR_SOS_SIZE_ZERO = -12,
    SGPU_ERR_IIR_SOS_SIZE_MISMATCH = -13,
    SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3 = -14,
    SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE = -15,
    SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE = -16,
    /* iir/sos.rs:18-21  SecondOrderErrorCode::CoefficientsNotInRange */
    SGPU_ERR_SOS_COEFFICIENTS_NOT_IN_RANGE = -17,
    /* firdes/mod.rs:17-25  FirdesErrorCode (the variants firdes_kaiser can return, :284-290) */
    SGPU_ERR_FIRDES_BANDWIDTH = -20,
    SGPU_ERR_FIRDES_STOP_BAND_LEVEL = -21,
    SGPU_ERR_FIRDES_MU = -22,
    /* library-side */
    SGPU_ERR_INVALID_ARGUMENT = -30,
    SGPU_ERR_CAPACITY = -31, /* `out` too small for the outputs this call produces */
    SGPU_ERR_CUDA = -32,
    SGPU_ERR_UNSUPPORTED = -33,
    SGPU_ERR_NO_DEVICE = -34,
    SGPU_ERR_ALLOC = -35
} sgpu_status;

typedef enum sgpu_mem { SGPU_HOST = 0, SGPU_DEVICE = 1 } sgpu_mem;
typedef enum sgpu_tapkind { SGPU_TAPS_REAL = 0, SGPU_TAPS_COMPLEX = 1 } sgpu_tapkind;
/* dot_product/mod.rs:31-34 */
typedef enum sgpu_direction { SGPU_FORWARD = 0, SGPU_REVERSE = 1 } sgpu_direction;
/* iir/mod.rs:62-66 */
typedef enum sgpu_iirtype { SGPU_IIR_NORMAL = 0, SGPU_IIR_SECOND_ORDER = 1 } sgpu_iirtype;
typedef enum sgpu_iirwrap { SGPU_IIR_PLAIN = 0, SGPU_IIR_DECIMATING = 1, SGPU_IIR_INTERPOLATING = 2 } sgpu_iirwrap;

typedef struct sgpu_fir sgpu_fir;       /* FIRFilter / DecimatingFIRFilter        */
typedef struct sgpu_interp sgpu_interp; /* InterpolatingFIRFilter (+ its PolyPhaseFilterBank) */
typedef struct sgpu_iir sgpu_iir;       /* IIRFilter / Decimating- / InterpolatingIIRFilter   */
typedef struct sgpu_dot sgpu_dot;       /* DotProduct                              */
typedef struct sgpu_autocorr sgpu_autocorr; /* AutoCorrelator                      */
typedef struct sgpu_ctx sgpu_ctx;         /* a set of GPUs of this box                */
typedef struct sgpu_sharded sgpu_sharded; /* one filter object spread over a context   */
typedef struct sgpu_nco sgpu_nco;       /* NCO (one phase accumulator per channel) */
typedef struct sgpu_ddc sgpu_ddc;       /* NCO mix-down -> DecimatingFIRFilter      */

#define SGPU_ALL_CHANNELS ((size_t)-1)

/* ---- library ------------------------------------------------------------------------ */
int sgpu_abi_version(void);
const char *sgpu_last_error(void);
const char *sgpu_status_name(int status);
/* Properties of the current CUDA device (cudaGetDevice). */
/* Pinned host memory placed on the NUMA node of `device` (-1: the current device) for the SGPU_HOST calls: the
 * reference's callers hand over `&[In]` slices in ordinary host memory (filter/mod.rs:14); a GPU caller that wants the
 * PCIe ceiling allocates its sample buffers here.  Free with sgpu_host_free. */
int sgpu_host_alloc(size_t bytes, int device, void **out);
int sgpu_host_free(void *p);
int sgpu_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t sgpu_launch_count(void);

/* ---- FIRFilter / DecimatingFIRFilter -------------------------------------------------
 * y[n] = scale * sum_{i<T} h[T-1-i] * x[n-i]   (taps applied REVERSED: dot_product/mod.rs:75-84
 * via DotProduct::new(.., REVERSE) at fir/mod.rs:86; newest sample first, window/mod.rs:63-71)
 * Decimator: an output is emitted for every pushed sample n with (count+1) % M == 0, the
 * counter persisting across calls (fir/decim.rs:115-118,221-228).
 *
 * sgpu_fir_create          replaces FIRFilter::new (fir/mod.rs:79-88) when is_decimator == 0
 *                          and DecimatingFIRFilter::new (fir/decim.rs:27-42) when 1.
 *   taps: n_taps doubles (REAL) or 2*n_taps doubles (COMPLEX), in the caller's order h[0..T).
 *   scale: Coef-typed; scale_im is ignored for real taps.
 */
int sgpu_fir_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                    double scale_re, double scale_im, int is_decimator, size_t decimation,
                    sgpu_fir **out);
/* As sgpu_fir_create, but every channel owns its taps: taps holds n_channels * n_taps values
 * (x2 when complex), channel c uses [c * n_taps, (c + 1) * n_taps) -- n_channels independent
 * FIRFilter::new / DecimatingFIRFilter::new objects of equal length in one handle.  These
 * handles run on the FP32 kernels (the tensor-core band matrix is shared by all channels). */
int sgpu_fir_create_per_channel(const double *taps, size_t n_taps, sgpu_tapkind kind,
                                size_t n_channels, double scale_re, double scale_im,
                                int is_decimator, size_t decimation, sgpu_fir **out);
int sgpu_fir_taps_per_channel(const sgpu_fir *f); /* 1 for a *_per_channel handle */
int sgpu_fir_destroy(sgpu_fir *f); /* Drop */
/* #[derive(Clone)] fir/mod.rs:58 -- deep copy incl. history and decimator phase */
int sgpu_fir_clone(const sgpu_fir *f, sgpu_fir **out);
/* Filter::execute_block (fir/mod.rs:235-241, fir/decim.rs:250-256); n_in == 1 is
 * Filter::execute.  Writes *n_out outputs per channel (== sgpu_fir_out_len(f, n_in)).
 * out_stride is also the per-channel capacity of `out`. */
int sgpu_fir_execute_block(sgpu_fir *f, const float *in, size_t n_in, size_t in_stride,
                           float *out, size_t out_stride, size_t *n_out, sgpu_mem mem,
                           void *stream);
/* DecimatingFIRFilter::write / ::push (fir/decim.rs:115-118,136-139): feed history and
 * advance the phase counter without producing output.  Also valid on a plain FIR handle
 * (Window::write, window/mod.rs:73-77). */
int sgpu_fir_write(sgpu_fir *f, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
                   void *stream);
/* Outputs per channel the next execute_block(n_in) produces: n_in for a FIR,
 * floor((current_item + n_in) / M) for a decimator. */
size_t sgpu_fir_out_len(const sgpu_fir *f, size_t n_in);
int sgpu_fir_set_scale(sgpu_fir *f, double re, double im);       /* fir/mod.rs:106, decim.rs:60 */
int sgpu_fir_get_scale(const sgpu_fir *f, double *re, double *im); /* fir/mod.rs:124, decim.rs:78 */
size_t sgpu_fir_len(const sgpu_fir *f);                           /* fir/mod.rs:142 */
size_t sgpu_fir_decimation(const sgpu_fir *f);                    /* decim.rs:96; 1 for a FIR */
size_t sgpu_fir_channels(const sgpu_fir *f);
/* Which arithmetic path the last execute_block took: 0 = FP32 FFMA2 kernels, 1 = tcgen05 tensor cores (long
 * real-tap FIR as a banded-Toeplitz product, csrc/fir_tc.cu).  Reporting only (bench.py roofline). */
int sgpu_fir_last_path(const sgpu_fir *f);
/* FIRFilter::coefficients (fir/mod.rs:176-178): the STORED (reversed) order, as doubles of
 * the f32 values used on the device.  out: n_taps (REAL) or 2*n_taps (COMPLEX) doubles. */
int sgpu_fir_coefficients(const sgpu_fir *f, double *out);
/* The same for one channel of a *_per_channel handle (any handle: channel < n_channels). */
int sgpu_fir_channel_coefficients(const sgpu_fir *f, size_t channel, double *out);
/* Streaming state (what Window + current_item hold): history = last T-1 inputs per channel,
 * oldest first, [n_channels][T-1] cf32 on the HOST; current_item as in fir/decim.rs:8. */
int sgpu_fir_get_state(sgpu_fir *f, float *history, uint64_t *current_item);
int sgpu_fir_set_state(sgpu_fir *f, const float *history, uint64_t current_item);
int sgpu_fir_reset(sgpu_fir *f); /* zero history, current_item = 0 (Window::reset, window/mod.rs:54) */

/* ---- InterpolatingFIRFilter / PolyPhaseFilterBank ------------------------------------
 * y[n*L + p] = sum_{j<S} hpad[p + (S-1-j)*L] * x[n-j],  S = ceil(T/L) computed in f32 as
 * fir/interp.rs:35-40 does, hpad = taps zero-padded to S*L (interp.rs:43-46); NO scale is
 * applied (pfb.rs:85-90 never reads self.scale); the stored scale is observable only
 * through get_scale.
 * sgpu_interp_create replaces InterpolatingFIRFilter::new (fir/interp.rs:27-54).
 * sgpu_pfb_create    replaces PolyPhaseFilterBank::new (fir/pfb.rs:24-49): sub_len =
 *                    n_taps / filters (truncating, pfb.rs:32); filters > n_taps is
 *                    SGPU_ERR_FIR_NOT_ENOUGH_FILTERS instead of the reference's panic. */
int sgpu_interp_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                       size_t interpolation, sgpu_interp **out);
int sgpu_pfb_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                    size_t filters, double scale_re, double scale_im, sgpu_interp **out);
/* Every channel owns its taps ([n_channels][n_taps]), see sgpu_fir_create_per_channel. */
int sgpu_interp_create_per_channel(const double *taps, size_t n_taps, sgpu_tapkind kind,
                                   size_t n_channels, size_t interpolation, sgpu_interp **out);
int sgpu_interp_destroy(sgpu_interp *f);
int sgpu_interp_clone(const sgpu_interp *f, sgpu_interp **out);
/* Filter::execute_block (fir/interp.rs:102-111): *n_out = n_in * L per channel. */
int sgpu_interp_execute_block(sgpu_interp *f, const float *in, size_t n_in, size_t in_stride,
                              float *out, size_t out_stride, size_t *n_out, sgpu_mem mem,
                              void *stream);
/* PolyPhaseFilterBank::push (pfb.rs:81-83) for n_in samples per channel. */
int sgpu_interp_push(sgpu_interp *f, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
                     void *stream);
/* PolyPhaseFilterBank::execute(index) (pfb.rs:85-90): one output per channel, out[c] cf32. */
int sgpu_interp_execute_phase(sgpu_interp *f, size_t index, float *out, sgpu_mem mem, void *stream);
int sgpu_interp_set_scale(sgpu_interp *f, double re, double im);        /* interp.rs:57, pfb.rs:52 */
int sgpu_interp_get_scale(const sgpu_interp *f, double *re, double *im);/* interp.rs:62, pfb.rs:57 */
size_t sgpu_interp_interpolation(const sgpu_interp *f);                  /* interp.rs:82 / pfb.rs:62 */
size_t sgpu_interp_sub_len(const sgpu_interp *f);
size_t sgpu_interp_channels(const sgpu_interp *f);
int sgpu_interp_last_path(const sgpu_interp *f); /* 0 = FP32 kernels, 1 = tcgen05 tensor cores (see sgpu_fir_last_path) */
/* InterpolatingFIRFilter::coefficents (interp.rs:77-79): per-phase stored order flattened,
 * L*S values. */
int sgpu_interp_coefficients(const sgpu_interp *f, double *out);
/* history = last S-1 inputs per channel, oldest first, [n_channels][S-1] cf32 on the HOST */
int sgpu_interp_get_state(sgpu_interp *f, float *history);
int sgpu_interp_set_state(sgpu_interp *f, const float *history);
int sgpu_interp_reset(sgpu_interp *f); /* pfb.rs:76-78 */

/* ---- IIRFilter (+ Decimating / Interpolating wrappers) -------------------------------
 * SecondOrder: ff, fb are flat arrays of 3*nsec values, section i uses [3i, 3i+3)
 * (iir/mod.rs:144-153), each normalised by its own fb[3i] (sos.rs:62-68); per section
 *   v0 = x - (a1*v1 + a2*v2);  y = b0*v0 + b1*v1 + b2*v2            (sos.rs:92-114)
 * Normal: one direct-form II of arbitrary order (iir/mod.rs:98-130,272-280).
 * wrap/factor: DecimatingIIRFilter keeps outputs where (index+1) % M == 0
 * (iir/decim.rs:190-198,222-233); InterpolatingIIRFilter feeds each sample then L-1 zeros
 * (iir/interp.rs:184-190).
 * sgpu_iir_create replaces IIRFilter::new (iir/mod.rs:92-164), DecimatingIIRFilter::new
 * (iir/decim.rs:30-47) and InterpolatingIIRFilter::new (iir/interp.rs:29-46). */
int sgpu_iir_create(sgpu_iirtype type, const double *ff, size_t n_ff, const double *fb, size_t n_fb,
                    size_t n_channels, sgpu_iirwrap wrap, size_t factor, sgpu_iir **out);
int sgpu_iir_destroy(sgpu_iir *f);
int sgpu_iir_clone(const sgpu_iir *f, sgpu_iir **out);
/* Filter::execute_block (iir/mod.rs:310-316, iir/decim.rs:222-233, iir/interp.rs:215-221) */
int sgpu_iir_execute_block(sgpu_iir *f, const float *in, size_t n_in, size_t in_stride,
                           float *out, size_t out_stride, size_t *n_out, sgpu_mem mem,
                           void *stream);
size_t sgpu_iir_out_len(const sgpu_iir *f, size_t n_in);
size_t sgpu_iir_sections(const sgpu_iir *f);  /* second_order_filters().len(), iir/mod.rs:222 */
size_t sgpu_iir_channels(const sgpu_iir *f);
int sgpu_iir_type(const sgpu_iir *f);         /* iir/mod.rs:239 */
/* numerator_coefs / denominator_coefs (iir/mod.rs:182-205): as stored by the reference --
 * SecondOrder: the RAW flat ff / fb; Normal: ff/a0 and fb[1..]/a0. */
int sgpu_iir_numerator_coefs(const sgpu_iir *f, double *out, size_t *n);
int sgpu_iir_denominator_coefs(const sgpu_iir *f, double *out, size_t *n);
/* state: SecondOrder [n_channels][nsec][2] cf32 = (v1, v2) per section; Normal
 * [n_channels][order] cf32 newest first; HOST memory.  index = decimator counter. */
int sgpu_iir_get_state(sgpu_iir *f, float *state, uint64_t *index);
int sgpu_iir_set_state(sgpu_iir *f, const float *state, uint64_t index);
int sgpu_iir_reset(sgpu_iir *f);
size_t sgpu_iir_state_len(const sgpu_iir *f); /* complex values per channel */
/* Execution strategy: -1 auto, 0 one channel per thread (batch), 1 chunked scan (fused warm-up scan
 * when the filter's memory decays within 2^16 samples, else the three-pass scan), 2 three-pass scan
 * (zero-state pass, f64 carry recurrence, output pass) regardless of the decay. */
int sgpu_iir_set_mode(sgpu_iir *f, int mode);
/* Memory of a second-order cascade in samples: the smallest multiple of 32 after which the influence
 * of an older state is below 1e-10 (infinity norm of the zero-input transition matrix power), 0 when
 * the filter does not decay within 2^16 samples or is in Normal mode.  A stream cut into segments
 * (one per GPU) stays within 1e-10 of the unbroken recurrence when every segment but the first is
 * preceded by that many samples of the previous segment: reset, execute_block(halo) with the output
 * discarded, then execute_block(segment).  No counterpart in the reference (it is single-threaded). */
int sgpu_iir_decay_length(sgpu_iir *f, size_t *n);
/* A^n: how the state of a second-order cascade evolves over n samples of ZERO input, as a D x D matrix of doubles
 * (row-major, D = *dim = 2 * sections, state order and scaling of sgpu_iir_get_state), built in f64 from the f32
 * coefficients the kernels use.  With the end state z of a segment run from zero state, the state after the segment
 * from any start state s is A^n s + z (iir/mod.rs:281-287 and sos.rs:92-114 are linear): that is all the ranks of a
 * stream cut into time segments have to exchange to make the cut exact for ANY filter (SURVEY 8e row 3;
 * solid_dsp_b200/sharding.py: iir_segment_exact).  A == NULL: only *dim is written. */
int sgpu_iir_transition(sgpu_iir *f, uint64_t n, double *A, size_t *dim);

/* ---- AutoCorrelator ------------------------------------------------------------------
 * filter/auto_correlator/mod.rs: new(window_size, delay) :51-62, push :99-111, write :130-141,
 * execute :165-172, execute_block :184-191, get_energy :214-216, reset :76-85.
 * With W = window_size, d = delay and the reference's Window(capacity, delay) semantics
 * (window/mod.rs:17-34,44-51,63-71: the delayed window's tail is never written),
 *     r[n] = sum_{i < W-d} x[n-i] * conj(x[n-d-i])      (0 for d >= W)
 *     energy = sum_{i < W} |x[n-i]|^2
 * One handle = n_channels independent correlators (one reference object each), channel-major
 * buffers like every other handle.  execute_block: one output per input.  execute / get_energy
 * report the current window: [n_channels] complex doubles / doubles in HOST memory.
 * State = the last window_size samples per channel, oldest first. */
int sgpu_autocorr_create(size_t window_size, size_t delay, size_t n_channels, sgpu_autocorr **out);
int sgpu_autocorr_destroy(sgpu_autocorr *f);
int sgpu_autocorr_clone(const sgpu_autocorr *f, sgpu_autocorr **out);
size_t sgpu_autocorr_window_size(const sgpu_autocorr *f);
size_t sgpu_autocorr_delay(const sgpu_autocorr *f);
size_t sgpu_autocorr_channels(const sgpu_autocorr *f);
int sgpu_autocorr_execute_block(sgpu_autocorr *f, const float *in, size_t n_in, size_t in_stride,
                                float *out, size_t out_stride, size_t *n_out, sgpu_mem mem,
                                void *stream);
int sgpu_autocorr_write(sgpu_autocorr *f, const float *in, size_t n_in, size_t in_stride,
                        sgpu_mem mem, void *stream);
int sgpu_autocorr_execute(sgpu_autocorr *f, double *out);
int sgpu_autocorr_get_energy(sgpu_autocorr *f, double *out);
int sgpu_autocorr_reset(sgpu_autocorr *f);
int sgpu_autocorr_get_state(sgpu_autocorr *f, float *state);
int sgpu_autocorr_set_state(sgpu_autocorr *f, const float *state);

/* ---- DotProduct ----------------------------------------------------------------------
 * sum_{i < min(len_c, len_x)} c[i] * x[i], c stored FORWARD or REVERSED
 * (dot_product/mod.rs:57-87,153-171). */
int sgpu_dot_create(const double *coefs, size_t n, sgpu_tapkind kind, sgpu_direction dir,
                    sgpu_dot **out);
int sgpu_dot_destroy(sgpu_dot *d);
size_t sgpu_dot_len(const sgpu_dot *d);                 /* dot_product/mod.rs:124 */
int sgpu_dot_coefficients(const sgpu_dot *d, double *out); /* dot_product/mod.rs:102-109 */
/* Execute::execute (dot_product/mod.rs:159-170) on n_vec independent sample vectors of
 * length n_x each (vector v at x + v*x_stride); result[v] cf32. */
int sgpu_dot_execute(sgpu_dot *d, const float *x, size_t n_x, size_t x_stride, size_t n_vec,
                     float *result, sgpu_mem mem, void *stream);

/* ---- tap design on the device ---------------------------------------
 * firdes_kaiser (firdes/mod.rs:278-305; kaiser_beta :243-253, windows/kaiser.rs:33-46, math/mod.rs:17-27,41-100,
 * 171-183) for n_designs filters of filter_length taps each in one launch, f64, one thread per tap: design d uses
 * cutoff_frequency[d], stop_band_attenuation[d], fractional_sample_offset[d] (host arrays; the last may be NULL = 0)
 * and writes out[d * filter_length ..] (`mem` says where `out` lives; SGPU_HOST returns after the copy).  The result
 * is what sgpu_fir_create / _create_per_channel take as `taps`.  Errors in the reference's order (:284-290):
 * SGPU_ERR_FIRDES_MU, _BANDWIDTH, _STOP_BAND_LEVEL. */
int sgpu_firdes_kaiser(size_t filter_length, const double *cutoff_frequency, const double *stop_band_attenuation,
                       const double *fractional_sample_offset, size_t n_designs, double *out, sgpu_mem mem,
                       void *stream);

/* ---- NCO (nco/mod.rs) and the digital down-converter ----------------------------------
 * the per-sample step next to the decimator.  An sgpu_nco is C reference NCO
 * objects (nco/mod.rs:26-33): a 32-bit phase `theta`, a 32-bit step `delta_theta`, the
 * 1024-entry sine table (:36-41; rounded to f32 on the device).  `channel` is an index or
 * SGPU_ALL_CHANNELS.  The reference's mix_up_block / mix_down_block (:153-172) index an empty
 * Vec and panic; sgpu_nco_mix_block is the loop they were written to be: y[i] = mix(x[i]); step().
 * An sgpu_ddc is that loop (mix_down) feeding a DecimatingFIRFilter (fir/decim.rs:221-256), one
 * pair per channel; on the hot shapes (M in {2, 4, 8}, real taps) the mixing happens inside the
 * decimator's tile loader and the mixed stream never reaches HBM (sgpu_ddc_last_fused). */
int sgpu_nco_create(size_t n_channels, sgpu_nco **out);                           /* NCO::new, :36-50 */
int sgpu_nco_destroy(sgpu_nco *n);
int sgpu_nco_clone(const sgpu_nco *n, sgpu_nco **out);
size_t sgpu_nco_channels(const sgpu_nco *n);
int sgpu_nco_reset(sgpu_nco *n);                                                  /* :53-56 */
int sgpu_nco_set_frequency(sgpu_nco *n, size_t channel, double delta_theta);      /* :59-61 */
int sgpu_nco_adjust_frequency(sgpu_nco *n, size_t channel, double dt);            /* :64-66 */
int sgpu_nco_set_phase(sgpu_nco *n, size_t channel, double phi);                  /* :79-81 */
int sgpu_nco_adjust_phase(sgpu_nco *n, size_t channel, double delta_phi);         /* :84-86 */
int sgpu_nco_step(sgpu_nco *n, uint64_t count);                                   /* :93-96, `count` times */
/* raw accumulator words (theta as of the next sample) */
int sgpu_nco_get(const sgpu_nco *n, size_t channel, uint32_t *theta, uint32_t *delta_theta);
int sgpu_nco_set(sgpu_nco *n, size_t channel, uint32_t theta, uint32_t delta_theta);
uint32_t sgpu_nco_constrain(double theta);                                        /* :176-188 */
/* up != 0: mix_up (:141-145), else mix_down (:147-151); every channel's NCO steps n_in times */
int sgpu_nco_mix_block(sgpu_nco *n, int up, const float *in, size_t n_in, size_t in_stride,
                       float *out, size_t out_stride, sgpu_mem mem, void *stream);

int sgpu_ddc_create(const double *taps, size_t n_taps, sgpu_tapkind kind, size_t n_channels,
                    double scale_re, double scale_im, size_t decimation, sgpu_ddc **out);
int sgpu_ddc_destroy(sgpu_ddc *d);
int sgpu_ddc_clone(const sgpu_ddc *d, sgpu_ddc **out);
sgpu_fir *sgpu_ddc_filter(sgpu_ddc *d); /* the decimator: scale, taps, state through sgpu_fir_* (its history holds MIXED samples) */
sgpu_nco *sgpu_ddc_nco(sgpu_ddc *d);    /* the oscillators: frequency / phase through sgpu_nco_* */
size_t sgpu_ddc_out_len(const sgpu_ddc *d, size_t n_in);
int sgpu_ddc_execute_block(sgpu_ddc *d, const float *in, size_t n_in, size_t in_stride, float *out,
                           size_t out_stride, size_t *n_out, sgpu_mem mem, void *stream);
int sgpu_ddc_write(sgpu_ddc *d, const float *in, size_t n_in, size_t in_stride, sgpu_mem mem,
                   void *stream);        /* mixed, pushed, no output (decim.rs:136-139) */
int sgpu_ddc_reset(sgpu_ddc *d);         /* NCO::reset + the decimator's window and counter */
int sgpu_ddc_last_fused(const sgpu_ddc *d); /* 1: the last call mixed inside the decimator kernel */

/* ---- multi-GPU context -----------------------------------------
 * One caller thread, host buffers, several GPUs behind one filter object: what a caller like
 * the reference's main.rs:39-41 (one Vec in, one Vec out) needs to use the whole box.
 *   - C > 1 channels (independent reference objects): contiguous channel ranges per GPU.
 *   - one FIR / decimating-FIR stream (C == 1): contiguous time segments per call; segment
 *     d > 0 starts where the decimator's counter is 0 (fir/decim.rs:221-228) and is primed with
 *     the T-1 samples in front of it, sliced from the caller's buffer.  Calls too short to
 *     give every GPU 65536 samples use fewer GPUs.  One IIR stream stays on the first GPU.
 *   - every GPU writes its outputs into the caller's one host buffer (that is the gather: no
 *     collective on this path); results and streaming state are those of the single handle.
 * sgpu_ctx_create(0, ..) takes every visible GPU; sgpu_ctx_create_devices takes an explicit
 * list, in which a device may appear more than once (several shards on one GPU). */
int sgpu_ctx_create(int n_gpus, sgpu_ctx **out);
int sgpu_ctx_create_devices(const int *devices, int n, sgpu_ctx **out);
int sgpu_ctx_destroy(sgpu_ctx *ctx);
int sgpu_ctx_devices(const sgpu_ctx *ctx);
int sgpu_ctx_fir_create(sgpu_ctx *ctx, const double *taps, size_t n_taps, sgpu_tapkind kind,
                        size_t n_channels, double scale_re, double scale_im, int is_decimator,
                        size_t decimation, sgpu_sharded **out);       /* fir/mod.rs:79, decim.rs:27 */
int sgpu_ctx_interp_create(sgpu_ctx *ctx, const double *taps, size_t n_taps, sgpu_tapkind kind,
                           size_t n_channels, size_t interpolation, sgpu_sharded **out); /* interp.rs:27 */
int sgpu_ctx_iir_create(sgpu_ctx *ctx, sgpu_iirtype type, const double *ff, size_t n_ff,
                        const double *fb, size_t n_fb, size_t n_channels, sgpu_iirwrap wrap,
                        size_t factor, sgpu_sharded **out);           /* iir/mod.rs:92 */
int sgpu_sharded_destroy(sgpu_sharded *f);
int sgpu_sharded_shards(const sgpu_sharded *f);
int sgpu_sharded_shard_info(const sgpu_sharded *f, int index, int *device, size_t *first_channel,
                            size_t *n_channels);
int sgpu_sharded_last_segments(const sgpu_sharded *f); /* GPUs the last call really used */
size_t sgpu_sharded_out_len(const sgpu_sharded *f, size_t n_in);
int sgpu_sharded_reset(sgpu_sharded *f);
/* Filter::execute_block (filter/mod.rs:14) over host buffers; returns when `out` is complete */
int sgpu_sharded_execute_block(sgpu_sharded *f, const float *in, size_t n_in, size_t in_stride,
                               float *out, size_t out_stride, size_t *n_out);

/* ---- sharding helpers (pure host arithmetic; no collective) --------------------------
 * Channels: contiguous ranges, remainder spread over the first ranks.
 * Stream: contiguous segments whose starts are multiples of `align` (the decimation factor,
 * so every rank's segment starts at decimator phase 0); rank r > 0 must prime its handle
 * with the `halo` = T-1 samples preceding `first` (sgpu_fir_set_state / sgpu_fir_write). */
int sgpu_shard_channels(size_t n_channels, int world, int rank, size_t *first, size_t *count);
int sgpu_shard_stream(size_t n_samples, size_t align, int world, int rank, size_t *first,
                      size_t *count);

#ifdef __cplusplus
}
#endif
#endif /* SOLID_GPU_H */
