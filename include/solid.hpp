// solid.hpp -- C++ host-side mirror of the reference crate's filtering interface over the C ABI
// (include/solid_gpu.h).  The reference is Rust and no Rust toolchain exists in this image, so the
// compiled host layer is C++ (header-only); rust/ holds the equivalent -sys and safe crates as
// source.  Names, argument meaning and error behaviour follow the reference:
//
//   solid::dot_product::{Direction, DotProduct}                     dot_product/mod.rs:31-196
//   solid::filter::Filter<I, O>                                     filter/mod.rs:9-22
//   solid::filter::fir::{FIRErrorCode, FIRError, FIRFilter}         filter/fir/mod.rs:39-316
//   solid::filter::fir::decim::DecimatingFIRFilter                  filter/fir/decim.rs:5-295
//   solid::filter::fir::interp::InterpolatingFIRFilter              filter/fir/interp.rs:6-137
//   solid::filter::fir::pfb::PolyPhaseFilterBank                    filter/fir/pfb.rs:3-90
//   solid::filter::iir::{IIRErrorCode, IIRError, IIRFilterType, IIRFilter}   filter/iir/mod.rs:40-419
//   solid::filter::iir::decim::DecimatingIIRFilter                  filter/iir/decim.rs:5-285
//   solid::filter::iir::interp::InterpolatingIIRFilter              filter/iir/interp.rs:5-273
//
// Rust `new(..) -> Result<Self, Box<dyn Error>>` becomes a constructor that throws the matching
// error type; `#[derive(Clone)]` becomes the copy constructor (deep copy incl. filter state);
// execute() never fails in the reference -- here a CUDA failure throws solid::GpuError (there is no
// CPU fallback to fall back to).  Sample type is cf32 = std::complex<float> (the f32 instantiation
// named by the north star); coefficients are double like the reference's Coef = f64.
#pragma once

#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "solid_gpu.h"
#include "solid_host.hpp"  // Window<T>, CircularBuffer<T>: host-only types of the same path

namespace solid {

using cf32 = std::complex<float>;

struct GpuError : std::runtime_error {
    int status;
    explicit GpuError(int s) : std::runtime_error(std::string(sgpu_status_name(s)) + ": " + sgpu_last_error()), status(s) {}
};

namespace detail {
inline void check(int st) {
    if (st != SGPU_OK) throw GpuError(st);
}
inline const float *fp(const cf32 *p) { return reinterpret_cast<const float *>(p); }
inline float *fp(cf32 *p) { return reinterpret_cast<float *>(p); }
}  // namespace detail

// ------------------------------------------------------------------------------------------------
namespace dot_product {

enum class Direction { FORWARD = SGPU_FORWARD, REVERSE = SGPU_REVERSE };  // dot_product/mod.rs:31-34

class DotProduct {  // dot_product/mod.rs:37-196
   public:
    DotProduct(const std::vector<double> &coefficients, Direction direction) {  // ::new :57
        detail::check(sgpu_dot_create(coefficients.data(), coefficients.size(), SGPU_TAPS_REAL,
                                      (sgpu_direction)direction, &h_));
    }
    ~DotProduct() { sgpu_dot_destroy(h_); }
    DotProduct(const DotProduct &) = delete;
    DotProduct &operator=(const DotProduct &) = delete;
    std::vector<double> coefficents() const {  // [sic] :102 -- stored order
        std::vector<double> v(len());
        if (!v.empty()) detail::check(sgpu_dot_coefficients(h_, v.data()));
        return v;
    }
    size_t len() const { return sgpu_dot_len(h_); }  // :124
    bool is_empty() const { return len() == 0; }     // :141
    cf32 execute(const std::vector<cf32> &samples) const {  // trait Execute, :153-171
        cf32 r;
        detail::check(sgpu_dot_execute(h_, detail::fp(samples.data()), samples.size(), samples.size(), 1,
                                       detail::fp(&r), SGPU_HOST, nullptr));
        return r;
    }

   private:
    sgpu_dot *h_ = nullptr;
};

}  // namespace dot_product

// ------------------------------------------------------------------------------------------------
namespace filter {

template <class I, class O>
struct Filter {  // filter/mod.rs:9-22
    virtual ~Filter() = default;
    virtual std::vector<O> execute(I sample) = 0;
    virtual std::vector<O> execute_block(const std::vector<I> &samples) = 0;
    virtual std::complex<double> frequency_response(double frequency) const = 0;
    virtual double group_delay(double frequency) const = 0;
};

namespace detail_analysis {
// group_delay/mod.rs:51-79; returns false where the reference returns Err
inline bool fir_group_delay(const std::vector<double> &c, double f, double &out) {
    if (c.empty() || f < -0.5 || f > 0.5) return false;
    std::complex<double> t0 = 0, t1 = 0;
    for (size_t i = 0; i < c.size(); ++i) {
        const std::complex<double> rot = std::polar(1.0, f * 2.0 * M_PI * (double)i);
        t0 += c[i] * rot * (double)i;
        t1 += c[i] * rot;
    }
    out = (t0 / t1).real();
    return true;
}
inline std::complex<double> fir_response(const std::vector<double> &c, double f) {
    std::complex<double> o = 0;
    for (size_t i = 0; i < c.size(); ++i) o += c[i] * std::polar(1.0, f * 2.0 * M_PI * (double)i);
    return o;
}
}  // namespace detail_analysis

namespace fir {

enum class FIRErrorCode {  // fir/mod.rs:39-45
    CoefficientsLengthZero,
    DecimationLessThanOne,
    InterpolationLessThanOne,
    NotEnoughFilters
};

struct FIRError : std::runtime_error {  // fir/mod.rs:47-56
    FIRErrorCode code;
    explicit FIRError(FIRErrorCode c) : std::runtime_error("FIR Filter Error"), code(c) {}
};

namespace detail_fir {
inline void check_ctor(int st) {
    switch (st) {
        case SGPU_OK: return;
        case SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO: throw FIRError(FIRErrorCode::CoefficientsLengthZero);
        case SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE: throw FIRError(FIRErrorCode::DecimationLessThanOne);
        case SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE: throw FIRError(FIRErrorCode::InterpolationLessThanOne);
        case SGPU_ERR_FIR_NOT_ENOUGH_FILTERS: throw FIRError(FIRErrorCode::NotEnoughFilters);
        default: throw GpuError(st);
    }
}
inline std::vector<double> flatten(const std::vector<std::vector<double>> &per_channel) {
    std::vector<double> flat;
    for (const auto &row : per_channel) {
        if (row.size() != per_channel[0].size()) throw std::invalid_argument("per-channel tap sets of different lengths");
        flat.insert(flat.end(), row.begin(), row.end());
    }
    return flat;
}
}  // namespace detail_fir

struct PerChannel {};                    // tag: the constructor takes one tap set per channel
inline constexpr PerChannel per_channel{};

class FIRFilter : public Filter<cf32, cf32> {  // fir/mod.rs:58-316
   public:
    FIRFilter(const std::vector<double> &coefficents, double scale, size_t n_channels = 1) {  // ::new :79
        detail_fir::check_ctor(sgpu_fir_create(coefficents.data(), coefficents.size(), SGPU_TAPS_REAL,
                                               n_channels, scale, 0.0, 0, 0, &h_));
    }
    /// one tap set per channel (`per_channel[c]` = the coefficients of reference object c; every reference filter owns
    /// its coefficients, fir/mod.rs:79-88): sgpu_fir_create_per_channel
    FIRFilter(PerChannel, const std::vector<std::vector<double>> &taps, double scale) {
        const std::vector<double> flat = detail_fir::flatten(taps);
        detail_fir::check_ctor(sgpu_fir_create_per_channel(flat.data(), taps.empty() ? 0 : taps[0].size(), SGPU_TAPS_REAL,
                                                           taps.size(), scale, 0.0, 0, 0, &h_));
    }
    FIRFilter(const FIRFilter &o) { detail::check(sgpu_fir_clone(o.h_, &h_)); }  // #[derive(Clone)]
    FIRFilter &operator=(const FIRFilter &) = delete;
    ~FIRFilter() override { sgpu_fir_destroy(h_); }

    void set_scale(double scale) { detail::check(sgpu_fir_set_scale(h_, scale, 0.0)); }  // :106
    double get_scale() const {                                                            // :124
        double re = 0, im = 0;
        detail::check(sgpu_fir_get_scale(h_, &re, &im));
        return re;
    }
    size_t len() const { return sgpu_fir_len(h_); }  // :142
    bool is_empty() const { return len() == 0; }     // :158
    std::vector<double> coefficients() const {       // :176 -- stored (reversed) order
        std::vector<double> v(len());
        detail::check(sgpu_fir_coefficients(h_, v.data()));
        return v;
    }
    size_t channels() const { return sgpu_fir_channels(h_); }
    /// true when the last execute_block ran on the tcgen05 tensor-core kernel (long filters; csrc/fir_tc.cu)
    bool last_path_tensor() const { return sgpu_fir_last_path(h_) == 1; }

    std::vector<cf32> execute(cf32 sample) override { return execute_block(std::vector<cf32>{sample}); }  // :209
    // :235 -- channel-major [channels][n]; one channel by default
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) override {
        const size_t C = channels(), n = samples.size() / C;
        std::vector<cf32> out(C * sgpu_fir_out_len(h_, n));
        size_t n_out = 0;
        const size_t cap = out.size() / C;
        detail::check(sgpu_fir_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()),
                                             cap ? cap : 1, &n_out, SGPU_HOST, nullptr));
        return out;
    }
    // hot path: device pointers, asynchronous on `stream`
    size_t execute_block_device(const cf32 *d_in, size_t n, size_t in_stride, cf32 *d_out, size_t out_stride,
                                void *stream) {
        size_t n_out = 0;
        detail::check(sgpu_fir_execute_block(h_, detail::fp(d_in), n, in_stride, detail::fp(d_out), out_stride,
                                             &n_out, SGPU_DEVICE, stream));
        return n_out;
    }
    std::complex<double> frequency_response(double f) const override {  // :263-273
        return get_scale() * detail_analysis::fir_response(coefficients(), f);
    }
    double group_delay(double f) const override {  // :293-303
        double d = 0.0;
        return detail_analysis::fir_group_delay(coefficients(), f, d) ? d : 0.0;
    }
    sgpu_fir *handle() { return h_; }

   protected:
    FIRFilter() = default;
    sgpu_fir *h_ = nullptr;
};

namespace decim {
class DecimatingFIRFilter : public FIRFilter {  // fir/decim.rs:5-295
   public:
    DecimatingFIRFilter(const std::vector<double> &coefficents, double scale, size_t decimation,
                        size_t n_channels = 1) {  // ::new :27
        detail_fir::check_ctor(sgpu_fir_create(coefficents.data(), coefficents.size(), SGPU_TAPS_REAL,
                                               n_channels, scale, 0.0, 1, decimation, &h_));
    }
    DecimatingFIRFilter(PerChannel, const std::vector<std::vector<double>> &taps, double scale, size_t decimation) {
        const std::vector<double> flat = detail_fir::flatten(taps);
        detail_fir::check_ctor(sgpu_fir_create_per_channel(flat.data(), taps.empty() ? 0 : taps[0].size(), SGPU_TAPS_REAL,
                                                           taps.size(), scale, 0.0, 1, decimation, &h_));
    }
    DecimatingFIRFilter(const DecimatingFIRFilter &o) : FIRFilter() { detail::check(sgpu_fir_clone(o.h_, &h_)); }
    size_t get_decimation() const { return sgpu_fir_decimation(h_); }  // :96
    void push(cf32 sample) { write(std::vector<cf32>{sample}); }       // :115
    void write(const std::vector<cf32> &samples) {                     // :136
        const size_t C = channels(), n = samples.size() / C;
        detail::check(sgpu_fir_write(h_, detail::fp(samples.data()), n, n, SGPU_HOST, nullptr));
    }
};
}  // namespace decim

namespace pfb {
class PolyPhaseFilterBank {  // fir/pfb.rs:3-90
   public:
    PolyPhaseFilterBank(const std::vector<double> &coefficients, size_t filters, double scale,
                        size_t n_channels = 1) {  // ::new :24
        detail_fir::check_ctor(sgpu_pfb_create(coefficients.data(), coefficients.size(), SGPU_TAPS_REAL,
                                               n_channels, filters, scale, 0.0, &h_));
    }
    PolyPhaseFilterBank(const PolyPhaseFilterBank &o) { detail::check(sgpu_interp_clone(o.h_, &h_)); }
    PolyPhaseFilterBank &operator=(const PolyPhaseFilterBank &) = delete;
    ~PolyPhaseFilterBank() { sgpu_interp_destroy(h_); }
    void set_scale(double s) { detail::check(sgpu_interp_set_scale(h_, s, 0.0)); }  // :52
    double get_scale() const {                                                       // :57
        double re = 0, im = 0;
        detail::check(sgpu_interp_get_scale(h_, &re, &im));
        return re;
    }
    size_t len() const { return sgpu_interp_interpolation(h_); }  // :62
    bool is_empty() const { return len() == 0; }                  // :67
    std::vector<std::vector<double>> coefficents() const {        // :71
        const size_t L = len(), S = sgpu_interp_sub_len(h_);
        std::vector<double> flat(L * S);
        detail::check(sgpu_interp_coefficients(h_, flat.data()));
        std::vector<std::vector<double>> v(L);
        for (size_t p = 0; p < L; ++p) v[p].assign(flat.begin() + p * S, flat.begin() + (p + 1) * S);
        return v;
    }
    void reset() { detail::check(sgpu_interp_reset(h_)); }  // :76
    void push(cf32 sample) {                                 // :81
        detail::check(sgpu_interp_push(h_, detail::fp(&sample), 1, 1, SGPU_HOST, nullptr));
    }
    cf32 execute(size_t index) {  // :85 -- NO scale applied
        cf32 r;
        detail::check(sgpu_interp_execute_phase(h_, index, detail::fp(&r), SGPU_HOST, nullptr));
        return r;
    }

   private:
    sgpu_interp *h_ = nullptr;
};
}  // namespace pfb

namespace interp {
class InterpolatingFIRFilter : public Filter<cf32, cf32> {  // fir/interp.rs:6-137
   public:
    InterpolatingFIRFilter(const std::vector<double> &coefficents, size_t interpolation,
                           size_t n_channels = 1) {  // ::new :27
        detail_fir::check_ctor(sgpu_interp_create(coefficents.data(), coefficents.size(), SGPU_TAPS_REAL,
                                                  n_channels, interpolation, &h_));
    }
    InterpolatingFIRFilter(const InterpolatingFIRFilter &o) { detail::check(sgpu_interp_clone(o.h_, &h_)); }
    InterpolatingFIRFilter &operator=(const InterpolatingFIRFilter &) = delete;
    ~InterpolatingFIRFilter() override { sgpu_interp_destroy(h_); }
    void set_scale(double s) { detail::check(sgpu_interp_set_scale(h_, s, 0.0)); }  // :57 (never applied)
    double get_scale() const {                                                       // :62
        double re = 0, im = 0;
        detail::check(sgpu_interp_get_scale(h_, &re, &im));
        return re;
    }
    size_t len() const { return sgpu_interp_interpolation(h_); }  // :67
    bool is_empty() const { return len() == 0; }                  // :72
    std::vector<double> coefficents() const {                     // :77
        std::vector<double> v(len() * sgpu_interp_sub_len(h_));
        detail::check(sgpu_interp_coefficients(h_, v.data()));
        return v;
    }
    size_t interpolation() const { return sgpu_interp_interpolation(h_); }  // :82
    size_t channels() const { return sgpu_interp_channels(h_); }
    bool last_path_tensor() const { return sgpu_interp_last_path(h_) == 1; }
    std::vector<cf32> execute(cf32 sample) override { return execute_block(std::vector<cf32>{sample}); }  // :93
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) override {                          // :102
        const size_t C = channels(), n = samples.size() / C, L = interpolation();
        std::vector<cf32> out(C * n * L);
        size_t n_out = 0;
        detail::check(sgpu_interp_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()),
                                                (n * L) > 0 ? n * L : 1, &n_out, SGPU_HOST, nullptr));
        return out;
    }
    size_t execute_block_device(const cf32 *d_in, size_t n, size_t in_stride, cf32 *d_out, size_t out_stride,
                                void *stream) {
        size_t n_out = 0;
        detail::check(sgpu_interp_execute_block(h_, detail::fp(d_in), n, in_stride, detail::fp(d_out), out_stride,
                                                &n_out, SGPU_DEVICE, stream));
        return n_out;
    }
    std::complex<double> frequency_response(double f) const override {  // :113-124
        return get_scale() * detail_analysis::fir_response(coefficents(), f);
    }
    double group_delay(double f) const override {  // :126-137
        double d = 0.0;
        return detail_analysis::fir_group_delay(coefficents(), f, d) ? d : 0.0;
    }

   private:
    sgpu_interp *h_ = nullptr;
};
}  // namespace interp
}  // namespace fir

// ------------------------------------------------------------------------------------------------
namespace iir {

enum class IIRErrorCode {  // iir/mod.rs:40-49
    NumeratorLengthZero,
    DenominatorLengthZero,
    SecondOrderSectionSizeZero,
    SecondOrderSectionSizeMismatch,
    SecondOrderSectionSizeNotMultpleOf3,
    DecimationLessThanOne,
    InterpolationLessThanOne
};
struct IIRError : std::runtime_error {  // iir/mod.rs:51-60
    IIRErrorCode code;
    explicit IIRError(IIRErrorCode c) : std::runtime_error("IIR Filter Error"), code(c) {}
};
enum class IIRFilterType { Normal = SGPU_IIR_NORMAL, SecondOrder = SGPU_IIR_SECOND_ORDER };  // :62-66

namespace detail_iir {
inline void check_ctor(int st) {
    switch (st) {
        case SGPU_OK: return;
        case SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO: throw IIRError(IIRErrorCode::NumeratorLengthZero);
        case SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO: throw IIRError(IIRErrorCode::DenominatorLengthZero);
        case SGPU_ERR_IIR_SOS_SIZE_ZERO: throw IIRError(IIRErrorCode::SecondOrderSectionSizeZero);
        case SGPU_ERR_IIR_SOS_SIZE_MISMATCH: throw IIRError(IIRErrorCode::SecondOrderSectionSizeMismatch);
        case SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3: throw IIRError(IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3);
        case SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE: throw IIRError(IIRErrorCode::DecimationLessThanOne);
        case SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE: throw IIRError(IIRErrorCode::InterpolationLessThanOne);
        default: throw GpuError(st);
    }
}
}  // namespace detail_iir

class IIRFilter : public Filter<cf32, cf32> {  // iir/mod.rs:68-419
   public:
    IIRFilter(const std::vector<double> &feed_forward, const std::vector<double> &feed_back, IIRFilterType iirtype,
              size_t n_channels = 1)  // ::new :92
        : IIRFilter(feed_forward, feed_back, iirtype, n_channels, SGPU_IIR_PLAIN, 0) {}
    IIRFilter(const IIRFilter &o) { detail::check(sgpu_iir_clone(o.h_, &h_)); }
    IIRFilter &operator=(const IIRFilter &) = delete;
    ~IIRFilter() override { sgpu_iir_destroy(h_); }

    std::vector<double> numerator_coefs() const { return coefs(sgpu_iir_numerator_coefs); }      // :182
    std::vector<double> denominator_coefs() const { return coefs(sgpu_iir_denominator_coefs); }  // :202
    size_t second_order_filters_len() const { return sgpu_iir_sections(h_); }                    // :222
    IIRFilterType iir_type() const { return (IIRFilterType)sgpu_iir_type(h_); }                  // :239
    size_t channels() const { return sgpu_iir_channels(h_); }
    void set_mode(int mode) { detail::check(sgpu_iir_set_mode(h_, mode)); }
    std::size_t decay_length() const {
        std::size_t n = 0;
        detail::check(sgpu_iir_decay_length(h_, &n));
        return n;
    }

    std::vector<cf32> execute(cf32 sample) override { return execute_block(std::vector<cf32>{sample}); }  // :270
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) override {                          // :310
        const size_t C = channels(), n = samples.size() / C;
        const size_t cap = sgpu_iir_out_len(h_, n);
        std::vector<cf32> out(C * cap);
        size_t n_out = 0;
        detail::check(sgpu_iir_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()),
                                             cap ? cap : 1, &n_out, SGPU_HOST, nullptr));
        return out;
    }
    size_t execute_block_device(const cf32 *d_in, size_t n, size_t in_stride, cf32 *d_out, size_t out_stride,
                                void *stream) {
        size_t n_out = 0;
        detail::check(sgpu_iir_execute_block(h_, detail::fp(d_in), n, in_stride, detail::fp(d_out), out_stride,
                                             &n_out, SGPU_DEVICE, stream));
        return n_out;
    }
    // iir/mod.rs:336-373: the SecondOrder branch multiplies into a zero product, so the reference
    // returns 0 there (its own doc-test asserts it, :328-334)
    std::complex<double> frequency_response(double f) const override {
        if (iir_type() == IIRFilterType::SecondOrder) return {0.0, 0.0};
        return detail_analysis::fir_response(numerator_coefs(), f) /
               detail_analysis::fir_response(denominator_coefs(), f);
    }
    double group_delay(double) const override { return 0.0; }  // host analysis lives in the Python mirror

   protected:
    IIRFilter(const std::vector<double> &ff, const std::vector<double> &fb, IIRFilterType t, size_t n_channels,
              sgpu_iirwrap wrap, size_t factor) {
        detail_iir::check_ctor(sgpu_iir_create((sgpu_iirtype)t, ff.data(), ff.size(), fb.data(), fb.size(),
                                               n_channels, wrap, factor, &h_));
    }
    template <class F>
    std::vector<double> coefs(F fn) const {
        size_t n = 0;
        detail::check(fn(h_, nullptr, &n));
        std::vector<double> v(n);
        detail::check(fn(h_, v.data(), &n));
        return v;
    }
    sgpu_iir *h_ = nullptr;
};

namespace decim {
class DecimatingIIRFilter : public IIRFilter {  // iir/decim.rs:5-285
   public:
    DecimatingIIRFilter(const std::vector<double> &ff, const std::vector<double> &fb, IIRFilterType t,
                        size_t decimation, size_t n_channels = 1)  // ::new :30
        : IIRFilter(ff, fb, t, n_channels, SGPU_IIR_DECIMATING, decimation), m_(decimation) {}
    size_t get_decimation() const { return m_; }  // :64

   private:
    size_t m_;
};
}  // namespace decim
namespace interp {
class InterpolatingIIRFilter : public IIRFilter {  // iir/interp.rs:5-273
   public:
    InterpolatingIIRFilter(const std::vector<double> &ff, const std::vector<double> &fb, IIRFilterType t,
                           size_t interpolation, size_t n_channels = 1)  // ::new :29
        : IIRFilter(ff, fb, t, n_channels, SGPU_IIR_INTERPOLATING, interpolation), l_(interpolation) {}
    size_t get_interpolation() const { return l_; }  // :62

   private:
    size_t l_;
};
}  // namespace interp
}  // namespace iir

// ------------------------------------------------------------------------------------------------
namespace auto_correlator {

// AutoCorrelator<C> -- filter/auto_correlator/mod.rs:24-216.  One object = `n_channels` independent
// correlators (one reference object per channel), channel-major buffers.
class AutoCorrelator {
   public:
    AutoCorrelator(size_t window_size, size_t delay, size_t n_channels = 1) {  // ::new :51
        detail::check(sgpu_autocorr_create(window_size, delay, n_channels, &h_));
    }
    AutoCorrelator(const AutoCorrelator &o) { detail::check(sgpu_autocorr_clone(o.h_, &h_)); }
    AutoCorrelator &operator=(const AutoCorrelator &) = delete;
    ~AutoCorrelator() { sgpu_autocorr_destroy(h_); }

    size_t window_size() const { return sgpu_autocorr_window_size(h_); }
    size_t delay() const { return sgpu_autocorr_delay(h_); }
    size_t channels() const { return sgpu_autocorr_channels(h_); }
    void reset() { detail::check(sgpu_autocorr_reset(h_)); }  // :76
    void push(cf32 sample) { write(std::vector<cf32>(channels(), sample)); }  // :99 (same sample on every channel)
    void write(const std::vector<cf32> &samples) {  // :130
        const size_t n = samples.size() / channels();
        detail::check(sgpu_autocorr_write(h_, detail::fp(samples.data()), n, n, SGPU_HOST, nullptr));
    }
    std::vector<std::complex<double>> execute() {  // :165, one value per channel
        std::vector<std::complex<double>> out(channels());
        detail::check(sgpu_autocorr_execute(h_, reinterpret_cast<double *>(out.data())));
        return out;
    }
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) {  // :184
        const size_t C = channels(), n = samples.size() / C;
        std::vector<cf32> out(C * n);
        size_t n_out = 0;
        detail::check(sgpu_autocorr_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()),
                                                  n ? n : 1, &n_out, SGPU_HOST, nullptr));
        return out;
    }
    size_t execute_block_device(const cf32 *d_in, size_t n, size_t in_stride, cf32 *d_out, size_t out_stride,
                                void *stream) {
        size_t n_out = 0;
        detail::check(sgpu_autocorr_execute_block(h_, detail::fp(d_in), n, in_stride, detail::fp(d_out), out_stride,
                                                  &n_out, SGPU_DEVICE, stream));
        return n_out;
    }
    std::vector<double> get_energy() {  // :214, one value per channel
        std::vector<double> out(channels());
        detail::check(sgpu_autocorr_get_energy(h_, out.data()));
        return out;
    }

   private:
    sgpu_autocorr *h_ = nullptr;
};
}  // namespace auto_correlator

// ------------------------------------------------------------------------------------------------
namespace firdes {

enum class FirdesErrorCode { Bandwidth, StopBandLevel, Mu };  // firdes/mod.rs:17-25 (those firdes_kaiser returns)

struct FirdesError : std::runtime_error {
    FirdesErrorCode code;
    explicit FirdesError(FirdesErrorCode c) : std::runtime_error("Firdes Error"), code(c) {}
};

// firdes_kaiser -- firdes/mod.rs:278-305, computed on the GPU (sgpu_firdes_kaiser)
inline std::vector<double> firdes_kaiser(size_t filter_length, double cutoff_frequency, double stop_band_attenuation,
                                         double fractional_sample_offset) {
    std::vector<double> h(filter_length);
    const int st = sgpu_firdes_kaiser(filter_length, &cutoff_frequency, &stop_band_attenuation, &fractional_sample_offset, 1,
                                      h.data(), SGPU_HOST, nullptr);
    switch (st) {
        case SGPU_OK: return h;
        case SGPU_ERR_FIRDES_BANDWIDTH: throw FirdesError(FirdesErrorCode::Bandwidth);
        case SGPU_ERR_FIRDES_STOP_BAND_LEVEL: throw FirdesError(FirdesErrorCode::StopBandLevel);
        case SGPU_ERR_FIRDES_MU: throw FirdesError(FirdesErrorCode::Mu);
        default: throw GpuError(st);
    }
}

}  // namespace firdes
// ------------------------------------------------------------------------------------------------
namespace ddc {

// NCO mix-down feeding a DecimatingFIRFilter (nco/mod.rs:147-151 -> fir/decim.rs:221-256), the mix fused into the
// decimator's tile on the hot shapes: sgpu_ddc_*.  Frequency / phase through nco().
class DigitalDownConverter {
   public:
    DigitalDownConverter(const std::vector<double> &coefficents, double scale, size_t decimation, double frequency,
                         size_t n_channels = 1) {
        fir::detail_fir::check_ctor(sgpu_ddc_create(coefficents.data(), coefficents.size(), SGPU_TAPS_REAL, n_channels, scale,
                                                    0.0, decimation, &h_));
        detail::check(sgpu_nco_set_frequency(sgpu_ddc_nco(h_), SGPU_ALL_CHANNELS, frequency));
    }
    DigitalDownConverter(const DigitalDownConverter &o) { detail::check(sgpu_ddc_clone(o.h_, &h_)); }
    DigitalDownConverter &operator=(const DigitalDownConverter &) = delete;
    ~DigitalDownConverter() { sgpu_ddc_destroy(h_); }
    sgpu_nco *nco() { return sgpu_ddc_nco(h_); }
    size_t channels() const { return sgpu_fir_channels(sgpu_ddc_filter(h_)); }
    bool last_fused() const { return sgpu_ddc_last_fused(h_) == 1; }
    void reset() { detail::check(sgpu_ddc_reset(h_)); }
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) {  // channel-major [channels][n]
        const size_t C = channels(), n = samples.size() / C;
        const size_t cap = sgpu_ddc_out_len(h_, n);
        std::vector<cf32> out(C * cap);
        size_t n_out = 0;
        detail::check(sgpu_ddc_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()), cap ? cap : 1, &n_out,
                                             SGPU_HOST, nullptr));
        return out;
    }

   private:
    sgpu_ddc *h_ = nullptr;
};

}  // namespace ddc
}  // namespace filter

// ------------------------------------------------------------------------------------------------
namespace nco {

// NCO -- nco/mod.rs:26-188: `n_channels` oscillators (32-bit phase and step, 1024-entry sine table)
class NCO {
   public:
    explicit NCO(size_t n_channels = 1) { detail::check(sgpu_nco_create(n_channels, &h_)); }  // ::new :36-50
    NCO(const NCO &o) { detail::check(sgpu_nco_clone(o.h_, &h_)); }
    NCO &operator=(const NCO &) = delete;
    ~NCO() { sgpu_nco_destroy(h_); }
    void reset() { detail::check(sgpu_nco_reset(h_)); }                                                          // :53
    void set_frequency(double delta_theta, size_t ch = SGPU_ALL_CHANNELS) { detail::check(sgpu_nco_set_frequency(h_, ch, delta_theta)); }  // :59
    void adjust_frequency(double dt, size_t ch = SGPU_ALL_CHANNELS) { detail::check(sgpu_nco_adjust_frequency(h_, ch, dt)); }          // :64
    void set_phase(double phi, size_t ch = SGPU_ALL_CHANNELS) { detail::check(sgpu_nco_set_phase(h_, ch, phi)); }                      // :79
    void adjust_phase(double dphi, size_t ch = SGPU_ALL_CHANNELS) { detail::check(sgpu_nco_adjust_phase(h_, ch, dphi)); }              // :84
    void step(uint64_t count = 1) { detail::check(sgpu_nco_step(h_, count)); }                                   // :93
    size_t channels() const { return sgpu_nco_channels(h_); }
    // the per-sample loop the reference's mix_up_block / mix_down_block intend (:141-172); channel-major [channels][n]
    std::vector<cf32> mix_block(const std::vector<cf32> &samples, bool up) {
        const size_t C = channels(), n = samples.size() / C;
        std::vector<cf32> out(samples.size());
        detail::check(sgpu_nco_mix_block(h_, up ? 1 : 0, detail::fp(samples.data()), n, n, detail::fp(out.data()), n ? n : 1,
                                         SGPU_HOST, nullptr));
        return out;
    }
    std::vector<cf32> mix_up_block(const std::vector<cf32> &x) { return mix_block(x, true); }
    std::vector<cf32> mix_down_block(const std::vector<cf32> &x) { return mix_block(x, false); }

   private:
    sgpu_nco *h_ = nullptr;
};

}  // namespace nco

// ------------------------------------------------------------------------------------------------
namespace multi_gpu {

// Every GPU of the box behind ONE filter object and one caller thread (host buffers): sgpu_ctx_* / sgpu_sharded_*.
class Context {
   public:
    explicit Context(int n_gpus = 0) { detail::check(sgpu_ctx_create(n_gpus, &h_)); }          // 0 = every visible GPU
    explicit Context(const std::vector<int> &devices) {                                        // a device may repeat
        detail::check(sgpu_ctx_create_devices(devices.data(), (int)devices.size(), &h_));
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ~Context() { sgpu_ctx_destroy(h_); }
    int devices() const { return sgpu_ctx_devices(h_); }
    sgpu_ctx *handle() { return h_; }

   private:
    sgpu_ctx *h_ = nullptr;
};

class ShardedFilter {
   public:
    // FIRFilter / DecimatingFIRFilter (decimation = 0: plain FIR) over the context's GPUs
    static ShardedFilter fir(Context &ctx, const std::vector<double> &taps, double scale, size_t decimation, size_t n_channels) {
        ShardedFilter f;
        filter::fir::detail_fir::check_ctor(sgpu_ctx_fir_create(ctx.handle(), taps.data(), taps.size(), SGPU_TAPS_REAL, n_channels,
                                                                scale, 0.0, decimation ? 1 : 0, decimation, &f.h_));
        f.C_ = n_channels;
        return f;
    }
    static ShardedFilter interp(Context &ctx, const std::vector<double> &taps, size_t interpolation, size_t n_channels) {
        ShardedFilter f;
        filter::fir::detail_fir::check_ctor(sgpu_ctx_interp_create(ctx.handle(), taps.data(), taps.size(), SGPU_TAPS_REAL, n_channels,
                                                                   interpolation, &f.h_));
        f.C_ = n_channels;
        return f;
    }
    ShardedFilter(ShardedFilter &&o) noexcept : h_(o.h_), C_(o.C_) { o.h_ = nullptr; }
    ShardedFilter(const ShardedFilter &) = delete;
    ShardedFilter &operator=(const ShardedFilter &) = delete;
    ~ShardedFilter() { if (h_) sgpu_sharded_destroy(h_); }
    int shards() const { return sgpu_sharded_shards(h_); }
    int last_segments() const { return sgpu_sharded_last_segments(h_); }
    void reset() { detail::check(sgpu_sharded_reset(h_)); }
    std::vector<cf32> execute_block(const std::vector<cf32> &samples) {  // Filter::execute_block, channel-major
        const size_t n = samples.size() / C_;
        const size_t cap = sgpu_sharded_out_len(h_, n);
        std::vector<cf32> out(C_ * cap);
        size_t n_out = 0;
        detail::check(sgpu_sharded_execute_block(h_, detail::fp(samples.data()), n, n, detail::fp(out.data()), cap ? cap : 1, &n_out));
        return out;
    }

   private:
    ShardedFilter() = default;
    sgpu_sharded *h_ = nullptr;
    size_t C_ = 1;
};

}  // namespace multi_gpu
}  // namespace solid
