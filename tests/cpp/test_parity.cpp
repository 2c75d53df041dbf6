// C++ parity tests: the compiled host mirror (include/solid.hpp) over the C ABI versus the CPU
// oracle (oracle/solid_oracle.c).  The cases follow the reference's doc-tests (cited per case) and
// add seeded random comparisons.  Exit code 0 = all passed.  Needs a B200; run by tests/test_cpp_gpu.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "solid.hpp"

extern "C" {
// oracle entry points (oracle/solid_oracle.c)
size_t so_fir_fast(const double *coefs, size_t T, int coef_complex, double sre, double sim, size_t decimation,
                   size_t count0, const double *hist, const double *x, size_t n, double *out);
size_t so_firinterp_fast(const double *coefs, size_t T, int coef_complex, size_t L, const double *hist,
                         const double *x, size_t n, double *out);
void so_sos_cascade_fast(const double *ff, const double *fb, size_t nsec, double *state, const double *x, size_t n,
                         double *out);
void so_autocorr_fast(size_t window_size, size_t delay, const double *hist, size_t nhist, const double *x, size_t n,
                      double *out);
int so_firdes_kaiser(size_t len, double fc, double as, double mu, double *h);
int so_pll_active_lag(double w, double zeta, double k, double *num, double *den);
typedef struct so_nco so_nco;
so_nco *so_nco_new(void);
void so_nco_free(so_nco *n);
void so_nco_set_frequency(so_nco *n, double dt);
void so_nco_set_phase(so_nco *n, double phi);
void so_nco_mix_block(so_nco *n, int up, const double *x, size_t len, double *out);
}

using namespace solid;
using solid::filter::fir::FIRFilter;
using solid::filter::fir::decim::DecimatingFIRFilter;
using solid::filter::fir::interp::InterpolatingFIRFilter;
using solid::filter::iir::IIRFilter;
using solid::filter::iir::IIRFilterType;

static int failures = 0;
#define EXPECT(cond, what)                                         \
    do {                                                           \
        if (!(cond)) {                                             \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, what); \
            ++failures;                                            \
        }                                                          \
    } while (0)

static std::vector<double> widen(const std::vector<cf32> &x) {
    std::vector<double> d(2 * x.size());
    for (size_t i = 0; i < x.size(); ++i) { d[2 * i] = x[i].real(); d[2 * i + 1] = x[i].imag(); }
    return d;
}
static double nerr(const std::vector<cf32> &got, const std::vector<double> &ref, size_t n) {
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) {
        const double dr = got[i].real() - ref[2 * i], di = got[i].imag() - ref[2 * i + 1];
        num = std::max(num, std::hypot(dr, di));
        den = std::max(den, std::hypot(ref[2 * i], ref[2 * i + 1]));
    }
    return den > 0 ? num / den : num;
}
static std::vector<cf32> rand_cf32(std::mt19937 &g, size_t n) {
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<cf32> x(n);
    for (auto &v : x) v = cf32(u(g), u(g));
    return x;
}
static std::vector<double> f32round(std::vector<double> h) {
    for (auto &v : h) v = (double)(float)v;
    return h;
}

int main() {
    const double TOL = 1e-5;
    // fir/mod.rs:200-206, 226-232
    {
        FIRFilter filter({1.0, 2.0, 3.0, 4.0, 5.0}, 1.0);
        std::vector<cf32> window = {{2.02f, 0}, {4.04f, 0}, {1.02f, 0}, {0.23f, 0}, {9.19f, 0}};
        FIRFilter one(filter);  // Clone
        auto first = one.execute(window[0]);
        EXPECT(first.size() == 1 && std::abs(first[0] - cf32(10.1f, 0)) < 1e-4f, "fir execute 10.1");
        auto out = filter.execute_block(window);
        EXPECT(out.size() == 5 && std::abs(out[4] - cf32(60.03f, 0)) < 6e-4f, "fir execute_block 60.03");
        EXPECT(filter.len() == 5 && !filter.is_empty() && filter.get_scale() == 1.0, "fir accessors");
        EXPECT(filter.coefficients() == std::vector<double>({5, 4, 3, 2, 1}), "fir stored (reversed) order");
    }
    // fir/decim.rs:213-219, 242-248
    {
        DecimatingFIRFilter f({1.0, 2.0, 3.0, 4.0, 5.0}, 1.0, 2);
        auto a = f.execute({2.02f, 0});
        auto b = f.execute({4.04f, 0});
        EXPECT(a.empty() && b.size() == 1 && std::abs(b[0] - cf32(28.28f, 0)) < 3e-4f, "decim execute 28.28");
        DecimatingFIRFilter g({1.0, 2.0, 3.0, 4.0, 5.0}, 1.0, 2);
        auto out = g.execute_block({{2.02f, 0}, {4.04f, 0}, {1.02f, 0}, {0.23f, 0}});
        EXPECT(out.size() == 2 && std::abs(out[1] - cf32(21.39f, 0)) < 3e-4f, "decim execute_block 21.39");
        EXPECT(g.get_decimation() == 2, "get_decimation");
    }
    // construction errors -- fir/mod.rs:80-82, decim.rs:28-31, interp.rs:28-31
    {
        using namespace solid::filter::fir;
        try { FIRFilter f({}, 1.0); EXPECT(false, "empty taps must throw"); }
        catch (const FIRError &e) { EXPECT(e.code == FIRErrorCode::CoefficientsLengthZero, "CoefficientsLengthZero"); }
        try { DecimatingFIRFilter f({1.0}, 1.0, 0); EXPECT(false, "decimation 0 must throw"); }
        catch (const FIRError &e) { EXPECT(e.code == FIRErrorCode::DecimationLessThanOne, "DecimationLessThanOne"); }
        try { InterpolatingFIRFilter f({1.0}, 0); EXPECT(false, "interpolation 0 must throw"); }
        catch (const FIRError &e) { EXPECT(e.code == FIRErrorCode::InterpolationLessThanOne, "InterpolationLessThanOne"); }
        try { IIRFilter f({1, 2, 3, 4}, {1, 2, 3, 4}, IIRFilterType::SecondOrder); EXPECT(false, "sos size must throw"); }
        catch (const solid::filter::iir::IIRError &e) {
            EXPECT(e.code == solid::filter::iir::IIRErrorCode::SecondOrderSectionSizeNotMultpleOf3, "NotMultpleOf3");
        }
    }
    std::mt19937 gen(42);
    // random FIR / decimator / interpolator versus the oracle
    {
        std::vector<double> h(512);
        so_firdes_kaiser(512, 0.1, 80.0, 0.0, h.data());
        h = f32round(h);
        auto x = rand_cf32(gen, 50000);
        auto xd = widen(x);
        std::vector<double> ref(2 * x.size() + 2);
        FIRFilter f(h, 1.0);
        auto y = f.execute_block(x);
        size_t n = so_fir_fast(h.data(), h.size(), 0, 1.0, 0.0, 0, 0, nullptr, xd.data(), x.size(), ref.data());
        EXPECT(n == y.size() && nerr(y, ref, n) <= TOL, "FIR 512 taps vs oracle");
        DecimatingFIRFilter d(h, 0.5, 8);
        auto yd = d.execute_block(x);
        n = so_fir_fast(h.data(), h.size(), 0, 0.5, 0.0, 8, 0, nullptr, xd.data(), x.size(), ref.data());
        EXPECT(n == yd.size() && nerr(yd, ref, n) <= TOL, "decimator M=8 vs oracle");
        std::vector<double> hi(h.begin(), h.begin() + 128);
        InterpolatingFIRFilter ip(hi, 4);
        std::vector<cf32> xs(x.begin(), x.begin() + 5000);
        auto yi = ip.execute_block(xs);
        std::vector<double> refi(2 * 4 * xs.size() + 2);
        n = so_firinterp_fast(hi.data(), hi.size(), 0, 4, nullptr, xd.data(), xs.size(), refi.data());
        EXPECT(n == yi.size() && nerr(yi, refi, n) <= TOL, "interpolator L=4 vs oracle");
        EXPECT(ip.interpolation() == 4 && ip.len() == 4, "interp accessors");
    }
    // iir/mod.rs:302-307 (5-sample golden; double pole at z~1, f32 coefficient rounding -> 1e-4)
    {
        double num[3], den[3];
        so_pll_active_lag(0.02, 1.0 / std::sqrt(2.0), 1000.0, num, den);
        IIRFilter f({num[0], num[1], num[2]}, {den[0], den[1], den[2]}, IIRFilterType::SecondOrder);
        auto y = f.execute_block({{1, 0}, {0, 0}, {1, 0}, {0, 0}, {1, 0}});
        const double exp[5] = {0.05816769596076701, 0.119535296293297, 0.18410279587774706, 0.2518701895942824,
                               0.32283747232307686};
        bool ok = y.size() == 5;
        for (int i = 0; ok && i < 5; ++i) ok = std::abs(y[i].real() - exp[i]) <= 1e-4 * exp[4];
        EXPECT(ok, "IIR SOS golden");
        EXPECT(f.second_order_filters_len() == 1 && f.iir_type() == IIRFilterType::SecondOrder, "iir accessors");
    }
    // 8-section cascade, 64 channels, versus the oracle
    {
        std::vector<double> ff, fb;
        const double rr[8] = {0.50, 0.60, 0.70, 0.78, 0.84, 0.88, 0.92, 0.95};
        const double th[8] = {0.10, 0.14, 0.18, 0.22, 0.26, 0.30, 0.34, 0.38};
        for (int i = 0; i < 8; ++i) {
            const double a1 = -2.0 * rr[i] * std::cos(M_PI * th[i]), a2 = rr[i] * rr[i], g = (1.0 + a1 + a2) / 4.0;
            ff.insert(ff.end(), {g, 2 * g, g});
            fb.insert(fb.end(), {1.0, a1, a2});
        }
        ff = f32round(ff);
        fb = f32round(fb);
        const size_t C = 64, n = 3000;
        auto x = rand_cf32(gen, C * n);
        IIRFilter f(ff, fb, IIRFilterType::SecondOrder, C);
        auto y = f.execute_block(x);
        double worst = 0;
        for (size_t c : {size_t(0), size_t(31), size_t(63)}) {
            std::vector<cf32> xc(x.begin() + c * n, x.begin() + (c + 1) * n), yc(y.begin() + c * n, y.begin() + (c + 1) * n);
            auto xd = widen(xc);
            std::vector<double> st(2 * 2 * 8, 0.0), ref(2 * n);
            so_sos_cascade_fast(ff.data(), fb.data(), 8, st.data(), xd.data(), n, ref.data());
            worst = std::max(worst, nerr(yc, ref, n));
        }
        EXPECT(worst <= TOL, "IIR 8 sections x 64 channels vs oracle");
    }
    // AutoCorrelator: reference doc-test (auto_correlator/mod.rs:199-211) and a random stream vs the oracle
    {
        using solid::filter::auto_correlator::AutoCorrelator;
        std::vector<cf32> x(500);
        for (int k = -250; k < 250; ++k) x[k + 250] = cf32((float)(std::cos((double)k) * 0.05), (float)(std::sin((double)k) * 0.05));
        AutoCorrelator a(5, 10);
        auto out = a.execute_block(x);
        bool zeros = true;
        for (auto v : out) zeros = zeros && v == cf32(0.f, 0.f);
        EXPECT(zeros && std::round(a.get_energy()[0] * 10000.0) == 125.0, "AutoCorrelator doc-test: energy 125, delay >= window");
        const size_t n = 5000;
        auto xr = rand_cf32(gen, n);
        AutoCorrelator b(48, 16);
        auto y = b.execute_block(xr);
        auto xd = widen(xr);
        std::vector<double> ref(2 * n);
        so_autocorr_fast(48, 16, nullptr, 0, xd.data(), n, ref.data());
        EXPECT(nerr(y, ref, n) <= TOL, "AutoCorrelator(48, 16) vs oracle");
    }
    // round-2 additions of the ABI through the compiled mirror: per-channel taps, NCO, the fused DDC, the multi-GPU context
    {
        const size_t C = 5, T = 200, n = 6000;
        std::vector<std::vector<double>> bank(C, std::vector<double>(T));
        for (size_t c = 0; c < C; ++c) {
            so_firdes_kaiser(T, 0.05 + 0.08 * (double)c, 60.0, 0.0, bank[c].data());
            bank[c] = f32round(bank[c]);
        }
        auto x = rand_cf32(gen, C * n);
        FIRFilter f(solid::filter::fir::per_channel, bank, 0.5);
        auto y = f.execute_block(x);
        DecimatingFIRFilter d(solid::filter::fir::per_channel, bank, 0.5, 4);
        auto yd = d.execute_block(x);
        double worst = 0, worst_d = 0;
        for (size_t c = 0; c < C; ++c) {
            std::vector<cf32> xc(x.begin() + c * n, x.begin() + (c + 1) * n), yc(y.begin() + c * n, y.begin() + (c + 1) * n);
            std::vector<cf32> ydc(yd.begin() + c * (n / 4), yd.begin() + (c + 1) * (n / 4));
            auto xd = widen(xc);
            std::vector<double> ref(2 * n + 2);
            so_fir_fast(bank[c].data(), T, 0, 0.5, 0.0, 0, 0, nullptr, xd.data(), n, ref.data());
            worst = std::max(worst, nerr(yc, ref, n));
            const size_t m = so_fir_fast(bank[c].data(), T, 0, 0.5, 0.0, 4, 0, nullptr, xd.data(), n, ref.data());
            EXPECT(m == n / 4, "per-channel decimator output count");
            worst_d = std::max(worst_d, nerr(ydc, ref, m));
        }
        EXPECT(worst <= TOL && worst_d <= TOL, "per-channel taps (FIR, decimator) vs one oracle object per channel");
    }
    {
        using solid::filter::ddc::DigitalDownConverter;
        using solid::nco::NCO;
        const size_t C = 3, T = 256, M = 8, n = 40000;
        std::vector<double> h(T);
        so_firdes_kaiser(T, 0.5 / 8 * 0.9, 80.0, 0.0, h.data());
        h = f32round(h);
        auto x = rand_cf32(gen, C * n);
        NCO osc(C);
        osc.set_frequency(0.1234);
        osc.set_phase(1.0, 2);
        auto mixed = osc.mix_down_block(x);
        DigitalDownConverter ddc(h, 1.0, M, 0.1234, C);
        sgpu_nco_set_phase(ddc.nco(), 2, 1.0);
        auto y = ddc.execute_block(x);
        EXPECT(ddc.last_fused(), "DDC took the fused kernel");
        double worst_mix = 0, worst = 0;
        for (size_t c = 0; c < C; ++c) {
            so_nco *o = so_nco_new();
            so_nco_set_frequency(o, 0.1234);
            if (c == 2) so_nco_set_phase(o, 1.0);
            std::vector<cf32> xc(x.begin() + c * n, x.begin() + (c + 1) * n), mc(mixed.begin() + c * n, mixed.begin() + (c + 1) * n);
            std::vector<cf32> yc(y.begin() + c * (n / M), y.begin() + (c + 1) * (n / M));
            auto xd = widen(xc);
            std::vector<double> md(2 * n), ref(2 * n + 2);
            so_nco_mix_block(o, 0, xd.data(), n, md.data());
            so_nco_free(o);
            worst_mix = std::max(worst_mix, nerr(mc, md, n));
            const size_t m = so_fir_fast(h.data(), T, 0, 1.0, 0.0, M, 0, nullptr, md.data(), n, ref.data());
            worst = std::max(worst, nerr(yc, ref, m));
        }
        EXPECT(worst_mix <= TOL, "NCO mix_down block vs oracle (nco/mod.rs:147-151)");
        EXPECT(worst <= TOL, "fused DDC vs oracle mix_down -> decimator");
    }
    {
        using solid::multi_gpu::Context;
        using solid::multi_gpu::ShardedFilter;
        Context ctx(std::vector<int>{0, 0, 0});  // three shards on this GPU: the code path of three GPUs
        EXPECT(ctx.devices() == 3, "context of three shards");
        const size_t T = 300, n = 400000;
        std::vector<double> h(T);
        so_firdes_kaiser(T, 0.1, 70.0, 0.0, h.data());
        h = f32round(h);
        auto x = rand_cf32(gen, n);
        auto f = ShardedFilter::fir(ctx, h, 1.0, 0, 1);  // ONE stream cut into time segments, halo sliced from the host buffer
        auto y = f.execute_block(x);
        auto xd = widen(x);
        std::vector<double> ref(2 * n + 2);
        so_fir_fast(h.data(), T, 0, 1.0, 0.0, 0, 0, nullptr, xd.data(), n, ref.data());
        EXPECT(f.last_segments() == 3 && nerr(y, ref, n) <= TOL, "one FIR stream over three shards vs oracle");
        auto g = ShardedFilter::fir(ctx, h, 1.0, 5, 7);  // 7 decimator channels in contiguous ranges
        const size_t nc = 9000;
        auto xc = rand_cf32(gen, 7 * nc);
        auto yc = g.execute_block(xc);
        double worst = 0;
        for (size_t c = 0; c < 7; ++c) {
            std::vector<cf32> xi(xc.begin() + c * nc, xc.begin() + (c + 1) * nc), yi(yc.begin() + c * (nc / 5), yc.begin() + (c + 1) * (nc / 5));
            auto xw = widen(xi);
            const size_t m = so_fir_fast(h.data(), T, 0, 1.0, 0.0, 5, 0, nullptr, xw.data(), nc, ref.data());
            worst = std::max(worst, nerr(yi, ref, m));
        }
        EXPECT(g.shards() == 3 && worst <= TOL, "decimator channels over three shards vs oracle");
    }
    std::printf("%s (%d failure%s), kernels launched: %llu\n", failures ? "FAILED" : "PASSED", failures,
                failures == 1 ? "" : "s", (unsigned long long)sgpu_launch_count());
    return failures ? 1 : 0;
}
