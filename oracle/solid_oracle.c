/*
 * solid_oracle.c -- CPU restatement (f64) of juliantos/solid-dsp's filtering hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libsolid_gpu.so) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED for FIR, FIR decimator, SOS / IIR, IIR decimator, IIR
 * interpolator, DotProduct and the Kaiser / notch designers -- every doc-test golden the
 * reference holds for this path (tests/golden/reference_doctests.json) is reproduced
 * bit-exactly in f64 by tests/test_oracle_golden.py.  UNPINNED by the reference's own
 * tests: FIR interpolator / PolyPhaseFilterBank numerics and Window (the reference has
 * construction-only doc-tests for them); the restatement follows the reference source
 * literally there.  The reference is Rust and no Rust toolchain exists in this image, so
 * there is no oracle/_ref build.
 *
 * Two flavours of every execute loop:
 *   structural  so_*_execute_block   : mirrors the Rust objects operation by operation
 *                                      (Window::push memmove, Window::to_vec malloc+copy,
 *                                      sequential DotProduct, one output append per
 *                                      sample).  This is the timed CPU baseline.
 *   closed form so_*_fast            : same arithmetic in the same accumulation order
 *                                      (bit-identical results) without the per-sample
 *                                      memmove/malloc; used as the checker on big inputs.
 *
 * Build with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 *
 * Complex numbers are interleaved (re, im) doubles everywhere.  Arithmetic follows
 * num-complex 0.4 (Cargo.toml:9, `num = "0.4"`; Cargo.lock is git-ignored so the patch
 * version is unpinned): (a+bi)(c+di) = (ac-bd) + (ad+bc)i, real*complex scales both
 * parts, += adds both parts.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_API __attribute__((visibility("default")))

typedef struct { double re, im; } cplx;

static inline cplx c_add(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cplx c_sub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
/* num-complex: impl Mul<Complex<T>> for Complex<T> */
static inline cplx c_mul(cplx a, cplx b) {
    cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
/* num-complex: impl Mul<Complex<f64>> for f64 / impl Mul<f64> for Complex<f64> */
static inline cplx r_mul(double a, cplx b) { cplx r = { a * b.re, a * b.im }; return r; }

/* ------------------------------------------------------------------------------------ */
/* resources::msb_index -- resources/mod.rs:21-23                                        */
SO_API size_t so_msb_index(size_t x) {
    size_t n = 0;
    while (x) { n++; x >>= 1; }
    return n;
}

/* ------------------------------------------------------------------------------------ */
/* Window<T> -- window/mod.rs:9-77.  Shift register, newest element at index 0.          */
typedef struct {
    size_t capacity, delay;
    cplx *buf; /* capacity + delay elements, zero initialised (window/mod.rs:26) */
} so_window;

static void win_init(so_window *w, size_t capacity, size_t delay) {
    w->capacity = capacity;
    w->delay = delay;
    w->buf = (cplx *)calloc(capacity + delay, sizeof(cplx));
}
static void win_free(so_window *w) { free(w->buf); w->buf = NULL; }
/* Window::push -- window/mod.rs:63-71: memmove capacity-1 elements right, write at 0 */
static inline void win_push(so_window *w, cplx e) {
    memmove(w->buf + 1, w->buf, (w->capacity - 1) * sizeof(cplx));
    w->buf[0] = e;
}
/* Window::to_vec -- window/mod.rs:44-51: fresh heap vector, copy `capacity` from +delay */
static inline cplx *win_to_vec(const so_window *w) {
    cplx *v = (cplx *)malloc(w->capacity * sizeof(cplx));
    memcpy(v, w->buf + w->delay, w->capacity * sizeof(cplx));
    return v;
}

/* ------------------------------------------------------------------------------------ */
/* DotProduct<T> -- dot_product/mod.rs:37-87,153-171                                     */
typedef struct {
    size_t len;
    int is_complex; /* Coef = Complex<f64> (1) or f64 (0) */
    cplx *buf;      /* stored order; real coefs keep im = 0 and use r_mul */
} so_dotprod;

enum { SO_FORWARD = 0, SO_REVERSE = 1 };

static void dp_init(so_dotprod *d, const double *coefs, size_t n, int is_complex, int direction) {
    d->len = n;
    d->is_complex = is_complex;
    d->buf = (cplx *)calloc(n ? n : 1, sizeof(cplx));
    for (size_t i = 0; i < n; i++) {
        size_t src = direction == SO_REVERSE ? n - 1 - i : i; /* dot_product/mod.rs:75-84 */
        if (is_complex) { d->buf[i].re = coefs[2 * src]; d->buf[i].im = coefs[2 * src + 1]; }
        else            { d->buf[i].re = coefs[src];     d->buf[i].im = 0.0; }
    }
}
static void dp_free(so_dotprod *d) { free(d->buf); d->buf = NULL; }
/* Execute::execute -- dot_product/mod.rs:159-170: min(len) terms, sequential from zero */
static inline cplx dp_execute(const so_dotprod *d, const cplx *x, size_t nx) {
    size_t it = nx < d->len ? nx : d->len;
    cplx sum = { 0.0, 0.0 };
    if (d->is_complex) {
        for (size_t i = 0; i < it; i++) sum = c_add(sum, c_mul(d->buf[i], x[i]));
    } else {
        for (size_t i = 0; i < it; i++) sum = c_add(sum, r_mul(d->buf[i].re, x[i]));
    }
    return sum;
}

/* stand-alone DotProduct for the golden tests */
SO_API void so_dot_execute(const double *coefs, size_t n, int is_complex, int direction,
                           const double *x, size_t nx, double *out2) {
    so_dotprod d;
    dp_init(&d, coefs, n, is_complex, direction);
    cplx r = dp_execute(&d, (const cplx *)x, nx);
    out2[0] = r.re; out2[1] = r.im;
    dp_free(&d);
}
/* DotProduct::coefficents -- dot_product/mod.rs:102-109: the STORED order */
SO_API void so_dot_coefficients(const double *coefs, size_t n, int is_complex, int direction,
                                double *out) {
    so_dotprod d;
    dp_init(&d, coefs, n, is_complex, direction);
    for (size_t i = 0; i < n; i++) {
        if (is_complex) { out[2 * i] = d.buf[i].re; out[2 * i + 1] = d.buf[i].im; }
        else out[i] = d.buf[i].re;
    }
    dp_free(&d);
}

/* ------------------------------------------------------------------------------------ */
/* FIRFilter / DecimatingFIRFilter -- fir/mod.rs:58-88,209-241; fir/decim.rs:5-42,115-139,
 * 221-256.  One object covers both: decimation == 0 means the plain FIRFilter.           */
typedef struct {
    so_window window;
    so_dotprod coefs;
    cplx scale;          /* Coef-typed: im ignored when coefs are real */
    size_t decimation;   /* 0 = FIRFilter */
    size_t current_item; /* fir/decim.rs:8 */
} so_fir;

enum {
    SO_OK = 0,
    SO_ERR_COEF_LEN_ZERO = -1,   /* FIRErrorCode::CoefficientsLengthZero  fir/mod.rs:41 */
    SO_ERR_DECIM_LT_ONE = -2,    /* FIRErrorCode::DecimationLessThanOne   fir/mod.rs:42 */
    SO_ERR_INTERP_LT_ONE = -3,   /* FIRErrorCode::InterpolationLessThanOne fir/mod.rs:43 */
    SO_ERR_NOT_ENOUGH_FILTERS = -4, /* FIRErrorCode::NotEnoughFilters     fir/mod.rs:44 */
    SO_ERR_IIR_NUM_ZERO = -10,   /* IIRErrorCode::NumeratorLengthZero  iir/mod.rs:42 */
    SO_ERR_IIR_DEN_ZERO = -11,   /* IIRErrorCode::DenominatorLengthZero */
    SO_ERR_SOS_SIZE_ZERO = -12,  /* SecondOrderSectionSizeZero */
    SO_ERR_SOS_MISMATCH = -13,   /* SecondOrderSectionSizeMismatch */
    SO_ERR_SOS_NOT_MULT3 = -14,  /* SecondOrderSectionSizeNotMultpleOf3 */
    SO_ERR_IIR_DECIM = -15,      /* DecimationLessThanOne */
    SO_ERR_IIR_INTERP = -16,     /* InterpolationLessThanOne */
    SO_ERR_SOS_RANGE = -17       /* SecondOrderErrorCode::CoefficientsNotInRange sos.rs:20 */
};

static inline cplx apply_scale(cplx v, cplx scale, int coef_complex) {
    /* fir/mod.rs:211: Out * Coef -> Complex*f64 scales both parts; Complex*Complex textbook */
    if (coef_complex) return c_mul(v, scale);
    cplx r = { v.re * scale.re, v.im * scale.re };
    return r;
}

/* decimation: 0 => FIRFilter::new (fir/mod.rs:79-88); >=1 => DecimatingFIRFilter::new
 * (fir/decim.rs:27-42).  `is_decim` distinguishes "decimation argument of 0" (an error). */
SO_API so_fir *so_fir_new(const double *coefs, size_t n, int coef_complex, double scale_re,
                          double scale_im, int is_decim, size_t decimation, int *err) {
    if (n == 0) { *err = SO_ERR_COEF_LEN_ZERO; return NULL; }
    if (is_decim && decimation < 1) { *err = SO_ERR_DECIM_LT_ONE; return NULL; }
    so_fir *f = (so_fir *)calloc(1, sizeof(so_fir));
    win_init(&f->window, (size_t)1 << so_msb_index(n), 0); /* fir/mod.rs:85 */
    dp_init(&f->coefs, coefs, n, coef_complex, SO_REVERSE); /* fir/mod.rs:86 */
    f->scale.re = scale_re; f->scale.im = scale_im;
    f->decimation = is_decim ? decimation : 0;
    f->current_item = 0;
    *err = SO_OK;
    return f;
}
SO_API void so_fir_free(so_fir *f) {
    if (!f) return;
    win_free(&f->window); dp_free(&f->coefs); free(f);
}
SO_API void so_fir_set_scale(so_fir *f, double re, double im) { f->scale.re = re; f->scale.im = im; }
SO_API size_t so_fir_window_capacity(const so_fir *f) { return f->window.capacity; }
SO_API size_t so_fir_current_item(const so_fir *f) { return f->current_item; }

/* DecimatingFIRFilter::write -- fir/decim.rs:136-139 */
SO_API void so_fir_write(so_fir *f, const double *x, size_t n) {
    const cplx *in = (const cplx *)x;
    if (f->decimation) f->current_item = (f->current_item + n) % f->decimation;
    for (size_t i = 0; i < n; i++) win_push(&f->window, in[i]);
}

/* output vector with amortised growth, standing in for Vec::append (fir/mod.rs:238) */
typedef struct { cplx *p; size_t len, cap; } ovec;
static inline void ovec_push(ovec *v, cplx e) {
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4;
        v->p = (cplx *)realloc(v->p, v->cap * sizeof(cplx));
    }
    v->p[v->len++] = e;
}

/* execute_block -- fir/mod.rs:235-241 -> :209-212; fir/decim.rs:250-256 -> :221-228.
 * Returns the number of outputs written to `out` (capacity out_cap complex values). */
SO_API size_t so_fir_execute_block(so_fir *f, const double *x, size_t n, double *out,
                                   size_t out_cap) {
    const cplx *in = (const cplx *)x;
    ovec block = { NULL, 0, 0 };
    for (size_t i = 0; i < n; i++) {
        if (f->decimation) f->current_item = (f->current_item + 1) % f->decimation; /* decim.rs:116 */
        win_push(&f->window, in[i]);
        if (!f->decimation || f->current_item == 0) {
            cplx *vec = win_to_vec(&f->window);
            cplx dot = dp_execute(&f->coefs, vec, f->window.capacity);
            free(vec);
            cplx *one = (cplx *)malloc(sizeof(cplx)); /* vec![..] -- fir/mod.rs:211 */
            *one = apply_scale(dot, f->scale, f->coefs.is_complex);
            ovec_push(&block, *one);
            free(one);
        }
    }
    size_t nout = block.len < out_cap ? block.len : out_cap;
    if (nout) memcpy(out, block.p, nout * sizeof(cplx));
    free(block.p);
    return block.len;
}

/* Closed form, identical accumulation order (newest sample first):
 *   y[n] = scale * sum_{i<T} h[T-1-i] * x[n-i],  emitted when (count+1) % M == 0.
 * `hist` holds the T-1 samples preceding x[0], oldest first; `count0` is current_item on
 * entry.  hist may be NULL (zeros).                                                      */
SO_API size_t so_fir_fast(const double *coefs, size_t T, int coef_complex, double scale_re,
                          double scale_im, size_t decimation, size_t count0,
                          const double *hist, const double *x, size_t n, double *out) {
    const cplx *in = (const cplx *)x;
    const cplx *h = (const cplx *)hist;
    cplx *o = (cplx *)out;
    cplx scale = { scale_re, scale_im };
    size_t M = decimation ? decimation : 1;
    size_t cur = decimation ? count0 % M : 0;
    size_t nout = 0;
    /* contiguous buffer [hist(T-1) | x(n)] so the inner loop is a plain backwards walk */
    cplx *buf = (cplx *)calloc(T - 1 + n + 1, sizeof(cplx));
    if (h && T > 1) memcpy(buf, h, (T - 1) * sizeof(cplx));
    if (n) memcpy(buf + (T - 1), in, n * sizeof(cplx));
    for (size_t i = 0; i < n; i++) {
        cur = (cur + 1) % M;
        if (cur != 0) continue;
        const cplx *p = buf + (T - 1) + i; /* newest */
        cplx sum = { 0.0, 0.0 };
        if (coef_complex) {
            for (size_t k = 0; k < T; k++) {
                cplx c = { coefs[2 * (T - 1 - k)], coefs[2 * (T - 1 - k) + 1] };
                sum = c_add(sum, c_mul(c, p[-(ptrdiff_t)k]));
            }
        } else {
            for (size_t k = 0; k < T; k++)
                sum = c_add(sum, r_mul(coefs[T - 1 - k], p[-(ptrdiff_t)k]));
        }
        o[nout++] = apply_scale(sum, scale, coef_complex);
    }
    free(buf);
    return nout;
}

/* ------------------------------------------------------------------------------------ */
/* PolyPhaseFilterBank -- fir/pfb.rs:3-8,24-49,81-90                                     */
typedef struct {
    so_window window;
    so_dotprod *coefs; /* `filters` sub-filters, each stored reversed, FORWARD dot */
    size_t filters, sub_len;
    int coef_complex;
    cplx scale; /* stored, never applied (pfb.rs:85-90) */
} so_pfb;

SO_API so_pfb *so_pfb_new(const double *coefs, size_t n, int coef_complex, size_t filters,
                          double scale_re, double scale_im, int *err) {
    if (filters == 0) { *err = SO_ERR_NOT_ENOUGH_FILTERS; return NULL; } /* pfb.rs:25 */
    if (n == 0) { *err = SO_ERR_COEF_LEN_ZERO; return NULL; }            /* pfb.rs:27 */
    size_t sub_len = n / filters;                                        /* pfb.rs:32 */
    if (sub_len == 0) { *err = SO_ERR_NOT_ENOUGH_FILTERS; return NULL; } /* ref: Window::new(0) assert panic */
    so_pfb *p = (so_pfb *)calloc(1, sizeof(so_pfb));
    p->filters = filters; p->sub_len = sub_len; p->coef_complex = coef_complex;
    p->scale.re = scale_re; p->scale.im = scale_im;
    p->coefs = (so_dotprod *)calloc(filters, sizeof(so_dotprod));
    size_t w = coef_complex ? 2 : 1;
    double *rev = (double *)calloc(sub_len * w, sizeof(double));
    for (size_t f = 0; f < filters; f++) {
        for (size_t idx = 0; idx < sub_len; idx++) /* pfb.rs:36-38 */
            for (size_t c = 0; c < w; c++)
                rev[(sub_len - idx - 1) * w + c] = coefs[(f + idx * filters) * w + c];
        dp_init(&p->coefs[f], rev, sub_len, coef_complex, SO_FORWARD); /* pfb.rs:40 */
    }
    free(rev);
    win_init(&p->window, sub_len, 0); /* pfb.rs:46 */
    *err = SO_OK;
    return p;
}
SO_API void so_pfb_free(so_pfb *p) {
    if (!p) return;
    for (size_t f = 0; f < p->filters; f++) dp_free(&p->coefs[f]);
    free(p->coefs); win_free(&p->window); free(p);
}
SO_API void so_pfb_push(so_pfb *p, double re, double im) { cplx e = { re, im }; win_push(&p->window, e); }
SO_API void so_pfb_execute(so_pfb *p, size_t index, double *out2) { /* pfb.rs:85-90 */
    cplx *vec = win_to_vec(&p->window);
    cplx r = dp_execute(&p->coefs[index], vec, p->window.capacity);
    free(vec);
    out2[0] = r.re; out2[1] = r.im;
}
SO_API size_t so_pfb_sub_len(const so_pfb *p) { return p->sub_len; }
/* PolyPhaseFilterBank::coefficents -- pfb.rs:71-73: per-filter stored order, flattened */
SO_API void so_pfb_coefficients(const so_pfb *p, double *out) {
    size_t w = p->coef_complex ? 2 : 1, k = 0;
    for (size_t f = 0; f < p->filters; f++)
        for (size_t i = 0; i < p->sub_len; i++) {
            out[k++] = p->coefs[f].buf[i].re;
            if (w == 2) out[k++] = p->coefs[f].buf[i].im;
        }
}

/* InterpolatingFIRFilter -- fir/interp.rs:6-10,27-54,93-111 */
typedef struct { so_pfb *bank; size_t interpolation; } so_firinterp;

/* sub-filter length exactly as fir/interp.rs:35-40 computes it (in f32!) */
SO_API size_t so_interp_sub_len(size_t n, size_t interpolation) {
    float q = (float)n / (float)interpolation;
    return (q == floorf(q)) ? (size_t)q : (size_t)ceilf(q);
}

SO_API so_firinterp *so_firinterp_new(const double *coefs, size_t n, int coef_complex,
                                      size_t interpolation, int *err) {
    if (n == 0) { *err = SO_ERR_COEF_LEN_ZERO; return NULL; }
    if (interpolation < 1) { *err = SO_ERR_INTERP_LT_ONE; return NULL; }
    size_t sub = so_interp_sub_len(n, interpolation);
    size_t eff = sub * interpolation; /* interp.rs:43 */
    size_t w = coef_complex ? 2 : 1;
    /* interp.rs:44-46: copy then Vec::resize(effective_length, zero) -- resize TRUNCATES if the
     * f32 quotient rounded down far enough that eff < n (only for n > 2^24). */
    double *padded = (double *)calloc((eff > n ? eff : n) * w, sizeof(double));
    memcpy(padded, coefs, n * w * sizeof(double));
    so_firinterp *f = (so_firinterp *)calloc(1, sizeof(so_firinterp));
    f->bank = so_pfb_new(padded, eff, coef_complex, interpolation, 1.0, 0.0, err); /* interp.rs:48 */
    free(padded);
    if (!f->bank) { free(f); return NULL; }
    f->interpolation = interpolation;
    return f;
}
SO_API void so_firinterp_free(so_firinterp *f) { if (f) { so_pfb_free(f->bank); free(f); } }
SO_API void so_firinterp_set_scale(so_firinterp *f, double re, double im) {
    f->bank->scale.re = re; f->bank->scale.im = im; /* interp.rs:57; never applied */
}
SO_API void so_firinterp_coefficients(const so_firinterp *f, double *out) { so_pfb_coefficients(f->bank, out); }
SO_API size_t so_firinterp_sub_len(const so_firinterp *f) { return f->bank->sub_len; }

/* execute_block -- fir/interp.rs:102-111 */
SO_API size_t so_firinterp_execute_block(so_firinterp *f, const double *x, size_t n,
                                         double *out, size_t out_cap) {
    const cplx *in = (const cplx *)x;
    ovec block = { NULL, 0, 0 };
    for (size_t i = 0; i < n; i++) {
        win_push(&f->bank->window, in[i]);
        for (size_t p = 0; p < f->bank->filters; p++) {
            cplx *vec = win_to_vec(&f->bank->window);
            cplx r = dp_execute(&f->bank->coefs[p], vec, f->bank->window.capacity);
            free(vec);
            ovec_push(&block, r);
        }
    }
    size_t nout = block.len < out_cap ? block.len : out_cap;
    if (nout) memcpy(out, block.p, nout * sizeof(cplx));
    free(block.p);
    return block.len;
}

/* Closed form: y[nL+p] = sum_{j<S} hpad[p + (S-1-j)L] * x[n-j], j = 0 (newest) first, no
 * scale.  hist = the S-1 samples before x[0], oldest first (NULL = zeros).              */
SO_API size_t so_firinterp_fast(const double *coefs, size_t T, int coef_complex, size_t L,
                                const double *hist, const double *x, size_t n, double *out) {
    size_t S = so_interp_sub_len(T, L);
    size_t w = coef_complex ? 2 : 1;
    size_t eff = S * L;
    double *hp = (double *)calloc((eff > T ? eff : T) * w, sizeof(double));
    memcpy(hp, coefs, T * w * sizeof(double));
    cplx *buf = (cplx *)calloc(S - 1 + n + 1, sizeof(cplx));
    if (hist && S > 1) memcpy(buf, hist, (S - 1) * sizeof(cplx));
    if (n) memcpy(buf + (S - 1), x, n * sizeof(cplx));
    cplx *o = (cplx *)out;
    for (size_t i = 0; i < n; i++) {
        const cplx *pnew = buf + (S - 1) + i;
        for (size_t p = 0; p < L; p++) {
            cplx sum = { 0.0, 0.0 };
            for (size_t j = 0; j < S; j++) {
                size_t t = p + (S - 1 - j) * L;
                if (coef_complex) {
                    cplx c = { hp[2 * t], hp[2 * t + 1] };
                    sum = c_add(sum, c_mul(c, pnew[-(ptrdiff_t)j]));
                } else {
                    sum = c_add(sum, r_mul(hp[t], pnew[-(ptrdiff_t)j]));
                }
            }
            o[i * L + p] = sum;
        }
    }
    free(buf); free(hp);
    return n * L;
}

/* ------------------------------------------------------------------------------------ */
/* SecondOrderFilter -- iir/sos.rs:34-39,55-75,92-114.  Real coefficients only: the IIR
 * Filter impl requires Coef: Conj + Real, i.e. Coef = f64 (iir/mod.rs:244-250).         */
typedef struct {
    so_window form_buffer_ii; /* Window(3) */
    so_dotprod numerator_coefs;   /* holds a1,a2  (field names swapped in the reference) */
    so_dotprod denominator_coefs; /* holds b0,b1,b2 */
} so_sos;

static int sos_init(so_sos *s, const double *ff, size_t nff, const double *fb, size_t nfb) {
    if (nff < 3 || nfb < 3) return SO_ERR_SOS_RANGE; /* sos.rs:56-60 */
    double a0 = fb[0];
    double b[3] = { ff[0] / a0, ff[1] / a0, ff[2] / a0 }; /* sos.rs:62-68 */
    double a[3] = { fb[0] / a0, fb[1] / a0, fb[2] / a0 };
    win_init(&s->form_buffer_ii, 3, 0);
    dp_init(&s->numerator_coefs, a + 1, 2, 0, SO_FORWARD);
    dp_init(&s->denominator_coefs, b, 3, 0, SO_FORWARD);
    return SO_OK;
}
static void sos_free(so_sos *s) {
    win_free(&s->form_buffer_ii); dp_free(&s->numerator_coefs); dp_free(&s->denominator_coefs);
}
/* SecondOrderFilter::execute -- sos.rs:92-114 */
static inline cplx sos_execute(so_sos *s, cplx input) {
    cplx *buffer = win_to_vec(&s->form_buffer_ii);      /* :98 */
    buffer[2] = buffer[1];                               /* :99 */
    buffer[1] = buffer[0];                               /* :100 */
    cplx denom_output = dp_execute(&s->numerator_coefs, buffer + 1, 2); /* :102 */
    free(buffer);
    cplx mixed = c_sub(input, denom_output);             /* :104-108 */
    win_push(&s->form_buffer_ii, mixed);                 /* :110 */
    cplx *b2 = win_to_vec(&s->form_buffer_ii);           /* :111 */
    cplx y = dp_execute(&s->denominator_coefs, b2, 3);   /* :113 */
    free(b2);
    return y;
}

SO_API so_sos *so_sos_new(const double *ff, size_t nff, const double *fb, size_t nfb, int *err) {
    so_sos *s = (so_sos *)calloc(1, sizeof(so_sos));
    *err = sos_init(s, ff, nff, fb, nfb);
    if (*err) { free(s); return NULL; }
    return s;
}
SO_API void so_sos_free(so_sos *s) { if (s) { sos_free(s); free(s); } }
SO_API void so_sos_execute(so_sos *s, double re, double im, double *out2) {
    cplx in = { re, im };
    cplx y = sos_execute(s, in);
    out2[0] = y.re; out2[1] = y.im;
}
/* numerator_coefs()/denominator_coefs() accessors -- sos.rs:116-134 (swapped names kept) */
SO_API void so_sos_numerator_coefs(const so_sos *s, double *out2) {
    out2[0] = s->numerator_coefs.buf[0].re; out2[1] = s->numerator_coefs.buf[1].re;
}
SO_API void so_sos_denominator_coefs(const so_sos *s, double *out3) {
    for (int i = 0; i < 3; i++) out3[i] = s->denominator_coefs.buf[i].re;
}

/* ------------------------------------------------------------------------------------ */
/* IIRFilter (+ Decimating / Interpolating wrappers) -- iir/mod.rs:68-75,92-164,270-316;
 * iir/decim.rs:190-233; iir/interp.rs:184-221                                            */
typedef struct {
    int second_order;          /* IIRFilterType */
    so_window buffer;          /* Normal: Window(max(len_b, len_a)); SOS: Window(2*nsec), unused */
    so_dotprod numerator_coefs, denominator_coefs;
    so_sos *sections; size_t nsec;
    size_t decimation, index;  /* DecimatingIIRFilter: iir/decim.rs:6-10 (0 = none) */
    size_t interpolation;      /* InterpolatingIIRFilter (0 = none) */
} so_iir;

SO_API so_iir *so_iir_new(const double *ff, size_t nff, const double *fb, size_t nfb,
                          int second_order, int wrapper /*0 none,1 decim,2 interp*/,
                          size_t factor, int *err) {
    if (wrapper) { /* iir/decim.rs:31-41, iir/interp.rs:30-40 */
        if (nff == 0) { *err = SO_ERR_IIR_NUM_ZERO; return NULL; }
        if (nfb == 0) { *err = SO_ERR_IIR_DEN_ZERO; return NULL; }
        if (factor < 1) { *err = wrapper == 1 ? SO_ERR_IIR_DECIM : SO_ERR_IIR_INTERP; return NULL; }
    }
    so_iir *f = (so_iir *)calloc(1, sizeof(so_iir));
    f->second_order = second_order;
    if (!second_order) { /* iir/mod.rs:98-130 */
        if (nff == 0) { *err = SO_ERR_IIR_NUM_ZERO; free(f); return NULL; }
        if (nfb == 0) { *err = SO_ERR_IIR_DEN_ZERO; free(f); return NULL; }
        size_t wl = nfb > nff ? nfb : nff;
        win_init(&f->buffer, wl, 0);
        double a0 = fb[0];
        double *num = (double *)malloc(nff * sizeof(double));
        double *den = (double *)malloc(nfb * sizeof(double));
        for (size_t i = 0; i < nff; i++) num[i] = ff[i] / a0;
        for (size_t i = 0; i < nfb; i++) den[i] = fb[i] / a0;
        dp_init(&f->numerator_coefs, num, nff, 0, SO_FORWARD);
        dp_init(&f->denominator_coefs, den + 1, nfb - 1, 0, SO_FORWARD);
        free(num); free(den);
    } else { /* iir/mod.rs:131-163 */
        if (nff != nfb) { *err = SO_ERR_SOS_MISMATCH; free(f); return NULL; }
        if (nff == 0) { *err = SO_ERR_SOS_SIZE_ZERO; free(f); return NULL; }
        if (nff % 3 != 0) { *err = SO_ERR_SOS_NOT_MULT3; free(f); return NULL; }
        f->nsec = nff / 3;
        win_init(&f->buffer, f->nsec * 2, 0);
        f->sections = (so_sos *)calloc(f->nsec, sizeof(so_sos));
        for (size_t i = 0; i < f->nsec; i++) {
            int e = sos_init(&f->sections[i], ff + 3 * i, 3, fb + 3 * i, 3);
            if (e) { *err = e; free(f->sections); win_free(&f->buffer); free(f); return NULL; }
        }
        dp_init(&f->numerator_coefs, ff, nff, 0, SO_FORWARD);
        dp_init(&f->denominator_coefs, fb, nfb, 0, SO_FORWARD);
    }
    if (wrapper == 1) f->decimation = factor;
    if (wrapper == 2) f->interpolation = factor;
    *err = SO_OK;
    return f;
}
SO_API void so_iir_free(so_iir *f) {
    if (!f) return;
    for (size_t i = 0; i < f->nsec; i++) sos_free(&f->sections[i]);
    free(f->sections);
    win_free(&f->buffer); dp_free(&f->numerator_coefs); dp_free(&f->denominator_coefs);
    free(f);
}
/* IIRFilter::execute -- iir/mod.rs:270-289 */
static inline cplx iir_execute(so_iir *f, cplx input) {
    if (!f->second_order) {
        cplx *buffer = win_to_vec(&f->buffer);
        cplx denom = dp_execute(&f->denominator_coefs, buffer, f->buffer.capacity - 1); /* :274 */
        free(buffer);
        cplx mixed = c_sub(input, denom);
        win_push(&f->buffer, mixed);
        cplx *b2 = win_to_vec(&f->buffer);
        cplx y = dp_execute(&f->numerator_coefs, b2, f->buffer.capacity);
        free(b2);
        return y;
    }
    cplx y = sos_execute(&f->sections[0], input);
    for (size_t i = 1; i < f->nsec; i++) y = sos_execute(&f->sections[i], y);
    return y;
}
SO_API size_t so_iir_execute_block(so_iir *f, const double *x, size_t n, double *out,
                                   size_t out_cap) {
    const cplx *in = (const cplx *)x;
    const cplx zero = { 0.0, 0.0 };
    ovec block = { NULL, 0, 0 };
    for (size_t i = 0; i < n; i++) {
        if (f->decimation) { /* iir/decim.rs:222-233 */
            f->index = (f->index + 1) % f->decimation;
            cplx y = iir_execute(f, in[i]);
            if (f->index == 0) ovec_push(&block, y);
        } else if (f->interpolation) { /* iir/interp.rs:184-190 */
            ovec_push(&block, iir_execute(f, in[i]));
            for (size_t k = 1; k < f->interpolation; k++) ovec_push(&block, iir_execute(f, zero));
        } else {
            ovec_push(&block, iir_execute(f, in[i]));
        }
    }
    size_t nout = block.len < out_cap ? block.len : out_cap;
    if (nout) memcpy(out, block.p, nout * sizeof(cplx));
    free(block.p);
    return block.len;
}

/* Closed-form SOS cascade, identical arithmetic order, no allocation.
 *   v0 = x - ((0 + a1*v1) + a2*v2);  y = ((0 + b0*v0) + b1*v1) + b2*v2
 * state: [nsec][2] complex (v1, v2), updated in place.  ff/fb are the RAW flat arrays
 * (normalised by fb[3i] here, as sos.rs:62-68 does).                                      */
SO_API void so_sos_cascade_fast(const double *ff, const double *fb, size_t nsec, double *state,
                                const double *x, size_t n, double *out) {
    const cplx *in = (const cplx *)x;
    cplx *o = (cplx *)out;
    cplx *st = (cplx *)state;
    double *b = (double *)malloc(nsec * 3 * sizeof(double));
    double *a = (double *)malloc(nsec * 3 * sizeof(double));
    for (size_t s = 0; s < nsec; s++)
        for (int k = 0; k < 3; k++) {
            b[3 * s + k] = ff[3 * s + k] / fb[3 * s];
            a[3 * s + k] = fb[3 * s + k] / fb[3 * s];
        }
    const cplx zero = { 0.0, 0.0 };
    for (size_t i = 0; i < n; i++) {
        cplx y = in[i];
        for (size_t s = 0; s < nsec; s++) {
            cplx v1 = st[2 * s], v2 = st[2 * s + 1];
            cplx fbk = c_add(c_add(zero, r_mul(a[3 * s + 1], v1)), r_mul(a[3 * s + 2], v2));
            cplx v0 = c_sub(y, fbk);
            y = c_add(c_add(c_add(zero, r_mul(b[3 * s], v0)), r_mul(b[3 * s + 1], v1)),
                      r_mul(b[3 * s + 2], v2));
            st[2 * s + 1] = v1; st[2 * s] = v0;
        }
        o[i] = y;
    }
    free(a); free(b);
}

/* ------------------------------------------------------------------------------------ */
/* AutoCorrelator<C> -- filter/auto_correlator/mod.rs:24-35,51-62,99-111,165-191,214-216  */
/* Structural mirror.  Note what Window(window_size, delay) really does (window/mod.rs:17-34,   */
/* 44-51,63-71): the buffer has window_size + delay elements, push() shifts only the first      */
/* window_size - 1 and to_vec() copies window_size elements starting at +delay, so the delayed  */
/* window holds conj(x[n-delay-i]) for i < window_size - delay and ZEROS after that:            */
/*   execute() = sum_{i < window_size - delay} x[n-i] * conj(x[n-delay-i])   (0 if delay >= size) */
typedef struct {
    size_t window_size, delay;
    so_window window, window_with_delay;
    double *energy_buffer;
    double energy_sum;
    size_t energy_index;
} so_autocorr;

SO_API so_autocorr *so_autocorr_new(size_t window_size, size_t delay) {
    if (window_size == 0) return NULL; /* Window::new asserts capacity > 0 */
    so_autocorr *a = (so_autocorr *)calloc(1, sizeof(so_autocorr));
    a->window_size = window_size;
    a->delay = delay;
    win_init(&a->window, window_size, 0);
    win_init(&a->window_with_delay, window_size, delay);
    a->energy_buffer = (double *)calloc(window_size, sizeof(double));
    return a;
}
SO_API void so_autocorr_free(so_autocorr *a) {
    if (!a) return;
    win_free(&a->window);
    win_free(&a->window_with_delay);
    free(a->energy_buffer);
    free(a);
}
/* push -- auto_correlator/mod.rs:99-111 */
SO_API void so_autocorr_push(so_autocorr *a, double re, double im) {
    cplx s = { re, im }, sc = { re, -im };
    win_push(&a->window, s);
    win_push(&a->window_with_delay, sc);
    double e2 = c_mul(s, sc).re;
    a->energy_sum -= a->energy_buffer[a->energy_index];
    a->energy_sum += e2;
    a->energy_buffer[a->energy_index] = e2;
    a->energy_index = (a->energy_index + 1) % a->window_size;
}
/* execute -- auto_correlator/mod.rs:165-172: two to_vec copies, zip, map, sum (from zero) */
SO_API void so_autocorr_execute(const so_autocorr *a, double *out2) {
    cplx *x = win_to_vec(&a->window), *y = win_to_vec(&a->window_with_delay);
    cplx sum = { 0.0, 0.0 };
    for (size_t i = 0; i < a->window_size; i++) sum = c_add(sum, c_mul(x[i], y[i]));
    free(x);
    free(y);
    out2[0] = sum.re;
    out2[1] = sum.im;
}
/* execute_block -- auto_correlator/mod.rs:184-191 */
SO_API size_t so_autocorr_execute_block(so_autocorr *a, const double *x, size_t n, double *out) {
    for (size_t k = 0; k < n; k++) {
        so_autocorr_push(a, x[2 * k], x[2 * k + 1]);
        so_autocorr_execute(a, out + 2 * k);
    }
    return n;
}
SO_API double so_autocorr_get_energy(const so_autocorr *a) { return a->energy_sum; }

/* closed form, same accumulation order (i ascending from the newest sample), zero history:    */
/* out[n] = sum_{i < W-d} x[n-i] * conj(x[n-d-i]); the terms with i >= W-d multiply by the     */
/* window's zero tail and are added as exact zeros, which cannot change an f64 sum.             */
SO_API void so_autocorr_fast(size_t window_size, size_t delay, const double *hist, size_t nhist,
                             const double *x, size_t n, double *out) {
    /* hist: the nhist samples preceding x[0], oldest first (may be NULL) */
    const size_t wd = delay < window_size ? window_size - delay : 0;
    for (size_t k = 0; k < n; k++) {
        cplx sum = { 0.0, 0.0 };
        for (size_t i = 0; i < wd; i++) {
            ptrdiff_t ia = (ptrdiff_t)k - (ptrdiff_t)i, ib = ia - (ptrdiff_t)delay;
            cplx xa = { 0.0, 0.0 }, xb = { 0.0, 0.0 };
            if (ia >= 0) { xa.re = x[2 * ia]; xa.im = x[2 * ia + 1]; }
            else if ((ptrdiff_t)nhist + ia >= 0) { xa.re = hist[2 * (nhist + ia)]; xa.im = hist[2 * (nhist + ia) + 1]; }
            if (ib >= 0) { xb.re = x[2 * ib]; xb.im = -x[2 * ib + 1]; }
            else if ((ptrdiff_t)nhist + ib >= 0) { xb.re = hist[2 * (nhist + ib)]; xb.im = -hist[2 * (nhist + ib) + 1]; }
            sum = c_add(sum, c_mul(xa, xb));
        }
        out[2 * k] = sum.re;
        out[2 * k + 1] = sum.im;
    }
}

/* ------------------------------------------------------------------------------------ */
/* Design helpers needed to produce identical taps on both sides.                        */
/* math/mod.rs:17-27 */
SO_API double so_sinc(double x) {
    if (fabs(x) < 0.01)
        return cos(M_PI * x / 2.0) * cos(M_PI * x / 4.0) * cos(M_PI * x / 8.0);
    return sin(M_PI * x) / (M_PI * x);
}
/* math/mod.rs:171-183 */
SO_API double so_lngamma(double x) {
    if (x < 0.0) return 0.0;
    if (x < 10.0) return so_lngamma(x + 1.0) - log(x);
    double g = 0.5 * (log(2.0 * M_PI) - log(x));
    return g + x * (log(x + (1.0 / (12.0 * x - 0.1 / x))) - 1.0);
}
/* math/mod.rs:156-169 */
SO_API double so_gamma(double x) {
    if (x < 0.0) {
        double t0 = so_gamma(1.0 - x);
        double t1 = sin(M_PI * x);
        return M_PI / (t0 * t1);
    }
    return exp(so_lngamma(x));
}
/* math/mod.rs:66-100 */
SO_API double so_lnbesseli(double z, double nu) {
    if (z == 0.0) return nu == 0.0 ? 0.0 : -1.7976931348623157e308;
    if (nu == 0.5) return 0.5 * log(2.0 / (M_PI * z)) + log(sinh(z));
    if (z < 0.001 * sqrt(nu + 1.0)) return -so_gamma(nu + 1.0) + nu * log(0.5 * z);
    double t0 = nu * log(0.5 * z);
    double y = 0.0;
    for (size_t k = 0; k < 64; k++) {
        double t1 = 2.0 * (double)k * log(0.5 * z);
        double t2 = so_lngamma((double)k + 1.0);
        double t3 = so_lngamma(nu + (double)k + 1.0);
        y += exp(t1 - t2 - t3);
    }
    return t0 + log(y);
}
/* math/mod.rs:41-64 */
SO_API double so_besseli(double z, double nu) {
    if (z == 0.0) return nu == 0.0 ? 1.0 : 0.0;
    if (nu == 0.5) return sqrt(2.0 / (M_PI * z)) * sinh(z);
    if (z < 0.001 * sqrt(nu + 1.0)) return pow(0.5 * z, nu) / so_gamma(nu + 1.0);
    return exp(so_lnbesseli(z, nu));
}
/* windows/kaiser.rs:33-46 */
SO_API double so_window_kaiser(size_t index, size_t len, double beta) {
    double t = (double)index - (double)(len - 1) / 2.0;
    double r = 2.0 * t / ((double)(len - 1));
    double a = so_besseli(beta * sqrt(1.0 - r * r), 0.0);
    double b = so_besseli(beta, 0.0);
    return a / b;
}
/* firdes/mod.rs:243-253 */
SO_API double so_kaiser_beta(double as) {
    double a = fabs(as);
    if (a > 50.0) return 0.1102 * (a - 8.7);
    if (a > 21.0) return 0.5842 * pow(a - 21.0, 0.4) + 0.07886 * (a - 21.0);
    return 0.0;
}
/* firdes/mod.rs:278-305.  Returns 0 on success, -1 on an argument the reference rejects. */
SO_API int so_firdes_kaiser(size_t len, double fc, double as, double mu, double *h) {
    if (!(mu >= -0.5 && mu <= 0.5)) return -1;
    if (!(fc >= 0.0 && fc <= 0.5)) return -1;
    if (as <= 0.0) return -1;
    double beta = so_kaiser_beta(as);
    for (size_t i = 0; i < len; i++) {
        double t = (double)i - ((double)(len - 1)) / 2.0 + mu;
        double h1 = so_sinc(2.0 * fc * t);
        double h2 = so_window_kaiser(i, len, beta);
        h[i] = h1 * h2;
    }
    return 0;
}
/* firdes/mod.rs:329-364; h must hold 2*semi_length+1 values */
SO_API int so_firdes_notch(size_t m, double f0, double as, double *h) {
    if (m < 1 || m > 1000) return -1;
    if (!(f0 >= 0.0 && f0 <= 0.5)) return -1;
    if (as <= 0.0) return -1;
    double beta = so_kaiser_beta(as);
    size_t n = 2 * m + 1;
    double scale = 0.0;
    for (size_t i = 0; i < n; i++) {
        double tone = -cos(2.0 * M_PI * f0 * ((double)i - (double)m));
        double w = so_window_kaiser(i, n, beta);
        h[i] = tone * w;
        scale += h[i] * tone;
    }
    for (size_t i = 0; i < n; i++) h[i] /= scale;
    h[m] += 1.0;
    return 0;
}
/* firdes/mod.rs:443-456 */
SO_API double so_filter_autocorrelation(const double *h, size_t n, ptrdiff_t lag) {
    size_t l = (size_t)(lag < 0 ? -lag : lag);
    if (l >= n) return 0.0;
    double r = 0.0;
    for (size_t i = l; i < n; i++) r += h[i] * h[i - l];
    return r;
}
/* firdes/mod.rs:487-526 */
SO_API double so_filter_crosscorrelation(const double *h, size_t nh, const double *g, size_t ng,
                                         ptrdiff_t lag) {
    if (nh < ng) return so_filter_crosscorrelation(g, ng, h, nh, lag);
    if (lag <= -(ptrdiff_t)ng) return 0.0;
    if (lag >= (ptrdiff_t)nh) return 0.0;
    size_t ig = 0, ih = 0;
    if (lag < 0) ig = (size_t)(-lag);
    if (lag > 0) ih = (size_t)lag;
    ptrdiff_t n;
    if (lag < 0) n = (ptrdiff_t)ng + lag;
    else if (lag < (ptrdiff_t)(nh - ng)) n = (ptrdiff_t)ng;
    else n = (ptrdiff_t)nh - lag;
    double r = 0.0;
    for (size_t i = 0; i < (size_t)n; i++) r += h[ih + i] * g[ig + i];
    return r;
}
/* filter_isi -- firdes/mod.rs:553-573: (rms, max) of |r(i * sps) / r(0)|, i = 1 .. 2 * delay - 1 */
SO_API void so_filter_isi(const double *h, size_t n, size_t sps, size_t delay, double *out2) {
    out2[0] = out2[1] = 0.0;
    if (2 * sps * delay + 1 != n) return;                       /* :554-561 */
    double rxx0 = so_filter_autocorrelation(h, n, 0);
    double isi_rms = 0.0, isi_max = 0.0;
    for (size_t i = 1; i < 2 * delay; i++) {
        double e = fabs(so_filter_autocorrelation(h, n, (ptrdiff_t)(i * sps)) / rxx0);
        isi_rms += e * e;
        if (i == 1 || e > isi_max) isi_max = e;
    }
    out2[0] = sqrt(isi_rms / (2.0 * (double)delay));
    out2[1] = isi_max;
}
/* filter_energy -- firdes/mod.rs:603-640: the crate's own caller of DotProduct::execute (FORWARD, f64 coefficients on
 * complex samples e^{j 2 pi f k}); relative energy above the cut-off.  Returns -1 / -2 / -3 for the Bandwidth /
 * FilterSize / FFTSize errors (:608-614). */
SO_API int so_filter_energy(const double *h, size_t n, double fc, size_t fft_size, double *out) {
    if (!(fc >= 0.0 && fc <= 0.5)) return -1;
    if (n == 0) return -2;
    if (fft_size == 0) return -3;
    cplx *ejwt = (cplx *)calloc(n, sizeof(cplx));
    so_dotprod dp;
    dp_init(&dp, h, n, 0, SO_FORWARD);
    double e_total = 0.0, e_stop = 0.0;
    for (size_t i = 0; i < fft_size; i++) {
        double f = 0.5 * (double)i / (double)fft_size;
        for (size_t k = 0; k < n; k++) {                        /* Complex::from_polar(1.0, theta) */
            double th = 2.0 * M_PI * f * (double)k;
            ejwt[k].re = 1.0 * cos(th);
            ejwt[k].im = 1.0 * sin(th);
        }
        cplx v = dp_execute(&dp, ejwt, n);
        double e2 = v.re * v.re - v.im * (-v.im);               /* (v * v.conj()).re */
        e_total += e2;
        if (f > fc) e_stop += e2;
    }
    dp_free(&dp);
    free(ejwt);
    *out = e_stop / e_total;
    return 0;
}
/* iirdes/pll/mod.rs:24-52; num/den each 3 values.  Returns -1 on a rejected argument. */
SO_API int so_pll_active_lag(double w, double zeta, double k, double *num, double *den) {
    if (w <= 0.0 || zeta <= 0.0 || k <= 0.0) return -1;
    double t1 = k / (w * w);
    double t2 = 2.0 * zeta / w - 1.0 / k;
    num[0] = 2.0 * k * (1.0 + t2 / 2.0);
    num[1] = 2.0 * k * 2.0;
    num[2] = 2.0 * k * (1.0 - t2 / 2.0);
    den[0] = 1.0 + t1 / 2.0;
    den[1] = -t1;
    den[2] = -1.0 + t1 / 2.0;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* NCO -- nco/mod.rs.  PARITY UNPINNED: the reference holds no doc-test or golden for the   */
/* NCO (nco/mod.rs has none; main.rs:29-36 only drives sincos/step), so this follows the    */
/* source literally.  mix_up_block / mix_down_block (:153-172) index an empty Vec and panic  */
/* in the reference; so_nco_mix_block is the loop they were written to be.                   */
typedef struct {
    double table[1024];
    uint32_t theta, delta_theta;
} so_nco;

/* nco/mod.rs:176-188 */
SO_API uint32_t so_nco_constrain(double theta) {
    double t = theta / (2.0 * 3.14159265358979323846264338327950288);
    double ip;
    double frac = modf(t, &ip);                 /* f64::fract */
    if (frac < 0.0) frac += 1.0;
    double v = frac * (double)0xffffffffu;      /* `0xffffffffu32 as f64` binds tighter than `*` */
    if (!(v > 0.0)) return 0;                   /* Rust `as u32` saturates; NaN -> 0 */
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
SO_API so_nco *so_nco_new(void) {              /* :36-50 */
    so_nco *n = (so_nco *)calloc(1, sizeof(so_nco));
    if (!n) return NULL;
    for (int i = 0; i < 1024; i++)
        n->table[i] = sin(2.0 * 3.14159265358979323846264338327950288 * (double)i / 1024.0);
    return n;
}
SO_API void so_nco_free(so_nco *n) { free(n); }
SO_API void so_nco_reset(so_nco *n) { n->theta = 0; n->delta_theta = 0; }                                  /* :53-56 */
SO_API void so_nco_set_frequency(so_nco *n, double dt) { n->delta_theta = so_nco_constrain(dt); }         /* :59-61 */
SO_API void so_nco_adjust_frequency(so_nco *n, double dt) { n->delta_theta += so_nco_constrain(dt); }     /* :64-66 */
SO_API void so_nco_set_phase(so_nco *n, double phi) { n->theta = so_nco_constrain(phi); }                 /* :79-81 */
SO_API void so_nco_adjust_phase(so_nco *n, double dphi) { n->theta += so_nco_constrain(dphi); }           /* :84-86 */
SO_API void so_nco_step(so_nco *n) { n->theta += n->delta_theta; }                                        /* :93-96 wrapping_add */
SO_API void so_nco_set_raw(so_nco *n, uint32_t theta, uint32_t delta) { n->theta = theta; n->delta_theta = delta; }
SO_API uint32_t so_nco_theta(const so_nco *n) { return n->theta; }
SO_API uint32_t so_nco_delta_theta(const so_nco *n) { return n->delta_theta; }
static inline size_t nco_index(const so_nco *n) {                                                          /* :98-101 */
    return (size_t)(((uint32_t)(n->theta + (1u << 21)) >> 22) & 0x3ff);
}
SO_API double so_nco_sin(const so_nco *n) { return n->table[nco_index(n)]; }                               /* :103-106 */
SO_API double so_nco_cos(const so_nco *n) { return n->table[(nco_index(n) + 256) & 0x3ff]; }               /* :108-112 */
SO_API void so_nco_mix(const so_nco *n, int up, double re, double im, double *out2) {                      /* :141-151 */
    cplx ph = { so_nco_cos(n), so_nco_sin(n) };   /* complex_exponential, :119-121 */
    if (!up) ph.im = -ph.im;                      /* .conj() */
    cplx x = { re, im };
    cplx y = c_mul(ph, x);
    out2[0] = y.re; out2[1] = y.im;
}
SO_API void so_nco_mix_block(so_nco *n, int up, const double *x, size_t len, double *out) {
    for (size_t i = 0; i < len; i++) {
        so_nco_mix(n, up, x[2 * i], x[2 * i + 1], out + 2 * i);
        so_nco_step(n);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Threaded drivers for the timed baseline: one independent filter OBJECT per channel (or
 * per stream segment), one pthread per worker -- what a user of the (single-threaded,
 * !Send) reference could do with its public API by giving every thread its own object.  */
typedef struct {
    int kind; /* 0 fir/decim, 1 interp, 2 iir sos */
    const double *coefs; size_t ncoef; int coef_complex; double scale_re, scale_im;
    int is_decim; size_t factor;
    const double *ff, *fb; size_t nco;
    const double *x; size_t in_stride, n_in; /* per-unit input (complex samples) */
    size_t preroll;                        /* inputs whose outputs are discarded (segment priming) */
    double *out; size_t out_stride;
    size_t unit_begin, unit_end;
    size_t n_out_last;
} so_job;

static void *so_worker(void *arg) {
    so_job *j = (so_job *)arg;
    int err = 0;
    for (size_t u = j->unit_begin; u < j->unit_end; u++) {
        const double *x = j->x + 2 * u * j->in_stride;
        double *out = j->out + 2 * u * j->out_stride;
        if (j->kind == 0) {
            so_fir *f = so_fir_new(j->coefs, j->ncoef, j->coef_complex, j->scale_re, j->scale_im,
                                   j->is_decim, j->factor, &err);
            if (!f) return NULL;
            if (j->preroll) { /* prime history through the public API, drop those outputs */
                double *tmp = (double *)malloc(2 * j->preroll * sizeof(double));
                so_fir_execute_block(f, x - 2 * j->preroll, j->preroll, tmp, j->preroll);
                free(tmp);
            }
            j->n_out_last = so_fir_execute_block(f, x, j->n_in, out, j->out_stride);
            so_fir_free(f);
        } else if (j->kind == 1) {
            so_firinterp *f = so_firinterp_new(j->coefs, j->ncoef, j->coef_complex, j->factor, &err);
            if (!f) return NULL;
            j->n_out_last = so_firinterp_execute_block(f, x, j->n_in, out, j->out_stride);
            so_firinterp_free(f);
        } else {
            so_iir *f = so_iir_new(j->ff, j->nco, j->fb, j->nco, 1, 0, 0, &err);
            if (!f) return NULL;
            j->n_out_last = so_iir_execute_block(f, x, j->n_in, out, j->out_stride);
            so_iir_free(f);
        }
    }
    return NULL;
}

/* Run `n_units` independent objects (channels or stream segments) on `n_threads` threads.
 * Unit u reads x + u*in_stride (complex samples) and writes out + u*out_stride.  For
 * stream segments pass preroll = T-1 for every unit except that unit 0 must have `preroll`
 * readable samples before it too (callers pass a pointer offset by preroll).  Returns the
 * per-unit output count.                                                                   */
SO_API size_t so_run_units(int kind, const double *coefs, size_t ncoef, int coef_complex,
                           double scale_re, double scale_im, int is_decim, size_t factor,
                           const double *ff, const double *fb, size_t nco,
                           const double *x, size_t in_stride, size_t n_in, size_t preroll,
                           double *out, size_t out_stride, size_t n_units, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_units) n_threads = (int)n_units;
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
    so_job *jobs = (so_job *)calloc(n_threads, sizeof(so_job));
    size_t per = (n_units + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        so_job *j = &jobs[t];
        j->kind = kind; j->coefs = coefs; j->ncoef = ncoef; j->coef_complex = coef_complex;
        j->scale_re = scale_re; j->scale_im = scale_im; j->is_decim = is_decim; j->factor = factor;
        j->ff = ff; j->fb = fb; j->nco = nco;
        j->x = x; j->in_stride = in_stride; j->n_in = n_in; j->preroll = preroll;
        j->out = out; j->out_stride = out_stride;
        j->unit_begin = (size_t)t * per;
        j->unit_end = j->unit_begin + per < n_units ? j->unit_begin + per : n_units;
        if (j->unit_begin > n_units) j->unit_begin = n_units;
        if (n_threads == 1) so_worker(j);
        else pthread_create(&th[t], NULL, so_worker, j);
    }
    if (n_threads > 1) for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    size_t r = jobs[0].n_out_last;
    free(jobs); free(th);
    return r;
}
