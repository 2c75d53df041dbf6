//! The two type parameters of the reference's filters, as sealed traits.
use num::complex::Complex;
use solid_gpu_sys as sys;

mod sealed {
    pub trait Sealed {}
    impl Sealed for f64 {}
    impl Sealed for num::complex::Complex<f64> {}
    impl Sealed for num::complex::Complex<f32> {}
}

/// `Coef` of `FIRFilter<Coef, In>` / `IIRFilter<Coef, In>`: `f64` or `Complex<f64>` (fir/mod.rs:181-186).
pub trait Coefficient: Copy + sealed::Sealed + 'static {
    /// sgpu_tapkind
    const KIND: i32;
    /// doubles per coefficient in the ABI's flat layout
    const WIDTH: usize;
    fn flatten(coefs: &[Self]) -> Vec<f64>;
    fn unflatten(flat: &[f64]) -> Vec<Self>;
    fn to_complex(self) -> Complex<f64>;
}
impl Coefficient for f64 {
    const KIND: i32 = sys::SGPU_TAPS_REAL;
    const WIDTH: usize = 1;
    fn flatten(coefs: &[Self]) -> Vec<f64> { coefs.to_vec() }
    fn unflatten(flat: &[f64]) -> Vec<Self> { flat.to_vec() }
    fn to_complex(self) -> Complex<f64> { Complex::new(self, 0.0) }
}
impl Coefficient for Complex<f64> {
    const KIND: i32 = sys::SGPU_TAPS_COMPLEX;
    const WIDTH: usize = 2;
    fn flatten(coefs: &[Self]) -> Vec<f64> { coefs.iter().flat_map(|c| [c.re, c.im]).collect() }
    fn unflatten(flat: &[f64]) -> Vec<Self> { flat.chunks_exact(2).map(|p| Complex::new(p[0], p[1])).collect() }
    fn to_complex(self) -> Complex<f64> { self }
}

/// `In` (= `Out`) of the filters: complex samples.  The device computes in `Complex<f32>`.
pub trait Sample: Copy + sealed::Sealed + 'static {
    /// the samples as the device's cf32 (borrowed when they already are)
    fn narrow(samples: &[Self]) -> std::borrow::Cow<'_, [Complex<f32>]>;
    fn widen(samples: Vec<Complex<f32>>) -> Vec<Self>;
    fn from_cf32(v: Complex<f32>) -> Self;
}
impl Sample for Complex<f32> {
    fn narrow(samples: &[Self]) -> std::borrow::Cow<'_, [Complex<f32>]> { std::borrow::Cow::Borrowed(samples) }
    fn widen(samples: Vec<Complex<f32>>) -> Vec<Self> { samples }
    fn from_cf32(v: Complex<f32>) -> Self { v }
}
impl Sample for Complex<f64> {
    fn narrow(samples: &[Self]) -> std::borrow::Cow<'_, [Complex<f32>]> {
        std::borrow::Cow::Owned(samples.iter().map(|s| Complex::new(s.re as f32, s.im as f32)).collect())
    }
    fn widen(samples: Vec<Complex<f32>>) -> Vec<Self> {
        samples.into_iter().map(|s| Complex::new(s.re as f64, s.im as f64)).collect()
    }
    fn from_cf32(v: Complex<f32>) -> Self { Complex::new(v.re as f64, v.im as f64) }
}
