//! reference: src/filter/fir/mod.rs, decim.rs, interp.rs, pfb.rs
use super::Filter;
use crate::group_delay::fir_group_delay;
use crate::scalar::{Coefficient, Sample};
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::marker::PhantomData;
use std::{fmt, ptr};

/// fir/mod.rs:39-45
#[derive(Debug, PartialEq, Eq)]
pub enum FIRErrorCode { CoefficientsLengthZero, DecimationLessThanOne, InterpolationLessThanOne, NotEnoughFilters }
/// fir/mod.rs:47-56
#[derive(Debug)]
pub struct FIRError(pub FIRErrorCode);
impl fmt::Display for FIRError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "FIR Filter Error {:?}", self.0) }
}
impl Error for FIRError {}

pub(crate) fn ctor_error(st: i32) -> Box<dyn Error> {
    match st {
        sys::SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO => Box::new(FIRError(FIRErrorCode::CoefficientsLengthZero)),
        sys::SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE => Box::new(FIRError(FIRErrorCode::DecimationLessThanOne)),
        sys::SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE => Box::new(FIRError(FIRErrorCode::InterpolationLessThanOne)),
        sys::SGPU_ERR_FIR_NOT_ENOUGH_FILTERS => Box::new(FIRError(FIRErrorCode::NotEnoughFilters)),
        _ => crate::last_error().into(),
    }
}

/// Σ c_i e^{j 2π f i} over the STORED coefficient order (fir/mod.rs:263-273)
fn response<C: Coefficient>(coefs: &[C], frequency: f64) -> Complex<f64> {
    coefs.iter().enumerate().fold(Complex::new(0.0, 0.0), |acc, (i, c)| {
        acc + c.to_complex() * Complex::from_polar(1.0, frequency * 2.0 * std::f64::consts::PI * (i as f64))
    })
}

/// The scale is a `Coef` in the reference (fir/mod.rs:61); it crosses the ABI as (re, im).
fn scale_parts<C: Coefficient>(scale: C) -> (f64, f64) { let s = scale.to_complex(); (s.re, s.im) }
fn scale_from<C: Coefficient>(re: f64, im: f64) -> C { C::unflatten(&[re, im][..C::WIDTH])[0] }

/// FIRFilter<Coef, In> -- fir/mod.rs:58-63.  y[n] = scale * Σ_i h[T-1-i] x[n-i]
pub struct FIRFilter<Coef: Coefficient, In: Sample> { pub(crate) h: *mut sys::sgpu_fir, _p: PhantomData<(Coef, In)> }

impl<Coef: Coefficient, In: Sample> FIRFilter<Coef, In> {
    pub(crate) fn create(coefficents: &[Coef], scale: Coef, is_decim: bool, decimation: usize) -> Result<Self, Box<dyn Error>> {
        let flat = Coef::flatten(coefficents);
        let (re, im) = scale_parts(scale);
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_fir_create(flat.as_ptr(), coefficents.len(), Coef::KIND, 1, re, im, is_decim as i32, decimation, &mut h) };
        if st != sys::SGPU_OK { return Err(ctor_error(st)); }
        Ok(FIRFilter { h, _p: PhantomData })
    }
    /// fir/mod.rs:79
    pub fn new(coefficents: &[Coef], scale: Coef) -> Result<Self, Box<dyn Error>> { Self::create(coefficents, scale, false, 0) }
    /// fir/mod.rs:106
    pub fn set_scale(&mut self, scale: Coef) { let (re, im) = scale_parts(scale); unsafe { sys::sgpu_fir_set_scale(self.h, re, im) }; }
    /// fir/mod.rs:124
    pub fn get_scale(&self) -> Coef { let (mut re, mut im) = (0.0, 0.0); unsafe { sys::sgpu_fir_get_scale(self.h, &mut re, &mut im) }; scale_from(re, im) }
    /// fir/mod.rs:142
    pub fn len(&self) -> usize { unsafe { sys::sgpu_fir_len(self.h) } }
    /// fir/mod.rs:158
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    /// fir/mod.rs:176 -- the stored (reversed) order
    pub fn coefficients(&self) -> Vec<Coef> {
        let mut flat = vec![0.0; self.len() * Coef::WIDTH];
        unsafe { sys::sgpu_fir_coefficients(self.h, flat.as_mut_ptr()) };
        Coef::unflatten(&flat)
    }
    /// true when the last execute_block ran on the tcgen05 tensor-core kernel (not part of the reference API)
    pub fn last_path_tensor(&self) -> bool { unsafe { sys::sgpu_fir_last_path(self.h) == 1 } }
    pub(crate) fn run(&mut self, samples: &[In]) -> Vec<In> {
        let x = In::narrow(samples);
        let n_out = unsafe { sys::sgpu_fir_out_len(self.h, x.len()) };
        let mut out = vec![Complex::new(0f32, 0f32); n_out];
        let mut got = 0usize;
        let st = unsafe {
            sys::sgpu_fir_execute_block(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1),
                                        out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_fir_execute_block");
        In::widen(out)
    }
    pub(crate) fn push_block(&mut self, samples: &[In]) {
        let x = In::narrow(samples);
        let st = unsafe { sys::sgpu_fir_write(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), sys::SGPU_HOST, ptr::null_mut()) };
        crate::expect_ok(st, "sgpu_fir_write");
    }
}
impl<Coef: Coefficient, In: Sample> Clone for FIRFilter<Coef, In> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_fir_clone(self.h, &mut h) }, "sgpu_fir_clone");
        FIRFilter { h, _p: PhantomData }
    }
}
impl<Coef: Coefficient, In: Sample> Drop for FIRFilter<Coef, In> { fn drop(&mut self) { unsafe { sys::sgpu_fir_destroy(self.h) }; } }
impl<Coef: Coefficient, In: Sample> fmt::Debug for FIRFilter<Coef, In> {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { fmt::Display::fmt(self, f) }
}
impl<Coef: Coefficient, In: Sample> fmt::Display for FIRFilter<Coef, In> {
    /// fir/mod.rs:306-316
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "FIR<{}> [Coefficients=DotProduct [Size={}]]", std::any::type_name::<Coef>(), self.len())
    }
}
impl<Coef: Coefficient, In: Sample> Filter<In, In> for FIRFilter<Coef, In> {
    /// fir/mod.rs:209
    fn execute(&mut self, sample: In) -> Vec<In> { self.run(&[sample]) }
    /// fir/mod.rs:235
    fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.run(samples) }
    /// fir/mod.rs:263-273
    fn frequency_response(&self, frequency: f64) -> Complex<f64> { response(&self.coefficients(), frequency) * self.get_scale().to_complex() }
    /// fir/mod.rs:293-303: 0.0 when fir_group_delay errors
    fn group_delay(&self, frequency: f64) -> f64 { fir_group_delay(&self.coefficients(), frequency).unwrap_or(0.0) }
}

pub mod decim {
    use super::*;
    /// DecimatingFIRFilter<Coef, In> -- fir/decim.rs:5-10: emits when (count + 1) % M == 0
    #[derive(Clone, Debug)]
    pub struct DecimatingFIRFilter<Coef: Coefficient, In: Sample> { filter: FIRFilter<Coef, In> }
    impl<Coef: Coefficient, In: Sample> DecimatingFIRFilter<Coef, In> {
        /// fir/decim.rs:27
        pub fn new(coefficents: &[Coef], scale: Coef, decimation: usize) -> Result<Self, Box<dyn Error>> {
            Ok(DecimatingFIRFilter { filter: FIRFilter::create(coefficents, scale, true, decimation)? })
        }
        /// decim.rs:60
        pub fn set_scale(&mut self, scale: Coef) { self.filter.set_scale(scale) }
        /// decim.rs:78
        pub fn get_scale(&self) -> Coef { self.filter.get_scale() }
        /// decim.rs:96
        pub fn get_decimation(&self) -> usize { unsafe { sys::sgpu_fir_decimation(self.filter.h) } }
        /// decim.rs:115
        pub fn push(&mut self, sample: In) { self.filter.push_block(&[sample]) }
        /// decim.rs:136 -- pushes without producing output; the counter advances
        pub fn write(&mut self, samples: &[In]) { self.filter.push_block(samples) }
        /// decim.rs:156
        pub fn len(&self) -> usize { self.filter.len() }
        /// decim.rs:172
        pub fn is_empty(&self) -> bool { self.filter.is_empty() }
        /// decim.rs:190
        pub fn coefficients(&self) -> Vec<Coef> { self.filter.coefficients() }
    }
    impl<Coef: Coefficient, In: Sample> fmt::Display for DecimatingFIRFilter<Coef, In> {
        /// decim.rs:281-295
        fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "{} [Decimation={}]", self.filter, self.get_decimation()) }
    }
    impl<Coef: Coefficient, In: Sample> Filter<In, In> for DecimatingFIRFilter<Coef, In> {
        /// decim.rs:221
        fn execute(&mut self, sample: In) -> Vec<In> { self.filter.run(&[sample]) }
        /// decim.rs:250
        fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.filter.run(samples) }
        fn frequency_response(&self, frequency: f64) -> Complex<f64> { self.filter.frequency_response(frequency) }
        fn group_delay(&self, frequency: f64) -> f64 { self.filter.group_delay(frequency) }
    }
}

/// One handle type behind both the interpolator and the bare filter bank.
struct Bank { h: *mut sys::sgpu_interp }
impl Bank {
    fn filters(&self) -> usize { unsafe { sys::sgpu_interp_interpolation(self.h) } }
    fn sub_len(&self) -> usize { unsafe { sys::sgpu_interp_sub_len(self.h) } }
    fn flat_coefs(&self, width: usize) -> Vec<f64> {
        let mut v = vec![0.0; self.filters() * self.sub_len() * width];
        if !v.is_empty() { unsafe { sys::sgpu_interp_coefficients(self.h, v.as_mut_ptr()) }; }
        v
    }
}
impl Clone for Bank {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_interp_clone(self.h, &mut h) }, "sgpu_interp_clone");
        Bank { h }
    }
}
impl Drop for Bank { fn drop(&mut self) { unsafe { sys::sgpu_interp_destroy(self.h) }; } }

pub mod pfb {
    use super::*;
    /// PolyPhaseFilterBank<Coef, In> -- fir/pfb.rs:3-8
    #[derive(Clone)]
    pub struct PolyPhaseFilterBank<Coef: Coefficient, In: Sample> { bank: Bank, _p: PhantomData<(Coef, In)> }
    impl<Coef: Coefficient, In: Sample> PolyPhaseFilterBank<Coef, In> {
        /// fir/pfb.rs:24 (sub_len = len / filters truncates, :32; filters > len is NotEnoughFilters instead of the reference's panic)
        pub fn new(coefficients: &[Coef], filters: usize, scale: Coef) -> Result<Self, Box<dyn Error>> {
            let flat = Coef::flatten(coefficients);
            let (re, im) = scale_parts(scale);
            let mut h = ptr::null_mut();
            let st = unsafe { sys::sgpu_pfb_create(flat.as_ptr(), coefficients.len(), Coef::KIND, 1, filters, re, im, &mut h) };
            if st != sys::SGPU_OK { return Err(ctor_error(st)); }
            Ok(PolyPhaseFilterBank { bank: Bank { h }, _p: PhantomData })
        }
        /// pfb.rs:52 -- stored, never applied (pfb.rs:85-90)
        pub fn set_scale(&mut self, scale: Coef) { let (re, im) = scale_parts(scale); unsafe { sys::sgpu_interp_set_scale(self.bank.h, re, im) }; }
        /// pfb.rs:57
        pub fn get_scale(&self) -> Coef { let (mut re, mut im) = (0.0, 0.0); unsafe { sys::sgpu_interp_get_scale(self.bank.h, &mut re, &mut im) }; scale_from(re, im) }
        /// pfb.rs:62 -- the number of sub-filters
        pub fn len(&self) -> usize { self.bank.filters() }
        /// pfb.rs:67
        pub fn is_empty(&self) -> bool { self.len() == 0 }
        /// [sic] pfb.rs:71
        pub fn coefficents(&self) -> Vec<Vec<Coef>> {
            let flat = Coef::unflatten(&self.bank.flat_coefs(Coef::WIDTH));
            flat.chunks(self.bank.sub_len().max(1)).map(|c| c.to_vec()).collect()
        }
        /// pfb.rs:76
        pub fn reset(&mut self) { unsafe { sys::sgpu_interp_reset(self.bank.h) }; }
        /// pfb.rs:81
        pub fn push(&mut self, sample: In) {
            let x = In::narrow(std::slice::from_ref(&sample));
            let st = unsafe { sys::sgpu_interp_push(self.bank.h, x.as_ptr() as *const f32, 1, 1, sys::SGPU_HOST, ptr::null_mut()) };
            crate::expect_ok(st, "sgpu_interp_push");
        }
        /// pfb.rs:85-90 -- Σ_j h[p + (S-1-j) L] x[n-j], no scale
        pub fn execute(&self, index: usize) -> In {
            let mut r = Complex::new(0f32, 0f32);
            let st = unsafe { sys::sgpu_interp_execute_phase(self.bank.h, index, &mut r as *mut Complex<f32> as *mut f32, sys::SGPU_HOST, ptr::null_mut()) };
            crate::expect_ok(st, "sgpu_interp_execute_phase");
            In::from_cf32(r)
        }
    }
}

pub mod interp {
    use super::*;
    /// InterpolatingFIRFilter<Coef, In> -- fir/interp.rs:6-10: y[nL + p], p = 0..L-1, no scale
    #[derive(Clone)]
    pub struct InterpolatingFIRFilter<Coef: Coefficient, In: Sample> { bank: Bank, _p: PhantomData<(Coef, In)> }
    impl<Coef: Coefficient, In: Sample> InterpolatingFIRFilter<Coef, In> {
        /// fir/interp.rs:27
        pub fn new(coefficents: &[Coef], interpolation: usize) -> Result<Self, Box<dyn Error>> {
            let flat = Coef::flatten(coefficents);
            let mut h = ptr::null_mut();
            let st = unsafe { sys::sgpu_interp_create(flat.as_ptr(), coefficents.len(), Coef::KIND, 1, interpolation, &mut h) };
            if st != sys::SGPU_OK { return Err(ctor_error(st)); }
            Ok(InterpolatingFIRFilter { bank: Bank { h }, _p: PhantomData })
        }
        /// interp.rs:57
        pub fn set_scale(&mut self, scale: Coef) { let (re, im) = scale_parts(scale); unsafe { sys::sgpu_interp_set_scale(self.bank.h, re, im) }; }
        /// interp.rs:62
        pub fn get_scale(&self) -> Coef { let (mut re, mut im) = (0.0, 0.0); unsafe { sys::sgpu_interp_get_scale(self.bank.h, &mut re, &mut im) }; scale_from(re, im) }
        /// interp.rs:67 -- the number of sub-filters (= L)
        pub fn len(&self) -> usize { self.bank.filters() }
        /// interp.rs:72
        pub fn is_empty(&self) -> bool { self.len() == 0 }
        /// [sic] interp.rs:77 -- flattened
        pub fn coefficents(&self) -> Vec<Coef> { Coef::unflatten(&self.bank.flat_coefs(Coef::WIDTH)) }
        /// interp.rs:82
        pub fn interpolation(&self) -> usize { self.len() }
        pub fn last_path_tensor(&self) -> bool { unsafe { sys::sgpu_interp_last_path(self.bank.h) == 1 } }
        fn run(&mut self, samples: &[In]) -> Vec<In> {
            let x = In::narrow(samples);
            let n_out = x.len() * self.len();
            let mut out = vec![Complex::new(0f32, 0f32); n_out];
            let mut got = 0usize;
            let st = unsafe {
                sys::sgpu_interp_execute_block(self.bank.h, x.as_ptr() as *const f32, x.len(), x.len().max(1),
                                               out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
            };
            crate::expect_ok(st, "sgpu_interp_execute_block");
            In::widen(out)
        }
    }
    impl<Coef: Coefficient, In: Sample> fmt::Debug for InterpolatingFIRFilter<Coef, In> {
        fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { fmt::Display::fmt(self, f) }
    }
    impl<Coef: Coefficient, In: Sample> fmt::Display for InterpolatingFIRFilter<Coef, In> {
        fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
            write!(f, "InterpolatingFIR<{}> [Interpolation={}] [SubLen={}]", std::any::type_name::<Coef>(), self.len(), self.bank.sub_len())
        }
    }
    impl<Coef: Coefficient, In: Sample> Filter<In, In> for InterpolatingFIRFilter<Coef, In> {
        /// interp.rs:93
        fn execute(&mut self, sample: In) -> Vec<In> { self.run(&[sample]) }
        /// interp.rs:102
        fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.run(samples) }
        /// interp.rs:113-123
        fn frequency_response(&self, frequency: f64) -> Complex<f64> { response(&self.coefficents(), frequency) * self.get_scale().to_complex() }
        /// interp.rs:125-136
        fn group_delay(&self, frequency: f64) -> f64 { fir_group_delay(&self.coefficents(), frequency).unwrap_or(0.0) }
    }
}
