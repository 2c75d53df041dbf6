//! reference: src/filter/iir/mod.rs, sos.rs, decim.rs, interp.rs
use super::Filter;
use crate::group_delay::iir_group_delay;
use crate::scalar::Sample;
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::marker::PhantomData;
use std::{fmt, ptr};

/// iir/mod.rs:40-49
#[derive(Debug, PartialEq, Eq)]
pub enum IIRErrorCode {
    NumeratorLengthZero, DenominatorLengthZero, SecondOrderSectionSizeZero, SecondOrderSectionSizeMismatch,
    SecondOrderSectionSizeNotMultpleOf3, DecimationLessThanOne, InterpolationLessThanOne,
}
/// iir/mod.rs:51-60
#[derive(Debug)]
pub struct IIRError(pub IIRErrorCode);
impl fmt::Display for IIRError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "IIR Filter Error {:?}", self.0) }
}
impl Error for IIRError {}
/// iir/mod.rs:62-66
#[derive(PartialEq, Eq, Debug, Clone, Copy)]
pub enum IIRFilterType { Normal, SecondOrder }

fn ctor_error(st: i32) -> Box<dyn Error> {
    use IIRErrorCode::*;
    let code = match st {
        sys::SGPU_ERR_IIR_NUMERATOR_LENGTH_ZERO => NumeratorLengthZero,
        sys::SGPU_ERR_IIR_DENOMINATOR_LENGTH_ZERO => DenominatorLengthZero,
        sys::SGPU_ERR_IIR_SOS_SIZE_ZERO => SecondOrderSectionSizeZero,
        sys::SGPU_ERR_IIR_SOS_SIZE_MISMATCH => SecondOrderSectionSizeMismatch,
        sys::SGPU_ERR_IIR_SOS_SIZE_NOT_MULTIPLE_OF_3 => SecondOrderSectionSizeNotMultpleOf3,
        sys::SGPU_ERR_IIR_DECIMATION_LESS_THAN_ONE => DecimationLessThanOne,
        sys::SGPU_ERR_IIR_INTERPOLATION_LESS_THAN_ONE => InterpolationLessThanOne,
        sys::SGPU_ERR_SOS_COEFFICIENTS_NOT_IN_RANGE => return Box::new(sos::SecondOrderError(sos::SecondOrderErrorCode::CoefficientsNotInRange)),
        _ => return crate::last_error().into(),
    };
    Box::new(IIRError(code))
}

fn poly(coefs: &[f64], frequency: f64) -> Complex<f64> {
    coefs.iter().enumerate().fold(Complex::new(0.0, 0.0), |acc, (i, c)| {
        acc + Complex::from_polar(1.0, frequency * 2.0 * std::f64::consts::PI * (i as f64)) * *c
    })
}

pub mod sos {
    use super::*;
    /// sos.rs:18-32
    #[derive(Debug, PartialEq, Eq)]
    pub enum SecondOrderErrorCode { CoefficientsNotInRange }
    #[derive(Debug)]
    pub struct SecondOrderError(pub SecondOrderErrorCode);
    impl fmt::Display for SecondOrderError {
        fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "Second Order Filter Error {:?}", self.0) }
    }
    impl Error for SecondOrderError {}

    /// SecondOrderFilter<C, T> -- sos.rs:34-39: one direct-form-II biquad; a one-section handle underneath.
    /// `C` is `f64` (the reference's only instantiation through IIRFilter, iir/mod.rs:144-153).
    #[derive(Clone, Debug)]
    pub struct SecondOrderFilter<T: Sample> { ff: [f64; 3], fb: [f64; 3], filter: IIRFilter<f64, T> }
    impl<T: Sample> SecondOrderFilter<T> {
        /// sos.rs:55 (fewer than three values on either side: CoefficientsNotInRange, :56-60)
        pub fn new(feed_forward: &[f64], feed_back: &[f64]) -> Result<Self, Box<dyn Error>> {
            if feed_forward.len() < 3 || feed_back.len() < 3 {
                return Err(Box::new(SecondOrderError(SecondOrderErrorCode::CoefficientsNotInRange)));
            }
            let (ff, fb) = ([feed_forward[0], feed_forward[1], feed_forward[2]], [feed_back[0], feed_back[1], feed_back[2]]);
            Ok(SecondOrderFilter { ff, fb, filter: IIRFilter::new(&ff, &fb, IIRFilterType::SecondOrder)? })
        }
        /// sos.rs:92-114: v0 = x - (a1 v1 + a2 v2); y = b0 v0 + b1 v1 + b2 v2 (Either<T, Out> collapses to one complex sample)
        pub fn execute(&mut self, sample: T) -> T { self.filter.run(&[sample])[0] }
        /// sos.rs:116 -- holds a1, a2 (the field names are swapped in the reference, sos.rs:70-75)
        pub fn numerator_coefs(&self) -> Vec<f64> { vec![self.fb[1] / self.fb[0], self.fb[2] / self.fb[0]] }
        /// sos.rs:136 -- holds b0, b1, b2
        pub fn denominator_coefs(&self) -> Vec<f64> { self.ff.iter().map(|b| b / self.fb[0]).collect() }
        /// sos.rs:151-172
        pub fn frequency_response(&self, frequency: f64) -> Complex<f64> {
            poly(&self.numerator_coefs(), frequency) / poly(&self.denominator_coefs(), frequency)
        }
        /// sos.rs:208-230
        pub fn group_delay(&self, frequency: f64) -> f64 {
            iir_group_delay(&self.numerator_coefs(), &self.denominator_coefs(), frequency).map(|d| d + 2.0).unwrap_or(0.0)
        }
    }
}

/// IIRFilter<Coef, In> -- iir/mod.rs:68-75.  `Coef` is `f64` (the reference's constructors take `&[f64]`-like real
/// coefficient lists; complex IIR coefficients are not reachable through iir/mod.rs:92-164).
pub struct IIRFilter<Coef, In: Sample> {
    h: *mut sys::sgpu_iir,
    iirtype: IIRFilterType,
    sections: Vec<sos::SecondOrderFilter<In>>,
    _p: PhantomData<Coef>,
}
impl<In: Sample> IIRFilter<f64, In> {
    pub(crate) fn create(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, wrap: i32, factor: usize) -> Result<Self, Box<dyn Error>> {
        let mut h = ptr::null_mut();
        let t = if iirtype == IIRFilterType::Normal { sys::SGPU_IIR_NORMAL } else { sys::SGPU_IIR_SECOND_ORDER };
        let st = unsafe { sys::sgpu_iir_create(t, ff.as_ptr(), ff.len(), fb.as_ptr(), fb.len(), 1, wrap, factor, &mut h) };
        if st != sys::SGPU_OK { return Err(ctor_error(st)); }
        Ok(IIRFilter { h, iirtype, sections: Vec::new(), _p: PhantomData })
    }
    /// iir/mod.rs:92
    pub fn new(feed_forward: &[f64], feed_back: &[f64], iirtype: IIRFilterType) -> Result<Self, Box<dyn Error>> {
        Self::create(feed_forward, feed_back, iirtype, sys::SGPU_IIR_PLAIN, 0)
    }
    fn coefs(&self, num: bool) -> Vec<f64> {
        let mut n = 0usize;
        unsafe { if num { sys::sgpu_iir_numerator_coefs(self.h, ptr::null_mut(), &mut n) } else { sys::sgpu_iir_denominator_coefs(self.h, ptr::null_mut(), &mut n) } };
        let mut v = vec![0.0; n];
        unsafe { if num { sys::sgpu_iir_numerator_coefs(self.h, v.as_mut_ptr(), &mut n) } else { sys::sgpu_iir_denominator_coefs(self.h, v.as_mut_ptr(), &mut n) } };
        v
    }
    /// iir/mod.rs:182
    pub fn numerator_coefs(&self) -> Vec<f64> { self.coefs(true) }
    /// iir/mod.rs:202
    pub fn denominator_coefs(&self) -> Vec<f64> { self.coefs(false) }
    /// iir/mod.rs:222 -- one SecondOrderFilter per section (views built on first use; the cascade itself runs in one kernel)
    pub fn second_order_filters(&mut self) -> &Vec<sos::SecondOrderFilter<In>> {
        if self.sections.is_empty() && self.iirtype == IIRFilterType::SecondOrder {
            let (ff, fb) = (self.numerator_coefs(), self.denominator_coefs());
            for i in 0..ff.len() / 3 {
                if let Ok(s) = sos::SecondOrderFilter::new(&ff[3 * i..3 * i + 3], &fb[3 * i..3 * i + 3]) { self.sections.push(s); }
            }
        }
        &self.sections
    }
    /// iir/mod.rs:239
    pub fn iir_type(&self) -> &IIRFilterType { &self.iirtype }
    pub(crate) fn run(&mut self, samples: &[In]) -> Vec<In> {
        let x = In::narrow(samples);
        let n_out = unsafe { sys::sgpu_iir_out_len(self.h, x.len()) };
        let mut out = vec![Complex::new(0f32, 0f32); n_out];
        let mut got = 0usize;
        let st = unsafe {
            sys::sgpu_iir_execute_block(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1),
                                        out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_iir_execute_block");
        In::widen(out)
    }
}
impl<In: Sample> Clone for IIRFilter<f64, In> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_iir_clone(self.h, &mut h) }, "sgpu_iir_clone");
        IIRFilter { h, iirtype: self.iirtype, sections: Vec::new(), _p: PhantomData }
    }
}
impl<Coef, In: Sample> Drop for IIRFilter<Coef, In> { fn drop(&mut self) { unsafe { sys::sgpu_iir_destroy(self.h) }; } }
impl<Coef, In: Sample> fmt::Debug for IIRFilter<Coef, In> {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "IIR [{:?}] [Sections={}]", self.iirtype, unsafe { sys::sgpu_iir_sections(self.h) }) }
}
impl<Coef, In: Sample> fmt::Display for IIRFilter<Coef, In> {
    /// iir/mod.rs:398-410
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { fmt::Debug::fmt(self, f) }
}
impl<In: Sample> Filter<In, In> for IIRFilter<f64, In> {
    /// iir/mod.rs:270
    fn execute(&mut self, sample: In) -> Vec<In> { self.run(&[sample]) }
    /// iir/mod.rs:310
    fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.run(samples) }
    /// iir/mod.rs:336-373.  SecondOrder mode multiplies into a zero-initialised product, so the reference always
    /// returns 0 there (asserted by its doc-test, iir/mod.rs:328-334) -- kept.
    fn frequency_response(&self, frequency: f64) -> Complex<f64> {
        if self.iirtype == IIRFilterType::Normal {
            poly(&self.numerator_coefs(), frequency) / poly(&self.denominator_coefs(), frequency)
        } else {
            Complex::new(0.0, 0.0)
        }
    }
    /// iir/mod.rs:374-395
    fn group_delay(&self, frequency: f64) -> f64 {
        if self.iirtype == IIRFilterType::SecondOrder {
            let (ff, fb) = (self.numerator_coefs(), self.denominator_coefs());
            let mut delay = 0.0;
            for i in 0..ff.len() / 3 {
                let (a0, b) = (fb[3 * i], &ff[3 * i..3 * i + 3]);
                let num = [fb[3 * i + 1] / a0, fb[3 * i + 2] / a0];                 // sos.rs:116 (swapped names)
                let den = [b[0] / a0, b[1] / a0, b[2] / a0];                       // sos.rs:136
                delay = delay + iir_group_delay(&num, &den, frequency).map(|d| d + 2.0).unwrap_or(0.0) + 2.0;
            }
            delay
        } else {
            iir_group_delay(&self.numerator_coefs(), &self.denominator_coefs(), frequency).unwrap_or(0.0)
        }
    }
}

pub mod decim {
    use super::*;
    /// DecimatingIIRFilter -- iir/decim.rs:5: the recurrence runs every sample, outputs are kept where (index + 1) % M == 0
    #[derive(Clone, Debug)]
    pub struct DecimatingIIRFilter<Coef, In: Sample> where IIRFilter<Coef, In>: Clone { filter: IIRFilter<Coef, In>, decimation: usize }
    impl<In: Sample> DecimatingIIRFilter<f64, In> {
        /// iir/decim.rs:30
        pub fn new(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, decimation: usize) -> Result<Self, Box<dyn Error>> {
            Ok(DecimatingIIRFilter { filter: IIRFilter::create(ff, fb, iirtype, sys::SGPU_IIR_DECIMATING, decimation)?, decimation })
        }
        /// iir/decim.rs:64
        pub fn get_decimation(&self) -> usize { self.decimation }
        pub fn numerator_coefs(&self) -> Vec<f64> { self.filter.numerator_coefs() }
        pub fn denominator_coefs(&self) -> Vec<f64> { self.filter.denominator_coefs() }
        pub fn iir_type(&self) -> &IIRFilterType { self.filter.iir_type() }
    }
    impl<In: Sample> Filter<In, In> for DecimatingIIRFilter<f64, In> {
        /// iir/decim.rs:190
        fn execute(&mut self, sample: In) -> Vec<In> { self.filter.run(&[sample]) }
        /// iir/decim.rs:222
        fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.filter.run(samples) }
        fn frequency_response(&self, f: f64) -> Complex<f64> { self.filter.frequency_response(f) }
        fn group_delay(&self, f: f64) -> f64 { self.filter.group_delay(f) }
    }
}
pub mod interp {
    use super::*;
    /// InterpolatingIIRFilter -- iir/interp.rs:5: each input, then L - 1 zeros, all L outputs kept
    #[derive(Clone, Debug)]
    pub struct InterpolatingIIRFilter<Coef, In: Sample> where IIRFilter<Coef, In>: Clone { filter: IIRFilter<Coef, In>, interpolation: usize }
    impl<In: Sample> InterpolatingIIRFilter<f64, In> {
        /// iir/interp.rs:29
        pub fn new(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, interpolation: usize) -> Result<Self, Box<dyn Error>> {
            Ok(InterpolatingIIRFilter { filter: IIRFilter::create(ff, fb, iirtype, sys::SGPU_IIR_INTERPOLATING, interpolation)?, interpolation })
        }
        /// iir/interp.rs:62
        pub fn get_interpolation(&self) -> usize { self.interpolation }
        pub fn numerator_coefs(&self) -> Vec<f64> { self.filter.numerator_coefs() }
        pub fn denominator_coefs(&self) -> Vec<f64> { self.filter.denominator_coefs() }
        pub fn iir_type(&self) -> &IIRFilterType { self.filter.iir_type() }
    }
    impl<In: Sample> Filter<In, In> for InterpolatingIIRFilter<f64, In> {
        /// iir/interp.rs:184
        fn execute(&mut self, sample: In) -> Vec<In> { self.filter.run(&[sample]) }
        /// iir/interp.rs:215
        fn execute_block(&mut self, samples: &[In]) -> Vec<In> { self.filter.run(samples) }
        fn frequency_response(&self, f: f64) -> Complex<f64> { self.filter.frequency_response(f) }
        fn group_delay(&self, f: f64) -> f64 { self.filter.group_delay(f) }
    }
}
