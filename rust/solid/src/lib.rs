//! Drop-in for the reference crate's filtering path: same module paths, type and method names,
//! constructor arguments and error variants; `execute` / `execute_block` run on the GPU.
//! Sample type is `Complex<f32>` (the f32 instantiation of the north star); coefficients are `f64`
//! like the reference's `Coef`.  NOT COMPILED in this repository's environment (no Rust toolchain).

pub mod dot_product {
    //! reference: src/dot_product/mod.rs, src/dot_product/execute.rs
    use num::complex::Complex;
    use solid_gpu_sys as sys;
    use std::ptr;

    pub enum Direction { FORWARD, REVERSE }

    pub struct DotProduct { h: *mut sys::sgpu_dot }

    pub mod execute {
        pub trait Execute<I, O> { fn execute(&self, samples: &[I]) -> O; }
    }

    impl DotProduct {
        /// DotProduct::new -- dot_product/mod.rs:57
        pub fn new(coefficients: &[f64], direction: Direction) -> Self {
            let mut h = ptr::null_mut();
            let dir = match direction { Direction::FORWARD => sys::SGPU_FORWARD, Direction::REVERSE => sys::SGPU_REVERSE };
            let st = unsafe { sys::sgpu_dot_create(coefficients.as_ptr(), coefficients.len(), sys::SGPU_TAPS_REAL, dir, &mut h) };
            assert_eq!(st, sys::SGPU_OK, "sgpu_dot_create failed: {}", crate::last_error());
            DotProduct { h }
        }
        /// [sic] dot_product/mod.rs:102 -- stored order
        pub fn coefficents(&self) -> Vec<f64> {
            let mut v = vec![0.0; self.len()];
            if !v.is_empty() { unsafe { sys::sgpu_dot_coefficients(self.h, v.as_mut_ptr()) }; }
            v
        }
        pub fn len(&self) -> usize { unsafe { sys::sgpu_dot_len(self.h) } }
        pub fn is_empty(&self) -> bool { self.len() == 0 }
    }
    impl execute::Execute<Complex<f32>, Complex<f32>> for DotProduct {
        /// Execute::execute -- dot_product/mod.rs:153-171
        fn execute(&self, samples: &[Complex<f32>]) -> Complex<f32> {
            let mut r = Complex::new(0f32, 0f32);
            let st = unsafe {
                sys::sgpu_dot_execute(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(), 1,
                                      &mut r as *mut Complex<f32> as *mut f32, sys::SGPU_HOST, ptr::null_mut())
            };
            assert_eq!(st, sys::SGPU_OK, "sgpu_dot_execute failed: {}", crate::last_error());
            r
        }
    }
    impl Drop for DotProduct { fn drop(&mut self) { unsafe { sys::sgpu_dot_destroy(self.h) }; } }
}

pub(crate) fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(solid_gpu_sys::sgpu_last_error()).to_string_lossy().into_owned() }
}

pub mod filter {
    //! reference: src/filter/mod.rs:9-22
    use num::complex::Complex;

    pub trait Filter<I, O> {
        fn execute(&mut self, sample: I) -> Vec<O>;
        fn execute_block(&mut self, samples: &[I]) -> Vec<O>;
        fn frequency_response(&self, frequency: f64) -> Complex<f64>;
        fn group_delay(&self, frequency: f64) -> f64;
    }

    pub mod fir {
        //! reference: src/filter/fir/mod.rs, decim.rs, interp.rs, pfb.rs
        use super::Filter;
        use num::complex::Complex;
        use solid_gpu_sys as sys;
        use std::error::Error;
        use std::{fmt, ptr};

        #[derive(Debug)]
        pub enum FIRErrorCode { CoefficientsLengthZero, DecimationLessThanOne, InterpolationLessThanOne, NotEnoughFilters }
        #[derive(Debug)]
        pub struct FIRError(pub FIRErrorCode);
        impl fmt::Display for FIRError {
            fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "FIR Filter Error {:?}", self.0) }
        }
        impl Error for FIRError {}

        fn ctor_error(st: i32) -> Box<dyn Error> {
            match st {
                sys::SGPU_ERR_FIR_COEFFICIENTS_LENGTH_ZERO => Box::new(FIRError(FIRErrorCode::CoefficientsLengthZero)),
                sys::SGPU_ERR_FIR_DECIMATION_LESS_THAN_ONE => Box::new(FIRError(FIRErrorCode::DecimationLessThanOne)),
                sys::SGPU_ERR_FIR_INTERPOLATION_LESS_THAN_ONE => Box::new(FIRError(FIRErrorCode::InterpolationLessThanOne)),
                sys::SGPU_ERR_FIR_NOT_ENOUGH_FILTERS => Box::new(FIRError(FIRErrorCode::NotEnoughFilters)),
                _ => crate::last_error().into(),
            }
        }
        fn response(coefs: &[f64], frequency: f64) -> Complex<f64> {
            coefs.iter().enumerate().fold(Complex::new(0.0, 0.0), |acc, (i, c)| {
                acc + *c * Complex::from_polar(1.0, frequency * 2.0 * std::f64::consts::PI * (i as f64))
            })
        }

        /// FIRFilter<Coef = f64, In = Complex<f32>> -- fir/mod.rs:58
        pub struct FIRFilter { pub(crate) h: *mut sys::sgpu_fir }

        impl FIRFilter {
            /// fir/mod.rs:79
            pub fn new(coefficents: &[f64], scale: f64) -> Result<Self, Box<dyn Error>> {
                let mut h = ptr::null_mut();
                let st = unsafe { sys::sgpu_fir_create(coefficents.as_ptr(), coefficents.len(), sys::SGPU_TAPS_REAL, 1, scale, 0.0, 0, 0, &mut h) };
                if st != sys::SGPU_OK { return Err(ctor_error(st)); }
                Ok(FIRFilter { h })
            }
            pub fn set_scale(&mut self, scale: f64) { unsafe { sys::sgpu_fir_set_scale(self.h, scale, 0.0) }; }
            pub fn get_scale(&self) -> f64 { let (mut re, mut im) = (0.0, 0.0); unsafe { sys::sgpu_fir_get_scale(self.h, &mut re, &mut im) }; re }
            pub fn len(&self) -> usize { unsafe { sys::sgpu_fir_len(self.h) } }
            pub fn is_empty(&self) -> bool { self.len() == 0 }
            /// true when the last execute_block ran on the tcgen05 tensor-core kernel (not part of the reference API)
            pub fn last_path_tensor(&self) -> bool { unsafe { sys::sgpu_fir_last_path(self.h) == 1 } }
            /// stored (reversed) order -- fir/mod.rs:176
            pub fn coefficients(&self) -> Vec<f64> { let mut v = vec![0.0; self.len()]; unsafe { sys::sgpu_fir_coefficients(self.h, v.as_mut_ptr()) }; v }
            pub(crate) fn run(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> {
                let n_out = unsafe { sys::sgpu_fir_out_len(self.h, samples.len()) };
                let mut out = vec![Complex::new(0f32, 0f32); n_out];
                let mut got = 0usize;
                let st = unsafe {
                    sys::sgpu_fir_execute_block(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(),
                                                out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
                };
                // execute never fails in the reference; there is no CPU fallback to take instead
                assert_eq!(st, sys::SGPU_OK, "sgpu_fir_execute_block failed: {}", crate::last_error());
                out
            }
        }
        impl Clone for FIRFilter {
            fn clone(&self) -> Self { let mut h = ptr::null_mut(); let st = unsafe { sys::sgpu_fir_clone(self.h, &mut h) }; assert_eq!(st, sys::SGPU_OK); FIRFilter { h } }
        }
        impl Drop for FIRFilter { fn drop(&mut self) { unsafe { sys::sgpu_fir_destroy(self.h) }; } }
        impl Filter<Complex<f32>, Complex<f32>> for FIRFilter {
            fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.run(&[sample]) }           // :209
            fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.run(samples) }   // :235
            fn frequency_response(&self, frequency: f64) -> Complex<f64> { response(&self.coefficients(), frequency) * self.get_scale() }
            fn group_delay(&self, _frequency: f64) -> f64 { unimplemented!("host-side analysis: port group_delay::fir_group_delay unchanged") }
        }

        pub mod decim {
            use super::*;
            /// DecimatingFIRFilter -- fir/decim.rs:5
            #[derive(Clone)]
            pub struct DecimatingFIRFilter { filter: FIRFilter }
            impl DecimatingFIRFilter {
                /// fir/decim.rs:27
                pub fn new(coefficents: &[f64], scale: f64, decimation: usize) -> Result<Self, Box<dyn Error>> {
                    let mut h = ptr::null_mut();
                    let st = unsafe { sys::sgpu_fir_create(coefficents.as_ptr(), coefficents.len(), sys::SGPU_TAPS_REAL, 1, scale, 0.0, 1, decimation, &mut h) };
                    if st != sys::SGPU_OK { return Err(ctor_error(st)); }
                    Ok(DecimatingFIRFilter { filter: FIRFilter { h } })
                }
                pub fn set_scale(&mut self, scale: f64) { self.filter.set_scale(scale) }
                pub fn get_scale(&self) -> f64 { self.filter.get_scale() }
                pub fn get_decimation(&self) -> usize { unsafe { sys::sgpu_fir_decimation(self.filter.h) } }
                pub fn push(&mut self, sample: Complex<f32>) { self.write(&[sample]) }                      // :115
                pub fn write(&mut self, samples: &[Complex<f32>]) {                                          // :136
                    unsafe { sys::sgpu_fir_write(self.filter.h, samples.as_ptr() as *const f32, samples.len(), samples.len(), sys::SGPU_HOST, ptr::null_mut()) };
                }
                pub fn len(&self) -> usize { self.filter.len() }
                pub fn is_empty(&self) -> bool { self.filter.is_empty() }
                pub fn coefficients(&self) -> Vec<f64> { self.filter.coefficients() }
            }
            impl Filter<Complex<f32>, Complex<f32>> for DecimatingFIRFilter {
                fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.filter.run(&[sample]) }         // :221
                fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.filter.run(samples) } // :250
                fn frequency_response(&self, frequency: f64) -> Complex<f64> { self.filter.frequency_response(frequency) }
                fn group_delay(&self, frequency: f64) -> f64 { self.filter.group_delay(frequency) }
            }
        }

        pub mod interp {
            use super::*;
            /// InterpolatingFIRFilter -- fir/interp.rs:6
            pub struct InterpolatingFIRFilter { h: *mut sys::sgpu_interp }
            impl InterpolatingFIRFilter {
                /// fir/interp.rs:27
                pub fn new(coefficents: &[f64], interpolation: usize) -> Result<Self, Box<dyn Error>> {
                    let mut h = ptr::null_mut();
                    let st = unsafe { sys::sgpu_interp_create(coefficents.as_ptr(), coefficents.len(), sys::SGPU_TAPS_REAL, 1, interpolation, &mut h) };
                    if st != sys::SGPU_OK { return Err(ctor_error(st)); }
                    Ok(InterpolatingFIRFilter { h })
                }
                pub fn set_scale(&mut self, scale: f64) { unsafe { sys::sgpu_interp_set_scale(self.h, scale, 0.0) }; }
                pub fn get_scale(&self) -> f64 { let (mut re, mut im) = (0.0, 0.0); unsafe { sys::sgpu_interp_get_scale(self.h, &mut re, &mut im) }; re }
                pub fn len(&self) -> usize { unsafe { sys::sgpu_interp_interpolation(self.h) } }
                pub fn is_empty(&self) -> bool { self.len() == 0 }
                pub fn interpolation(&self) -> usize { self.len() }
                pub fn last_path_tensor(&self) -> bool { unsafe { sys::sgpu_interp_last_path(self.h) == 1 } }
                pub fn coefficents(&self) -> Vec<f64> {
                    let mut v = vec![0.0; self.len() * unsafe { sys::sgpu_interp_sub_len(self.h) }];
                    unsafe { sys::sgpu_interp_coefficients(self.h, v.as_mut_ptr()) }; v
                }
                fn run(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> {
                    let n_out = samples.len() * self.len();
                    let mut out = vec![Complex::new(0f32, 0f32); n_out];
                    let mut got = 0usize;
                    let st = unsafe {
                        sys::sgpu_interp_execute_block(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(),
                                                       out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
                    };
                    assert_eq!(st, sys::SGPU_OK, "sgpu_interp_execute_block failed: {}", crate::last_error());
                    out
                }
            }
            impl Clone for InterpolatingFIRFilter {
                fn clone(&self) -> Self { let mut h = ptr::null_mut(); let st = unsafe { sys::sgpu_interp_clone(self.h, &mut h) }; assert_eq!(st, sys::SGPU_OK); InterpolatingFIRFilter { h } }
            }
            impl Drop for InterpolatingFIRFilter { fn drop(&mut self) { unsafe { sys::sgpu_interp_destroy(self.h) }; } }
            impl Filter<Complex<f32>, Complex<f32>> for InterpolatingFIRFilter {
                fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.run(&[sample]) }          // :93
                fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.run(samples) }  // :102
                fn frequency_response(&self, frequency: f64) -> Complex<f64> { response(&self.coefficents(), frequency) * self.get_scale() }
                fn group_delay(&self, _frequency: f64) -> f64 { unimplemented!("host-side analysis") }
            }
        }
    }

    pub mod iir {
        //! reference: src/filter/iir/mod.rs, decim.rs, interp.rs
        use super::Filter;
        use num::complex::Complex;
        use solid_gpu_sys as sys;
        use std::error::Error;
        use std::{fmt, ptr};

        #[derive(Debug)]
        pub enum IIRErrorCode {
            NumeratorLengthZero, DenominatorLengthZero, SecondOrderSectionSizeZero, SecondOrderSectionSizeMismatch,
            SecondOrderSectionSizeNotMultpleOf3, DecimationLessThanOne, InterpolationLessThanOne,
        }
        #[derive(Debug)]
        pub struct IIRError(pub IIRErrorCode);
        impl fmt::Display for IIRError {
            fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "IIR Filter Error {:?}", self.0) }
        }
        impl Error for IIRError {}
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
                _ => return crate::last_error().into(),
            };
            Box::new(IIRError(code))
        }

        /// IIRFilter<Coef = f64, In = Complex<f32>> -- iir/mod.rs:68
        pub struct IIRFilter { h: *mut sys::sgpu_iir, iirtype: IIRFilterType }
        impl IIRFilter {
            fn create(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, wrap: i32, factor: usize) -> Result<Self, Box<dyn Error>> {
                let mut h = ptr::null_mut();
                let t = if iirtype == IIRFilterType::Normal { sys::SGPU_IIR_NORMAL } else { sys::SGPU_IIR_SECOND_ORDER };
                let st = unsafe { sys::sgpu_iir_create(t, ff.as_ptr(), ff.len(), fb.as_ptr(), fb.len(), 1, wrap, factor, &mut h) };
                if st != sys::SGPU_OK { return Err(ctor_error(st)); }
                Ok(IIRFilter { h, iirtype })
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
            pub fn numerator_coefs(&self) -> Vec<f64> { self.coefs(true) }     // :182
            pub fn denominator_coefs(&self) -> Vec<f64> { self.coefs(false) }  // :202
            pub fn iir_type(&self) -> &IIRFilterType { &self.iirtype }         // :239
            fn run(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> {
                let n_out = unsafe { sys::sgpu_iir_out_len(self.h, samples.len()) };
                let mut out = vec![Complex::new(0f32, 0f32); n_out];
                let mut got = 0usize;
                let st = unsafe {
                    sys::sgpu_iir_execute_block(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(),
                                                out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
                };
                assert_eq!(st, sys::SGPU_OK, "sgpu_iir_execute_block failed: {}", crate::last_error());
                out
            }
        }
        impl Clone for IIRFilter {
            fn clone(&self) -> Self { let mut h = ptr::null_mut(); let st = unsafe { sys::sgpu_iir_clone(self.h, &mut h) }; assert_eq!(st, sys::SGPU_OK); IIRFilter { h, iirtype: self.iirtype } }
        }
        impl Drop for IIRFilter { fn drop(&mut self) { unsafe { sys::sgpu_iir_destroy(self.h) }; } }
        impl Filter<Complex<f32>, Complex<f32>> for IIRFilter {
            fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.run(&[sample]) }          // :270
            fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.run(samples) }  // :310
            fn frequency_response(&self, _frequency: f64) -> Complex<f64> { unimplemented!("host-side analysis: port iir/mod.rs:336-373 unchanged") }
            fn group_delay(&self, _frequency: f64) -> f64 { unimplemented!("host-side analysis") }
        }

        pub mod decim {
            use super::*;
            /// DecimatingIIRFilter -- iir/decim.rs:5
            #[derive(Clone)]
            pub struct DecimatingIIRFilter { filter: IIRFilter, decimation: usize }
            impl DecimatingIIRFilter {
                pub fn new(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, decimation: usize) -> Result<Self, Box<dyn Error>> {
                    Ok(DecimatingIIRFilter { filter: IIRFilter::create(ff, fb, iirtype, sys::SGPU_IIR_DECIMATING, decimation)?, decimation })
                }
                pub fn get_decimation(&self) -> usize { self.decimation }
            }
            impl Filter<Complex<f32>, Complex<f32>> for DecimatingIIRFilter {
                fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.filter.run(&[sample]) }
                fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.filter.run(samples) }
                fn frequency_response(&self, f: f64) -> Complex<f64> { self.filter.frequency_response(f) }
                fn group_delay(&self, f: f64) -> f64 { self.filter.group_delay(f) }
            }
        }
        pub mod interp {
            use super::*;
            /// InterpolatingIIRFilter -- iir/interp.rs:5
            #[derive(Clone)]
            pub struct InterpolatingIIRFilter { filter: IIRFilter, interpolation: usize }
            impl InterpolatingIIRFilter {
                pub fn new(ff: &[f64], fb: &[f64], iirtype: IIRFilterType, interpolation: usize) -> Result<Self, Box<dyn Error>> {
                    Ok(InterpolatingIIRFilter { filter: IIRFilter::create(ff, fb, iirtype, sys::SGPU_IIR_INTERPOLATING, interpolation)?, interpolation })
                }
                pub fn get_interpolation(&self) -> usize { self.interpolation }
            }
            impl Filter<Complex<f32>, Complex<f32>> for InterpolatingIIRFilter {
                fn execute(&mut self, sample: Complex<f32>) -> Vec<Complex<f32>> { self.filter.run(&[sample]) }
                fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> { self.filter.run(samples) }
                fn frequency_response(&self, f: f64) -> Complex<f64> { self.filter.frequency_response(f) }
                fn group_delay(&self, f: f64) -> f64 { self.filter.group_delay(f) }
            }
        }
    }
    pub mod auto_correlator {
        //! reference: src/filter/auto_correlator/mod.rs
        use num::complex::Complex;
        use solid_gpu_sys as sys;
        use std::error::Error;
        use std::fmt;
        use std::ptr;

        /// AutoCorrelator<C> -- auto_correlator/mod.rs:24-35 (the f32-sample instantiation)
        pub struct AutoCorrelator { h: *mut sys::sgpu_autocorr }

        impl AutoCorrelator {
            /// AutoCorrelator::new -- auto_correlator/mod.rs:51
            pub fn new(window_size: usize, delay: usize) -> Self {
                let mut h = ptr::null_mut();
                let st = unsafe { sys::sgpu_autocorr_create(window_size, delay, 1, &mut h) };
                assert_eq!(st, sys::SGPU_OK, "sgpu_autocorr_create failed: {}", crate::last_error());
                AutoCorrelator { h }
            }
            /// :76
            pub fn reset(&mut self) { unsafe { sys::sgpu_autocorr_reset(self.h) }; }
            /// :99
            pub fn push(&mut self, sample: Complex<f32>) { let _ = self.write(&[sample]); }
            /// :130
            pub fn write(&mut self, samples: &[Complex<f32>]) -> Result<(), Box<dyn Error>> {
                let st = unsafe { sys::sgpu_autocorr_write(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(),
                                                           sys::SGPU_HOST, ptr::null_mut()) };
                assert_eq!(st, sys::SGPU_OK, "sgpu_autocorr_write failed: {}", crate::last_error());
                Ok(())
            }
            /// :165
            pub fn execute(&self) -> Complex<f32> {
                let mut out = [0.0f64; 2];
                unsafe { sys::sgpu_autocorr_execute(self.h, out.as_mut_ptr()) };
                Complex::new(out[0] as f32, out[1] as f32)
            }
            /// :184 -- one output per input
            pub fn execute_block(&mut self, samples: &[Complex<f32>]) -> Vec<Complex<f32>> {
                let mut out = vec![Complex::new(0.0f32, 0.0f32); samples.len()];
                let mut n_out = 0usize;
                let st = unsafe { sys::sgpu_autocorr_execute_block(self.h, samples.as_ptr() as *const f32, samples.len(), samples.len(),
                                                                   out.as_mut_ptr() as *mut f32, out.len().max(1), &mut n_out,
                                                                   sys::SGPU_HOST, ptr::null_mut()) };
                assert_eq!(st, sys::SGPU_OK, "sgpu_autocorr_execute_block failed: {}", crate::last_error());
                out
            }
            /// :214
            pub fn get_energy(&self) -> f64 {
                let mut e = 0.0f64;
                unsafe { sys::sgpu_autocorr_get_energy(self.h, &mut e) };
                e
            }
        }
        impl Clone for AutoCorrelator {
            fn clone(&self) -> Self {
                let mut h = ptr::null_mut();
                let st = unsafe { sys::sgpu_autocorr_clone(self.h, &mut h) };
                assert_eq!(st, sys::SGPU_OK);
                AutoCorrelator { h }
            }
        }
        impl Drop for AutoCorrelator {
            fn drop(&mut self) { unsafe { sys::sgpu_autocorr_destroy(self.h) }; }
        }
        impl fmt::Display for AutoCorrelator {
            /// auto_correlator/mod.rs:219-228
            fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
                let (w, d) = unsafe { (sys::sgpu_autocorr_window_size(self.h), sys::sgpu_autocorr_delay(self.h)) };
                write!(f, "AutoCorrelator<f32> [Size={}] [Delay={}] [Energy={}]", w, d, self.get_energy())
            }
        }
    }
}
