//! reference: src/filter/auto_correlator/mod.rs
use crate::scalar::Sample;
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::marker::PhantomData;
use std::{fmt, ptr};

/// AutoCorrelator<C> -- auto_correlator/mod.rs:24-35
pub struct AutoCorrelator<C: Sample> { h: *mut sys::sgpu_autocorr, _p: PhantomData<C> }

impl<C: Sample> AutoCorrelator<C> {
    /// auto_correlator/mod.rs:51
    pub fn new(window_size: usize, delay: usize) -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_autocorr_create(window_size, delay, 1, &mut h) }, "sgpu_autocorr_create");
        AutoCorrelator { h, _p: PhantomData }
    }
    /// :76
    pub fn reset(&mut self) { unsafe { sys::sgpu_autocorr_reset(self.h) }; }
    /// :99
    pub fn push(&mut self, sample: C) { let _ = self.write(&[sample]); }
    /// :130
    pub fn write(&mut self, samples: &[C]) -> Result<(), Box<dyn Error>> {
        let x = C::narrow(samples);
        let st = unsafe { sys::sgpu_autocorr_write(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), sys::SGPU_HOST, ptr::null_mut()) };
        crate::expect_ok(st, "sgpu_autocorr_write");
        Ok(())
    }
    /// :165
    pub fn execute(&self) -> C {
        let mut out = [0.0f64; 2];
        unsafe { sys::sgpu_autocorr_execute(self.h, out.as_mut_ptr()) };
        C::from_cf32(Complex::new(out[0] as f32, out[1] as f32))
    }
    /// :184 -- one output per input
    pub fn execute_block(&mut self, samples: &[C]) -> Vec<C> {
        let x = C::narrow(samples);
        let mut out = vec![Complex::new(0.0f32, 0.0f32); x.len()];
        let mut n_out = 0usize;
        let st = unsafe {
            sys::sgpu_autocorr_execute_block(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), out.as_mut_ptr() as *mut f32,
                                             out.len().max(1), &mut n_out, sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_autocorr_execute_block");
        C::widen(out)
    }
    /// :214
    pub fn get_energy(&self) -> f64 {
        let mut e = 0.0f64;
        unsafe { sys::sgpu_autocorr_get_energy(self.h, &mut e) };
        e
    }
}
impl<C: Sample> Clone for AutoCorrelator<C> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_autocorr_clone(self.h, &mut h) }, "sgpu_autocorr_clone");
        AutoCorrelator { h, _p: PhantomData }
    }
}
impl<C: Sample> Drop for AutoCorrelator<C> { fn drop(&mut self) { unsafe { sys::sgpu_autocorr_destroy(self.h) }; } }
impl<C: Sample> fmt::Debug for AutoCorrelator<C> {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { fmt::Display::fmt(self, f) }
}
impl<C: Sample> fmt::Display for AutoCorrelator<C> {
    /// auto_correlator/mod.rs:219-228
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        let (w, d) = unsafe { (sys::sgpu_autocorr_window_size(self.h), sys::sgpu_autocorr_delay(self.h)) };
        write!(f, "AutoCorrelator<{}> [Size={}] [Delay={}] [Energy={}]", std::any::type_name::<C>(), w, d, self.get_energy())
    }
}
