//! Several GPUs of the box behind one filter object (sgpu_ctx_* / sgpu_sharded_*, include/solid_gpu.h): host slices in,
//! `Vec` out, like every `Filter` of the reference -- channels split into contiguous ranges, one FIR / decimating-FIR
//! stream split into time segments with the T-1 sample halo sliced from the caller's buffer.  Not a reference type.
use crate::filter::fir::ctor_error;
use crate::filter::iir::IIRFilterType;
use crate::scalar::{Coefficient, Sample};
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::marker::PhantomData;
use std::ptr;

pub struct Context { h: *mut sys::sgpu_ctx }
impl Context {
    /// every visible GPU (n_gpus = 0) or the first n
    pub fn new(n_gpus: usize) -> Result<Self, Box<dyn Error>> {
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ctx_create(n_gpus as i32, &mut h) };
        if st != sys::SGPU_OK { return Err(crate::last_error().into()); }
        Ok(Context { h })
    }
    /// an explicit device list; a device may appear more than once
    pub fn with_devices(devices: &[i32]) -> Result<Self, Box<dyn Error>> {
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ctx_create_devices(devices.as_ptr(), devices.len() as i32, &mut h) };
        if st != sys::SGPU_OK { return Err(crate::last_error().into()); }
        Ok(Context { h })
    }
    pub fn devices(&self) -> usize { unsafe { sys::sgpu_ctx_devices(self.h) as usize } }
    /// FIRFilter::new (fir/mod.rs:79) or, with decimation >= 1, DecimatingFIRFilter::new (decim.rs:27), n_channels objects
    pub fn fir<Coef: Coefficient, In: Sample>(&self, coefficents: &[Coef], scale: Coef, n_channels: usize, decimation: usize)
        -> Result<Sharded<In>, Box<dyn Error>> {
        let flat = Coef::flatten(coefficents);
        let s = scale.to_complex();
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ctx_fir_create(self.h, flat.as_ptr(), coefficents.len(), Coef::KIND, n_channels, s.re, s.im,
                                                   (decimation > 0) as i32, decimation, &mut h) };
        if st != sys::SGPU_OK { return Err(ctor_error(st)); }
        Ok(Sharded { h, n_channels, _p: PhantomData })
    }
    /// InterpolatingFIRFilter::new (interp.rs:27)
    pub fn interpolator<Coef: Coefficient, In: Sample>(&self, coefficents: &[Coef], interpolation: usize, n_channels: usize)
        -> Result<Sharded<In>, Box<dyn Error>> {
        let flat = Coef::flatten(coefficents);
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ctx_interp_create(self.h, flat.as_ptr(), coefficents.len(), Coef::KIND, n_channels, interpolation, &mut h) };
        if st != sys::SGPU_OK { return Err(ctor_error(st)); }
        Ok(Sharded { h, n_channels, _p: PhantomData })
    }
    /// IIRFilter::new (iir/mod.rs:92)
    pub fn iir<In: Sample>(&self, ff: &[f64], fb: &[f64], iirtype: IIRFilterType, n_channels: usize) -> Result<Sharded<In>, Box<dyn Error>> {
        let t = if iirtype == IIRFilterType::Normal { sys::SGPU_IIR_NORMAL } else { sys::SGPU_IIR_SECOND_ORDER };
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ctx_iir_create(self.h, t, ff.as_ptr(), ff.len(), fb.as_ptr(), fb.len(), n_channels, sys::SGPU_IIR_PLAIN, 0, &mut h) };
        if st != sys::SGPU_OK { return Err(crate::last_error().into()); }
        Ok(Sharded { h, n_channels, _p: PhantomData })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::sgpu_ctx_destroy(self.h) }; } }

/// One filter object spread over the context's GPUs.  Samples are channel-major: `samples[c * n .. (c + 1) * n]`.
pub struct Sharded<In: Sample> { h: *mut sys::sgpu_sharded, n_channels: usize, _p: PhantomData<In> }
impl<In: Sample> Sharded<In> {
    pub fn shards(&self) -> usize { unsafe { sys::sgpu_sharded_shards(self.h) as usize } }
    pub fn last_segments(&self) -> usize { unsafe { sys::sgpu_sharded_last_segments(self.h) as usize } }
    pub fn reset(&mut self) { unsafe { sys::sgpu_sharded_reset(self.h) }; }
    /// Filter::execute_block (filter/mod.rs:14) for every channel; returns the outputs channel-major
    pub fn execute_block(&mut self, samples: &[In]) -> Vec<In> {
        assert_eq!(samples.len() % self.n_channels, 0);
        let n = samples.len() / self.n_channels;
        let x = In::narrow(samples);
        let n_out = unsafe { sys::sgpu_sharded_out_len(self.h, n) };
        let mut out = vec![Complex::new(0f32, 0f32); n_out * self.n_channels];
        let mut got = 0usize;
        let st = unsafe {
            sys::sgpu_sharded_execute_block(self.h, x.as_ptr() as *const f32, n, n.max(1), out.as_mut_ptr() as *mut f32, n_out.max(1), &mut got)
        };
        crate::expect_ok(st, "sgpu_sharded_execute_block");
        In::widen(out)
    }
}
impl<In: Sample> Drop for Sharded<In> { fn drop(&mut self) { unsafe { sys::sgpu_sharded_destroy(self.h) }; } }
