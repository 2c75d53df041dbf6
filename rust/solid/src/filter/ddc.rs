//! Digital down-converter: `NCO::mix_down` + `step` per sample (nco/mod.rs:93-96,147-151) in front of a
//! `DecimatingFIRFilter` (fir/decim.rs) -- the loop a user of the reference writes as
//! `for x in samples { out.extend(decim.execute(nco.mix_down(x))); nco.step(); }` -- as ONE kernel per call on the
//! shapes the decimator's warp kernel serves (the mixed stream never reaches HBM).  Not a reference type.
use super::fir::ctor_error;
use crate::nco::NCO;
use crate::scalar::{Coefficient, Sample};
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::marker::PhantomData;
use std::ptr;

pub struct DigitalDownConverter<Coef: Coefficient, In: Sample> { h: *mut sys::sgpu_ddc, nco: NCO, _p: PhantomData<(Coef, In)> }

impl<Coef: Coefficient, In: Sample> DigitalDownConverter<Coef, In> {
    /// the decimator's arguments (fir/decim.rs:27) + the oscillator's frequency (nco/mod.rs:59)
    pub fn new(coefficents: &[Coef], scale: Coef, decimation: usize, frequency: f64) -> Result<Self, Box<dyn Error>> {
        let flat = Coef::flatten(coefficents);
        let s = scale.to_complex();
        let mut h = ptr::null_mut();
        let st = unsafe { sys::sgpu_ddc_create(flat.as_ptr(), coefficents.len(), Coef::KIND, 1, s.re, s.im, decimation, &mut h) };
        if st != sys::SGPU_OK { return Err(ctor_error(st)); }
        let mut nco = NCO::wrap(unsafe { sys::sgpu_ddc_nco(h) }, false);
        nco.set_frequency(frequency);
        Ok(DigitalDownConverter { h, nco, _p: PhantomData })
    }
    /// the oscillator: set_frequency / adjust_phase / pll_step ... between calls
    pub fn nco(&mut self) -> &mut NCO { &mut self.nco }
    pub fn get_decimation(&self) -> usize { unsafe { sys::sgpu_fir_decimation(sys::sgpu_ddc_filter(self.h)) } }
    /// true when the last call mixed inside the decimator kernel
    pub fn last_fused(&self) -> bool { unsafe { sys::sgpu_ddc_last_fused(self.h) == 1 } }
    pub fn reset(&mut self) { unsafe { sys::sgpu_ddc_reset(self.h) }; }
    pub fn execute_block(&mut self, samples: &[In]) -> Vec<In> {
        let x = In::narrow(samples);
        let n_out = unsafe { sys::sgpu_ddc_out_len(self.h, x.len()) };
        let mut out = vec![Complex::new(0f32, 0f32); n_out];
        let mut got = 0usize;
        let st = unsafe {
            sys::sgpu_ddc_execute_block(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), out.as_mut_ptr() as *mut f32,
                                        n_out.max(1), &mut got, sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_ddc_execute_block");
        In::widen(out)
    }
    /// mixed and pushed, no output (fir/decim.rs:136)
    pub fn write(&mut self, samples: &[In]) {
        let x = In::narrow(samples);
        let st = unsafe { sys::sgpu_ddc_write(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), sys::SGPU_HOST, ptr::null_mut()) };
        crate::expect_ok(st, "sgpu_ddc_write");
    }
}
impl<Coef: Coefficient, In: Sample> Drop for DigitalDownConverter<Coef, In> {
    fn drop(&mut self) { unsafe { sys::sgpu_ddc_destroy(self.h) }; }
}
