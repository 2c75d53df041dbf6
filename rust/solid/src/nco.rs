//! reference: src/nco/mod.rs -- numerically controlled oscillator.  The scalar methods are host arithmetic on the
//! handle's 32-bit words; `mix_up_block` / `mix_down_block` run on the GPU (the reference's versions index an
//! empty Vec and panic, nco/mod.rs:153-172; these are the loops they were written to be).
use crate::scalar::Sample;
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::error::Error;
use std::fmt;
use std::ptr;

/// nco/mod.rs:7-24
#[derive(Debug, PartialEq, Eq)]
pub enum NCOErrorCode { BandwidthOutOfRange }
#[derive(Debug)]
pub struct NCOError(pub NCOErrorCode);
impl fmt::Display for NCOError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "NCO Error Bandwidth out Range [0, inf)") }
}
impl Error for NCOError {}

/// nco/mod.rs:26-33
pub struct NCO {
    pub(crate) h: *mut sys::sgpu_nco,
    owned: bool,
    look_up_table: [f64; 1024],
    alpha: f64,
    beta: f64,
}

impl NCO {
    /// nco/mod.rs:36-50
    pub fn new() -> Self {
        let mut h = ptr::null_mut();
        crate::expect_ok(unsafe { sys::sgpu_nco_create(1, &mut h) }, "sgpu_nco_create");
        Self::wrap(h, true)
    }
    pub(crate) fn wrap(h: *mut sys::sgpu_nco, owned: bool) -> Self {
        let mut table = [0.0; 1024];
        for (i, item) in table.iter_mut().enumerate() {
            *item = (2.0 * std::f64::consts::PI * (i as f64) / 1024.0).sin();
        }
        let a = 0.1f64;
        NCO { h, owned, look_up_table: table, alpha: a, beta: a.sqrt() }
    }
    fn words(&self) -> (u32, u32) {
        let (mut t, mut d) = (0u32, 0u32);
        unsafe { sys::sgpu_nco_get(self.h, 0, &mut t, &mut d) };
        (t, d)
    }
    /// :53
    pub fn reset(&mut self) { unsafe { sys::sgpu_nco_reset(self.h) }; }
    /// :59
    pub fn set_frequency(&mut self, delta_theta: f64) { unsafe { sys::sgpu_nco_set_frequency(self.h, sys::SGPU_ALL_CHANNELS, delta_theta) }; }
    /// :64
    pub fn adjust_frequency(&mut self, dt: f64) { unsafe { sys::sgpu_nco_adjust_frequency(self.h, sys::SGPU_ALL_CHANNELS, dt) }; }
    /// :69-76 -- the integer division by 2^32 makes this 0 for every value (kept)
    pub fn get_frequency(&self) -> f64 {
        let dt = (self.words().1 as u64 / (1u64 << 32)) as f64 * 2.0f64 * std::f64::consts::PI;
        if dt > std::f64::consts::PI { dt - 2.0f64 * std::f64::consts::PI } else { dt }
    }
    /// :79
    pub fn set_phase(&mut self, phi: f64) { unsafe { sys::sgpu_nco_set_phase(self.h, sys::SGPU_ALL_CHANNELS, phi) }; }
    /// :84
    pub fn adjust_phase(&mut self, delta_phi: f64) { unsafe { sys::sgpu_nco_adjust_phase(self.h, sys::SGPU_ALL_CHANNELS, delta_phi) }; }
    /// :89-91
    pub fn get_phase(&self) -> f64 { (self.words().0 as u64 / (1u64 << 32)) as f64 * 2.0f64 * std::f64::consts::PI }
    /// :93
    pub fn step(&mut self) { unsafe { sys::sgpu_nco_step(self.h, 1) }; }
    fn index(&self) -> usize { ((self.words().0.wrapping_add(1 << 21) >> 22) & 0x3ff) as usize }
    /// :103
    pub fn sin(&self) -> f64 { self.look_up_table[self.index()] }
    /// :108
    pub fn cos(&self) -> f64 { self.look_up_table[(self.index() + 256) & 0x3ff] }
    /// :114
    pub fn sincos(&self) -> (f64, f64) { (self.sin(), self.cos()) }
    /// :119
    pub fn complex_exponential(&self) -> Complex<f64> { Complex::new(self.cos(), self.sin()) }
    /// :123-131
    pub fn set_internal_pll_bandwidth(&mut self, bandwidth: f64) -> Result<(), Box<dyn Error>> {
        if bandwidth < 0.0 { return Err(Box::new(NCOError(NCOErrorCode::BandwidthOutOfRange))); }
        self.alpha = bandwidth;
        self.beta = self.alpha.sqrt();
        Ok(())
    }
    /// :134-137
    pub fn pll_step(&mut self, delta_phi: f64) {
        self.adjust_frequency(delta_phi * self.alpha);
        self.adjust_phase(delta_phi * self.beta);
    }
    /// :141
    pub fn mix_up(&self, input: Complex<f64>) -> Complex<f64> { self.complex_exponential() * input }
    /// :147
    pub fn mix_down(&self, input: Complex<f64>) -> Complex<f64> { self.complex_exponential().conj() * input }
    fn mix_block<I: Sample>(&mut self, up: bool, input: &[I]) -> Vec<I> {
        let x = I::narrow(input);
        let mut out = vec![Complex::new(0f32, 0f32); x.len()];
        let st = unsafe {
            sys::sgpu_nco_mix_block(self.h, up as i32, x.as_ptr() as *const f32, x.len(), x.len().max(1),
                                    out.as_mut_ptr() as *mut f32, x.len().max(1), sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_nco_mix_block");
        I::widen(out)
    }
    /// :153 -- y[i] = mix_up(x[i]); step()
    pub fn mix_up_block<I: Sample>(&mut self, input: &[I]) -> Vec<I> { self.mix_block(true, input) }
    /// :164
    pub fn mix_down_block<I: Sample>(&mut self, input: &[I]) -> Vec<I> { self.mix_block(false, input) }
}
/// nco/mod.rs:176-188
pub fn constrain(theta: f64) -> u32 { unsafe { sys::sgpu_nco_constrain(theta) } }

impl Default for NCO { fn default() -> Self { Self::new() } }
impl Drop for NCO { fn drop(&mut self) { if self.owned { unsafe { sys::sgpu_nco_destroy(self.h) }; } } }
impl fmt::Debug for NCO {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { fmt::Display::fmt(self, f) }
}
impl fmt::Display for NCO {
    /// nco/mod.rs:196-203
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        let (t, d) = self.words();
        write!(f, "NCO [Theta={}] [ΔTheta={}] [Alpha={}] [Beta={}]", t, d, self.alpha, self.beta)
    }
}
