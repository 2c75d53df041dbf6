//! reference: src/group_delay/mod.rs:51-129 -- host-side f64 analysis, unchanged arithmetic.
use crate::scalar::Coefficient;
use num::complex::Complex;
use std::error::Error;
use std::fmt;

/// group_delay/mod.rs:24
pub const TOLERANCE: f64 = 0.00000000001;

/// group_delay/mod.rs:26-47
#[derive(Debug, PartialEq, Eq)]
pub enum DelayErrorCode { EmptyCoefficients, FrequencyOutOfBounds, DivideByZero }
#[derive(Debug)]
pub struct DelayError(pub DelayErrorCode);
impl fmt::Display for DelayError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        let msg = match self.0 {
            DelayErrorCode::EmptyCoefficients => "Empty Coefficients",
            DelayErrorCode::FrequencyOutOfBounds => "Frequency Out of Bounds [-0.5, 0.5]",
            DelayErrorCode::DivideByZero => "Denominator Coefficents Divide Numerator by Zero",
        };
        write!(f, "Delay Error: {}", msg)
    }
}
impl Error for DelayError {}

fn rot(frequency: f64, i: usize) -> Complex<f64> {
    Complex::from_polar(1.0, frequency * 2.0 * std::f64::consts::PI * (i as f64))
}

/// group_delay/mod.rs:51-79
pub fn fir_group_delay<C: Coefficient>(coefs: &[C], frequency: f64) -> Result<f64, Box<dyn Error>> {
    if coefs.is_empty() { return Err(Box::new(DelayError(DelayErrorCode::EmptyCoefficients))); }
    if !(-0.5..=0.5).contains(&frequency) { return Err(Box::new(DelayError(DelayErrorCode::FrequencyOutOfBounds))); }
    let mut t0 = Complex::new(0.0, 0.0);
    let mut t1 = Complex::new(0.0, 0.0);
    for (i, c) in coefs.iter().enumerate() {
        let v = c.to_complex() * rot(frequency, i);
        t0 += v * (i as f64);
        t1 += v;
    }
    Ok((t0 / t1).re)
}

/// group_delay/mod.rs:82-129
pub fn iir_group_delay<C: Coefficient>(num: &[C], den: &[C], frequency: f64) -> Result<f64, Box<dyn Error>> {
    if num.is_empty() || den.is_empty() { return Err(Box::new(DelayError(DelayErrorCode::EmptyCoefficients))); }
    if !(-0.5..=0.5).contains(&frequency) { return Err(Box::new(DelayError(DelayErrorCode::FrequencyOutOfBounds))); }
    let n = num.len() + den.len() - 1;
    let mut coefs = vec![Complex::new(0.0, 0.0); n];
    for i in 0..den.len() {
        for j in 0..num.len() {
            coefs[i + j] += den[den.len() - i - 1].to_complex().conj() * num[j].to_complex();
        }
    }
    let mut t0 = Complex::new(0.0, 0.0);
    let mut t1 = Complex::new(0.0, 0.0);
    for (i, c) in coefs.iter().enumerate() {
        let c0 = *c * rot(frequency, i);
        t0 += c0 * (i as f64);
        t1 += c0;
    }
    if t1.norm() <= TOLERANCE { return Err(Box::new(DelayError(DelayErrorCode::DivideByZero))); }
    Ok((t0 / t1).re - ((den.len() - 1) as f64))
}
