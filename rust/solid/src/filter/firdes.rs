//! reference: src/filter/firdes/mod.rs -- `firdes_kaiser` (:278-305) with the arithmetic on the GPU
//! (sgpu_firdes_kaiser, csrc/firdes.cu: one thread per tap, f64, the reference's operations in the reference's order).
use solid_gpu_sys as sys;
use std::error::Error;
use std::fmt;
use std::ptr;

/// firdes/mod.rs:17-25 (the variants `firdes_kaiser` can return)
#[derive(Debug)]
pub enum FirdesErrorCode { Bandwidth, StopBandLevel, Mu }
#[derive(Debug)]
pub struct FirdesError(pub FirdesErrorCode);
impl fmt::Display for FirdesError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        let text = match self.0 {
            FirdesErrorCode::Bandwidth => "Invalid Bandwidth [0, 0.5]",
            FirdesErrorCode::StopBandLevel => "Invalid Stop Band Attenuation (0, inf)",
            FirdesErrorCode::Mu => "Invalid Mu Range [-0.5, 0.5]",
        };
        write!(f, "Firdes Error: {}", text)
    }
}
impl Error for FirdesError {}

fn status(st: i32) -> Result<(), Box<dyn Error>> {
    match st {
        sys::SGPU_OK => Ok(()),
        sys::SGPU_ERR_FIRDES_BANDWIDTH => Err(Box::new(FirdesError(FirdesErrorCode::Bandwidth))),
        sys::SGPU_ERR_FIRDES_STOP_BAND_LEVEL => Err(Box::new(FirdesError(FirdesErrorCode::StopBandLevel))),
        sys::SGPU_ERR_FIRDES_MU => Err(Box::new(FirdesError(FirdesErrorCode::Mu))),
        _ => panic!("sgpu_firdes_kaiser failed: {}", crate::last_error()),
    }
}

/// firdes_kaiser -- firdes/mod.rs:278-305
pub fn firdes_kaiser(filter_length: usize, cutoff_frequency: f64, stop_band_attenuation: f64,
                     fractional_sample_offset: f64) -> Result<Vec<f64>, Box<dyn Error>> {
    let mut h = vec![0.0f64; filter_length];
    let st = unsafe {
        sys::sgpu_firdes_kaiser(filter_length, &cutoff_frequency, &stop_band_attenuation, &fractional_sample_offset, 1,
                                h.as_mut_ptr(), sys::SGPU_HOST, ptr::null_mut())
    };
    status(st)?;
    Ok(h)
}

/// One design per channel in a single launch: row d of the result uses `cutoff_frequency[d]` (the bank of per-channel
/// filters `FIRFilter::new_per_channel` takes).
pub fn firdes_kaiser_bank(filter_length: usize, cutoff_frequency: &[f64], stop_band_attenuation: f64,
                          fractional_sample_offset: f64) -> Result<Vec<Vec<f64>>, Box<dyn Error>> {
    let n = cutoff_frequency.len();
    let (a, m) = (vec![stop_band_attenuation; n], vec![fractional_sample_offset; n]);
    let mut flat = vec![0.0f64; n * filter_length];
    let st = unsafe {
        sys::sgpu_firdes_kaiser(filter_length, cutoff_frequency.as_ptr(), a.as_ptr(), m.as_ptr(), n, flat.as_mut_ptr(),
                                sys::SGPU_HOST, ptr::null_mut())
    };
    status(st)?;
    Ok(flat.chunks(filter_length.max(1)).map(|r| r.to_vec()).collect())
}
