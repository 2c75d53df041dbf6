//! reference: src/filter/firdes/mod.rs -- `firdes_kaiser` (:278-305) with the arithmetic on the GPU
//! (sgpu_firdes_kaiser, csrc/firdes.cu: one thread per tap, f64, the reference's operations in the reference's order).
use solid_gpu_sys as sys;
use std::error::Error;
use std::fmt;
use std::ptr;

/// firdes/mod.rs:17-25 (the variants `firdes_kaiser` can return)
#[derive(Debug)]
pub enum FirdesErrorCode { Bandwidth, StopBandLevel, Mu, FilterSize, FFTSize }
#[derive(Debug)]
pub struct FirdesError(pub FirdesErrorCode);
impl fmt::Display for FirdesError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        let text = match self.0 {
            FirdesErrorCode::Bandwidth => "Invalid Bandwidth [0, 0.5]",
            FirdesErrorCode::StopBandLevel => "Invalid Stop Band Attenuation (0, inf)",
            FirdesErrorCode::Mu => "Invalid Mu Range [-0.5, 0.5]",
            FirdesErrorCode::FilterSize => "Invalid Filter Size [1, inf)",
            FirdesErrorCode::FFTSize => "Invalid FFT Size [1, inf)",
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

/// filter_autocorrelation -- firdes/mod.rs:443-456 (host f64)
pub fn filter_autocorrelation(filter: &[f64], lag: isize) -> f64 {
    let lag = lag.unsigned_abs();
    if lag >= filter.len() { return 0.0; }
    let mut rxx = 0.0;
    for i in lag..filter.len() { rxx += filter[i] * filter[i - lag]; }
    rxx
}

/// filter_isi -- firdes/mod.rs:553-573 (host f64): (rms, max)
pub fn filter_isi(filter: &[f64], samples_per_symbol: usize, filter_delay: usize) -> (f64, f64) {
    if 2 * samples_per_symbol * filter_delay + 1 != filter.len() { return (0.0, 0.0); }
    let rxx0 = filter_autocorrelation(filter, 0);
    let (mut isi_rms, mut isi_max) = (0.0f64, 0.0f64);
    for i in 1..(2 * filter_delay) {
        let e = (filter_autocorrelation(filter, (i * samples_per_symbol) as isize) / rxx0).abs();
        isi_rms += e * e;
        if i == 1 || e > isi_max { isi_max = e; }
    }
    ((isi_rms / (2.0 * filter_delay as f64)).sqrt(), isi_max)
}

/// filter_energy -- firdes/mod.rs:603-640: the crate's own caller of `DotProduct::execute` (:620-629).  The `fft_size`
/// sample vectors e^{j 2 pi f k} go to the GPU as ONE batched `sgpu_dot_execute` (f32 on the device).
pub fn filter_energy(filter: &[f64], cutoff_frequency: f64, fft_size: usize) -> Result<f64, Box<dyn Error>> {
    use num::complex::Complex;
    if !(0.0..=0.5).contains(&cutoff_frequency) { return Err(Box::new(FirdesError(FirdesErrorCode::Bandwidth))); }
    if filter.is_empty() { return Err(Box::new(FirdesError(FirdesErrorCode::FilterSize))); }
    if fft_size == 0 { return Err(Box::new(FirdesError(FirdesErrorCode::FFTSize))); }
    let n = filter.len();
    let mut ejwt = vec![Complex::<f32>::new(0.0, 0.0); fft_size * n];
    for i in 0..fft_size {
        let f = 0.5 * i as f64 / fft_size as f64;
        for k in 0..n {
            let c = Complex::<f64>::from_polar(1.0, 2.0 * std::f64::consts::PI * f * k as f64);
            ejwt[i * n + k] = Complex::new(c.re as f32, c.im as f32);
        }
    }
    let mut dp = ptr::null_mut();
    let st = unsafe { sys::sgpu_dot_create(filter.as_ptr(), n, sys::SGPU_TAPS_REAL, sys::SGPU_FORWARD, &mut dp) };
    crate::expect_ok(st, "sgpu_dot_create");
    let mut v = vec![Complex::<f32>::new(0.0, 0.0); fft_size];
    let st = unsafe {
        sys::sgpu_dot_execute(dp, ejwt.as_ptr() as *const f32, n, n, fft_size, v.as_mut_ptr() as *mut f32, sys::SGPU_HOST,
                              ptr::null_mut())
    };
    unsafe { sys::sgpu_dot_destroy(dp) };
    crate::expect_ok(st, "sgpu_dot_execute");
    let (mut e_total, mut e_stopband) = (0.0f64, 0.0f64);
    for (i, vi) in v.iter().enumerate() {
        let f = 0.5 * i as f64 / fft_size as f64;
        let e2 = (vi.re as f64) * (vi.re as f64) + (vi.im as f64) * (vi.im as f64);
        e_total += e2;
        if f > cutoff_frequency { e_stopband += e2; }
    }
    Ok(e_stopband / e_total)
}
