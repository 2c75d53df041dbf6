//! Drop-in for the reference crate's filtering path: the same module paths, type names, generic parameters,
//! method names, constructor arguments and error variants as `juliantos/solid-dsp`; `execute` /
//! `execute_block` run on a B200 through `solid-gpu-sys` (include/solid_gpu.h).
//!
//! Generic surface.  The reference is generic over `Coef` and `In` (`FIRFilter<Coef, In>`,
//! filter/fir/mod.rs:58-63; its only caller writes `IIRFilter::<f64, Complex<f64>>::new`, main.rs:39).
//! Here `Coef: scalar::Coefficient` (`f64` or `Complex<f64>`, crossing the C ABI as doubles) and
//! `In: scalar::Sample` (`Complex<f32>`, the device's sample type, or `Complex<f64>`, narrowed to f32 on the
//! host before the call and widened afterwards -- the arithmetic is f32 on the GPU either way).
//!
//! NOT COMPILED in this repository's environment (no Rust toolchain in the image or on the GPU boxes); the
//! compiled host mirrors are include/solid.hpp (C++) and solid_dsp_b200/ (Python).  tests/test_abi_symbols.py
//! checks that every `pub fn` of SURVEY.md Appendix C exists in this source and that the -sys crate declares
//! every prototype of the header.

pub mod scalar;
pub mod dot_product;
pub mod window;
pub mod circular_buffer;
pub mod group_delay;
pub mod nco;
pub mod filter;
pub mod multi_gpu;

pub(crate) fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(solid_gpu_sys::sgpu_last_error()).to_string_lossy().into_owned() }
}

/// Execute never fails in the reference (it returns a `Vec`); a CUDA failure here has no CPU path to fall
/// back on, so it panics with the library's message -- like the reference's allocation failures
/// (dot_product/mod.rs:62, window/mod.rs:23).
pub(crate) fn expect_ok(status: i32, what: &str) {
    if status != solid_gpu_sys::SGPU_OK {
        panic!("{} failed ({}): {}", what, status, last_error());
    }
}
