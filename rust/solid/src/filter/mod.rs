//! reference: src/filter/mod.rs:9-22 -- the plugin trait of the path and its implementors.
use num::complex::Complex;

/// trait Filter<I, O> -- filter/mod.rs:9-22
pub trait Filter<I, O> {
    fn execute(&mut self, sample: I) -> Vec<O>;
    fn execute_block(&mut self, samples: &[I]) -> Vec<O>;
    fn frequency_response(&self, frequency: f64) -> Complex<f64>;
    fn group_delay(&self, frequency: f64) -> f64;
}

pub mod fir;
pub mod iir;
pub mod auto_correlator;
pub mod ddc;
pub mod firdes;
