//! reference: src/dot_product/mod.rs, src/dot_product/execute.rs
use crate::scalar::{Coefficient, Sample};
use num::complex::Complex;
use solid_gpu_sys as sys;
use std::fmt;
use std::marker::PhantomData;
use std::ptr;

/// dot_product/mod.rs:31-34
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Direction { FORWARD, REVERSE }

pub mod execute {
    /// dot_product/execute.rs:1-18
    pub trait Execute<I, O> { fn execute(&self, samples: &[I]) -> O; }
}

/// DotProduct<T> -- dot_product/mod.rs:37-42
pub struct DotProduct<T: Coefficient> { h: *mut sys::sgpu_dot, direction: Direction, _t: PhantomData<T> }

impl<T: Coefficient> DotProduct<T> {
    /// dot_product/mod.rs:57
    pub fn new(coefficients: &[T], direction: Direction) -> Self {
        let flat = T::flatten(coefficients);
        let mut h = ptr::null_mut();
        let dir = match direction { Direction::FORWARD => sys::SGPU_FORWARD, Direction::REVERSE => sys::SGPU_REVERSE };
        let st = unsafe { sys::sgpu_dot_create(flat.as_ptr(), coefficients.len(), T::KIND, dir, &mut h) };
        crate::expect_ok(st, "sgpu_dot_create");
        DotProduct { h, direction, _t: PhantomData }
    }
    /// [sic] dot_product/mod.rs:102 -- the STORED order (reversed for REVERSE)
    pub fn coefficents(&self) -> Vec<T> {
        let mut flat = vec![0.0f64; self.len() * T::WIDTH];
        if !flat.is_empty() { unsafe { sys::sgpu_dot_coefficients(self.h, flat.as_mut_ptr()) }; }
        T::unflatten(&flat)
    }
    /// dot_product/mod.rs:124
    pub fn len(&self) -> usize { unsafe { sys::sgpu_dot_len(self.h) } }
    /// dot_product/mod.rs:141
    pub fn is_empty(&self) -> bool { self.len() == 0 }
}

impl<T: Coefficient, I: Sample> execute::Execute<I, I> for DotProduct<T> {
    /// Execute::execute -- dot_product/mod.rs:153-171: sum over min(len_c, len_x) terms
    fn execute(&self, samples: &[I]) -> I {
        let x = I::narrow(samples);
        let mut r = Complex::new(0f32, 0f32);
        let st = unsafe {
            sys::sgpu_dot_execute(self.h, x.as_ptr() as *const f32, x.len(), x.len().max(1), 1,
                                  &mut r as *mut Complex<f32> as *mut f32, sys::SGPU_HOST, ptr::null_mut())
        };
        crate::expect_ok(st, "sgpu_dot_execute");
        I::from_cf32(r)
    }
}

impl<T: Coefficient> Clone for DotProduct<T> {
    /// dot_product/mod.rs:173-190 (deep copy)
    fn clone(&self) -> Self {
        // coefficents() returns the stored order: a FORWARD object built from it stores the same values
        let stored = self.coefficents();
        let mut c = DotProduct::new(&stored, Direction::FORWARD);
        c.direction = self.direction;
        c
    }
}
impl<T: Coefficient> Drop for DotProduct<T> {
    /// dot_product/mod.rs:192-196
    fn drop(&mut self) { unsafe { sys::sgpu_dot_destroy(self.h) }; }
}
impl<T: Coefficient> fmt::Debug for DotProduct<T> {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "DotProduct [Size={}] [{:?}]", self.len(), self.direction) }
}
impl<T: Coefficient> fmt::Display for DotProduct<T> {
    /// dot_product/mod.rs:146-151
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "DotProduct<{}> [Size={}]", std::any::type_name::<T>(), self.len())
    }
}
