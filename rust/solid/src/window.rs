//! reference: src/window/mod.rs -- fixed-capacity shift register, newest element at index 0.
//! Host-side type (on the GPU path the same state is the handle's history tail); the reference's leaks in
//! `reset` / `as_ptr` and the dropped delay tail in `clone` are not reproduced.
use std::fmt;

/// Window<T> -- window/mod.rs:9-14
#[derive(Debug, Clone)]
pub struct Window<T: Copy + Default> { buffer: Vec<T>, capacity: usize, delay: usize }

impl<T: Copy + Default> Window<T> {
    /// window/mod.rs:17 (asserts capacity > 0, :18)
    pub fn new(capacity: usize, delay: usize) -> Self {
        assert!(capacity > 0);
        Window { buffer: vec![T::default(); capacity + delay], capacity, delay }
    }
    /// window/mod.rs:36 -- a pointer to `capacity` elements starting at `delay`
    pub fn as_ptr(&self) -> *const T { self.buffer[self.delay..].as_ptr() }
    /// window/mod.rs:44-51
    pub fn to_vec(&self) -> Vec<T> { self.buffer[self.delay..self.delay + self.capacity].to_vec() }
    /// window/mod.rs:54
    pub fn reset(&mut self) { for v in self.buffer.iter_mut() { *v = T::default(); } }
    /// window/mod.rs:59
    pub fn capacity(&self) -> usize { self.capacity }
    /// window/mod.rs:63-71: moves capacity - 1 elements up by one, writes index 0
    pub fn push(&mut self, element: T) {
        self.buffer.copy_within(0..self.capacity - 1, 1);
        self.buffer[0] = element;
    }
    /// window/mod.rs:73-77
    pub fn write(&mut self, other: &[T]) { for e in other { self.push(*e); } }
    /// The `n` most recent samples, oldest first: the layout of sgpu_*_get_state / _set_state.
    pub fn to_history(&self, n: usize) -> Vec<T> { self.buffer[..n].iter().rev().copied().collect() }
}
impl<T: Copy + Default> fmt::Display for Window<T> {
    /// window/mod.rs:90-100
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "Window<{}> [Capacity={}] [Delay={}]", std::any::type_name::<T>(), self.capacity, self.delay)
    }
}
