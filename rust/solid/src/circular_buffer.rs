//! reference: src/circular_buffer/mod.rs -- ring FIFO with explicit indices.  No filter uses it (SURVEY section
//! 2 row 3); kept as API with the same method names, `isize` sizes, index behaviour and error codes.
use std::error::Error;
use std::fmt;
use std::ops::{Deref, DerefMut};

/// circular_buffer/mod.rs:27-33
#[derive(Debug, PartialEq, Eq)]
pub enum BufferErrorCode { EmptyBuffer, FullBuffer, NotEnoughBuffer, NegativeBuffer, NonExistantBuffer }
/// circular_buffer/mod.rs:36
#[derive(Debug)]
pub struct BufferError(pub BufferErrorCode);
impl fmt::Display for BufferError {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "Buffer Error {:?}", self.0) }
}
impl Error for BufferError {}

/// CircularBuffer<T> -- circular_buffer/mod.rs:55-62
#[derive(Debug, Clone)]
pub struct CircularBuffer<T: Copy + Default> {
    buffer: Vec<T>,
    capacity: isize,
    read_index: isize,
    write_index: isize,
    num_elements: isize,
}

impl<T: Copy + Default> CircularBuffer<T> {
    /// :79 (asserts capacity > 0)
    pub fn new(capacity: isize) -> Self {
        assert!(capacity > 0);
        CircularBuffer { buffer: vec![T::default(); capacity as usize], capacity, read_index: 0, write_index: 0, num_elements: 0 }
    }
    /// :114
    pub fn from_vec(vec: Vec<T>) -> Self { Self::from_slice(&vec) }
    /// :136
    pub fn from_slice(slice: &[T]) -> Self {
        let mut cb = Self::new(slice.len() as isize);
        cb.append(slice).expect("capacity equals the slice length");
        cb
    }
    /// :164
    pub fn as_ptr(&self) -> *const T { self.buffer.as_ptr() }
    /// :191 -- linearises first
    pub fn as_mut_ptr(&mut self) -> *mut T { self.linearize(); self.buffer.as_mut_ptr() }
    /// :220-238
    pub fn linearize(&mut self) {
        self.buffer.rotate_left(self.read_index as usize);
        // Rust's % keeps the dividend's sign: the reference can leave a negative write index here (:235)
        self.write_index = (self.write_index - self.read_index) % self.capacity;
        self.read_index = 0;
    }
    /// :261-269: all `capacity` slots starting at read_index
    pub fn to_vec(&self) -> Vec<T> {
        let mut v = self.buffer.clone();
        v.rotate_left(self.read_index as usize);
        v
    }
    /// :289
    pub fn reset(&mut self) { self.read_index = 0; self.write_index = 0; self.num_elements = 0; }
    /// :313
    pub fn len(&self) -> isize { self.num_elements }
    /// :326
    pub fn capacity(&self) -> isize { self.capacity }
    /// :343
    pub fn reserved(&self) -> isize { self.capacity - self.num_elements }
    /// :357
    pub fn is_empty(&self) -> bool { self.num_elements == 0 }
    /// :375
    pub fn is_full(&self) -> bool { self.num_elements == self.capacity }
    /// :395
    pub fn read_index(&self) -> isize { self.read_index }
    /// :414
    pub fn write_index(&self) -> isize { self.write_index }
    /// :433-447
    pub fn push(&mut self, element: T) -> Result<(), Box<dyn Error>> {
        if self.is_full() { return Err(Box::new(BufferError(BufferErrorCode::FullBuffer))); }
        let w = self.write_index.rem_euclid(self.capacity) as usize;
        self.buffer[w] = element;
        self.write_index = (self.write_index + 1) % self.capacity;
        self.num_elements += 1;
        Ok(())
    }
    /// :469-494
    pub fn append(&mut self, other: &[T]) -> Result<(), Box<dyn Error>> {
        let k = other.len() as isize;
        if self.num_elements + k > self.capacity { return Err(Box::new(BufferError(BufferErrorCode::NotEnoughBuffer))); }
        let w = self.write_index.rem_euclid(self.capacity) as usize;
        let room = self.capacity as usize - w;
        if other.len() <= room {
            self.buffer[w..w + other.len()].copy_from_slice(other);
        } else {
            self.buffer[w..].copy_from_slice(&other[..room]);
            // reference quirk kept (:486-490): the wrapped part is copied from offset k - room, not from offset room
            let rest = other.len() - room;
            let from = rest.min(other.len());
            let n = rest.min(other.len() - from);
            self.buffer[..n].copy_from_slice(&other[from..from + n]);
        }
        self.write_index = (self.write_index + k) % self.capacity;
        self.num_elements += k;
        Ok(())
    }
    /// :512-524
    pub fn pop(&mut self) -> Result<T, Box<dyn Error>> {
        if self.is_empty() { return Err(Box::new(BufferError(BufferErrorCode::EmptyBuffer))); }
        let v = self.buffer[self.read_index as usize];
        self.read_index = (self.read_index + 1) % self.capacity;
        self.num_elements -= 1;
        Ok(v)
    }
    /// :548-557
    pub fn release(&mut self, n: isize) -> Result<(), Box<dyn Error>> {
        if n < 0 { return Err(Box::new(BufferError(BufferErrorCode::NegativeBuffer))); }
        if n > self.num_elements { return Err(Box::new(BufferError(BufferErrorCode::NotEnoughBuffer))); }
        self.read_index = (self.read_index + n) % self.capacity;
        self.num_elements -= n;
        Ok(())
    }
}
/// :603-610: the first len() storage slots
impl<T: Copy + Default> Deref for CircularBuffer<T> {
    type Target = [T];
    fn deref(&self) -> &[T] { &self.buffer[..self.num_elements as usize] }
}
impl<T: Copy + Default> DerefMut for CircularBuffer<T> {
    fn deref_mut(&mut self) -> &mut [T] { let n = self.num_elements as usize; &mut self.buffer[..n] }
}
impl<T: Copy + Default + fmt::Display> fmt::Display for CircularBuffer<T> {
    /// :619-627
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "CircularBuffer<{}> [", std::any::type_name::<T>())?;
        for i in 0..self.num_elements {
            if i > 0 { write!(f, ", ")?; }
            write!(f, "{}", self.buffer[((self.read_index + i) % self.capacity) as usize])?;
        }
        write!(f, "]")
    }
}
