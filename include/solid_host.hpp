// Host-side types of the reference crate's filtering path that never touch the GPU: solid::window::Window<T>
// (window/mod.rs:9-126) and solid::circular_buffer::CircularBuffer<T> (circular_buffer/mod.rs:55-627).  Header-only,
// no CUDA, no libsolid_gpu.so: on the GPU path a filter's Window is the handle's history tail (sgpu_*_get_state /
// _set_state); these types keep the reference's API for callers that build their own pipelines, with the reference's
// index behaviour and error codes.  include/solid.hpp includes this file.
#pragma once

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace solid {

namespace window {

// Fixed-capacity shift register, newest element at index 0, zero-initialised (window/mod.rs:17-34).
template <typename T>
class Window {
   public:
    Window(size_t capacity, size_t delay = 0) : buffer_(capacity + delay, T{}), capacity_(capacity), delay_(delay) {  // :17
        if (capacity == 0) throw std::invalid_argument("Window: capacity must be > 0");                              // :18 assert
    }
    const T *as_ptr() const { return buffer_.data() + delay_; }                                                     // :36
    std::vector<T> to_vec() const { return std::vector<T>(buffer_.begin() + delay_, buffer_.begin() + delay_ + capacity_); }  // :44-51
    void reset() { std::fill(buffer_.begin(), buffer_.end(), T{}); }                                                // :54 (no leak)
    size_t capacity() const { return capacity_; }                                                                   // :59
    void push(T element) {  // :63-71: moves capacity - 1 elements up by one, writes index 0
        std::copy_backward(buffer_.begin(), buffer_.begin() + (capacity_ - 1), buffer_.begin() + capacity_);
        buffer_[0] = element;
    }
    void write(const std::vector<T> &other) { for (const T &e : other) push(e); }                                   // :73-77
    // the n most recent samples, oldest first: the layout of sgpu_*_get_state / _set_state
    std::vector<T> to_history(size_t n) const {
        std::vector<T> h(buffer_.begin(), buffer_.begin() + n);
        std::reverse(h.begin(), h.end());
        return h;
    }
    static Window from_history(const std::vector<T> &history, size_t capacity = 0) {
        Window w(capacity ? capacity : std::max<size_t>(history.size(), 1));
        w.write(history);
        return w;
    }

   private:
    std::vector<T> buffer_;
    size_t capacity_, delay_;
};

}  // namespace window

namespace circular_buffer {

enum class BufferErrorCode { EmptyBuffer, FullBuffer, NotEnoughBuffer, NegativeBuffer, NonExistantBuffer };  // :27-33

struct BufferError : std::runtime_error {  // :36-48
    BufferErrorCode code;
    explicit BufferError(BufferErrorCode c) : std::runtime_error("Buffer Error"), code(c) {}
};

// Ring FIFO with explicit read / write indices and `isize` sizes (circular_buffer/mod.rs:55-62).
template <typename T>
class CircularBuffer {
   public:
    explicit CircularBuffer(std::ptrdiff_t capacity) : buffer_((size_t)capacity, T{}), capacity_(capacity) {  // :79
        assert(capacity > 0);
    }
    static CircularBuffer from_vec(const std::vector<T> &v) {  // :114, :136 (from_slice)
        CircularBuffer cb((std::ptrdiff_t)v.size());
        cb.append(v);
        return cb;
    }
    static CircularBuffer from_slice(const std::vector<T> &v) { return from_vec(v); }
    const T *as_ptr() const { return buffer_.data(); }                // :164 -- the raw storage
    T *as_mut_ptr() { linearize(); return buffer_.data(); }           // :191 -- linearises first
    void linearize() {                                                // :220-238
        std::rotate(buffer_.begin(), buffer_.begin() + read_, buffer_.end());
        write_ = (write_ - read_) % capacity_;  // C++ and Rust `%` both keep the dividend's sign: can go negative (:235)
        read_ = 0;
    }
    std::vector<T> to_vec() const {                                   // :261-269: all `capacity` slots from read_index
        std::vector<T> v(buffer_.begin() + read_, buffer_.end());
        v.insert(v.end(), buffer_.begin(), buffer_.begin() + read_);
        return v;
    }
    void reset() { read_ = write_ = n_ = 0; }                         // :289
    std::ptrdiff_t len() const { return n_; }                         // :313
    std::ptrdiff_t capacity() const { return capacity_; }             // :326
    std::ptrdiff_t reserved() const { return capacity_ - n_; }        // :343
    bool is_empty() const { return n_ == 0; }                         // :357
    bool is_full() const { return n_ == capacity_; }                  // :375
    std::ptrdiff_t read_index() const { return read_; }               // :395
    std::ptrdiff_t write_index() const { return write_; }             // :414
    void push(T element) {                                            // :433-447
        if (is_full()) throw BufferError(BufferErrorCode::FullBuffer);
        buffer_[(size_t)(((write_ % capacity_) + capacity_) % capacity_)] = element;
        write_ = (write_ + 1) % capacity_;
        ++n_;
    }
    void append(const std::vector<T> &other) {                        // :469-494
        const std::ptrdiff_t k = (std::ptrdiff_t)other.size();
        if (n_ + k > capacity_) throw BufferError(BufferErrorCode::NotEnoughBuffer);
        const std::ptrdiff_t room = capacity_ - write_;
        if (k <= room) {
            std::copy(other.begin(), other.end(), buffer_.begin() + write_);
        } else {
            std::copy(other.begin(), other.begin() + room, buffer_.begin() + write_);
            // reference quirk kept (:486-490): the wrapped part is copied from offset k - room, not from offset room
            // (identical only when k == 2 * room); reads past the slice are clipped
            const std::ptrdiff_t from = k - room, cnt = std::min<std::ptrdiff_t>(k - room, k - from);
            std::copy(other.begin() + from, other.begin() + from + cnt, buffer_.begin());
        }
        write_ = (write_ + k) % capacity_;
        n_ += k;
    }
    T pop() {                                                         // :512-524
        if (is_empty()) throw BufferError(BufferErrorCode::EmptyBuffer);
        const T v = buffer_[(size_t)read_];
        read_ = (read_ + 1) % capacity_;
        --n_;
        return v;
    }
    void release(std::ptrdiff_t n) {                                  // :548-557
        if (n < 0) throw BufferError(BufferErrorCode::NegativeBuffer);
        if (n > n_) throw BufferError(BufferErrorCode::NotEnoughBuffer);
        read_ = (read_ + n) % capacity_;
        n_ -= n;
    }
    std::vector<T> deref() const { return std::vector<T>(buffer_.begin(), buffer_.begin() + n_); }  // Deref<[T]> :603-610

   private:
    std::vector<T> buffer_;
    std::ptrdiff_t capacity_, read_ = 0, write_ = 0, n_ = 0;
};

}  // namespace circular_buffer
}  // namespace solid
