// Host-only types of the C++ mirror (include/solid_host.hpp): Window<T> and CircularBuffer<T> against the reference's
// behaviour -- the doc-tests of circular_buffer/mod.rs:425-431, 461-467, 504-510, 536-545 and the same cases
// tests/test_host_types.py runs on the Python types.  No GPU, no libsolid_gpu.so.  Exit code 0 = all passed.
#include <complex>
#include <cstdio>

#include "solid_host.hpp"

using solid::circular_buffer::BufferError;
using solid::circular_buffer::BufferErrorCode;
using solid::circular_buffer::CircularBuffer;
using solid::window::Window;

static int failures = 0;
#define EXPECT(cond, what)                                            \
    do {                                                              \
        if (!(cond)) {                                                \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, what); \
            ++failures;                                               \
        }                                                             \
    } while (0)

template <typename F>
static bool throws_code(F f, BufferErrorCode want) {
    try {
        f();
    } catch (const BufferError &e) {
        return e.code == want;
    }
    return false;
}

int main() {
    {  // Window: newest at index 0 (window/mod.rs:63-71), to_vec (:44-51), reset, history bridge
        Window<double> w(4);
        w.write({1.0, 2.0, 3.0, 4.0, 5.0});
        EXPECT(w.to_vec() == std::vector<double>({5.0, 4.0, 3.0, 2.0}), "window order");
        EXPECT(w.capacity() == 4 && w.as_ptr()[0] == 5.0, "window accessors");
        Window<double> c = w;  // Clone
        w.reset();
        EXPECT(w.to_vec() == std::vector<double>(4, 0.0) && c.to_vec() == std::vector<double>({5.0, 4.0, 3.0, 2.0}), "window reset / clone");
        EXPECT(Window<double>::from_history({7.0, 8.0, 9.0}).to_vec() == std::vector<double>({9.0, 8.0, 7.0}), "from_history");
        EXPECT(c.to_history(3) == std::vector<double>({3.0, 4.0, 5.0}), "to_history: oldest first");
        Window<std::complex<float>> d(3, 2);  // capacity 3, delay 2: to_vec starts at the delay and is never written there
        d.push({1.f, 2.f});
        EXPECT(d.to_vec().size() == 3 && d.to_vec()[0] == std::complex<float>(0.f, 0.f), "delayed window");
        bool threw = false;
        try { Window<int> z(0); } catch (const std::invalid_argument &) { threw = true; }
        EXPECT(threw, "capacity 0 rejected");
    }
    {  // CircularBuffer
        CircularBuffer<unsigned char> b(4);
        b.append({2, 3, 4, 5});
        EXPECT(throws_code([&] { b.append({6}); }, BufferErrorCode::NotEnoughBuffer), "append to a full buffer");
        EXPECT(throws_code([&] { b.push(1); }, BufferErrorCode::FullBuffer), "push to a full buffer");
        EXPECT(b.pop() == 2 && b.len() == 3 && b.read_index() == 1, "pop");
        b.push(9);
        EXPECT(b.write_index() == 1 && b.is_full(), "wrapped write index");
        EXPECT(b.to_vec() == std::vector<unsigned char>({3, 4, 5, 9}), "to_vec from read_index");
        b.release(2);
        EXPECT(b.len() == 2 && b.reserved() == 2, "release");
        EXPECT(throws_code([&] { b.release(-1); }, BufferErrorCode::NegativeBuffer), "negative release");
        EXPECT(throws_code([&] { b.release(5); }, BufferErrorCode::NotEnoughBuffer), "release more than held");
        b.reset();
        EXPECT(b.is_empty(), "reset");
        EXPECT(throws_code([&] { b.pop(); }, BufferErrorCode::EmptyBuffer), "pop from an empty buffer");
        auto c = CircularBuffer<int>::from_vec({1, 2, 3});
        EXPECT(c.capacity() == 3 && c.is_full() && c.deref() == std::vector<int>({1, 2, 3}), "from_vec");
        c.pop();
        c.linearize();
        EXPECT(c.read_index() == 0 && std::vector<int>(c.as_ptr(), c.as_ptr() + 3) == std::vector<int>({2, 3, 1}), "linearize");
        // the reference's negative write index after linearize (circular_buffer/mod.rs:235): read 2, write 1 -> -1
        CircularBuffer<int> q(4);
        q.append({1, 2, 3, 4});
        q.pop();
        q.pop();
        q.push(5);  // write index 1, read index 2
        q.linearize();
        EXPECT(q.write_index() == -1 && q.read_index() == 0, "linearize keeps the reference's sign behaviour");
        // wrapped append: the reference copies the wrapped part from offset k - room (:486-490)
        CircularBuffer<int> r(4);
        r.append({1, 2, 3});
        r.release(3);  // read 3, write 3, empty
        r.append({7, 8, 9});  // room 1: slot 3 = 7; wrapped part from offset 2: slot 0 = 9 (slot 1 untouched)
        EXPECT(r.len() == 3 && r.as_ptr()[3] == 7 && r.as_ptr()[0] == 9 && r.write_index() == 2, "wrapped append quirk");
    }
    std::printf("%s (%d failure%s)\n", failures ? "FAILED" : "PASSED", failures, failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
