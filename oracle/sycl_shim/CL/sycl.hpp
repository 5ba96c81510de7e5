// oracle/sycl_shim/CL/sycl.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A few hundred lines of "SYCL for FPGA" surface, implemented with std::thread, so that the
// reference's OWN device code (device/keyswitch.cpp + device/keyswitch/*.hpp + device/mod_ops.hpp,
// written for dpcpp -fintelfpga) compiles UNMODIFIED with g++ and runs on the CPU -- the role the
// oneAPI FPGA emulator plays for the reference (RUN_CHOICE=1, host/src/fpga.cpp:1615-1623).
// Built only by oracle/Makefile into oracle/_ref/; nothing in the product includes or links it.
//
// What is modelled (exactly what the keyswitch sources use):
//   sycl::queue::submit + handler::single_task   one host thread per kernel (autorun kernels spin forever)
//   sycl::ext::intel::pipe<Id, T, depth>          bounded FIFO with blocking and non-blocking read / write
//   sycl::buffer / accessor / device_ptr          thin views of host memory
//   sycl::ulong2 / ulong4                          swizzle accessors s0()..s3()
//   [[intel::...]] attributes, #pragma unroll       ignored by g++ (they do not change results)
#pragma once
#include <sys/types.h>

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace sycl {

// ---- vector types -----------------------------------------------------------
template <int N>
struct ulongN {
    unsigned long v[N];
    ulongN() {
        for (int i = 0; i < N; ++i) v[i] = 0;
    }
    ulongN(unsigned long x) {   // broadcast, as sycl::vec(const T&)
        for (int i = 0; i < N; ++i) v[i] = x;
    }
    unsigned long& operator[](size_t i) { return v[i]; }
    const unsigned long& operator[](size_t i) const { return v[i]; }
    unsigned long& s0() { return v[0]; }
    unsigned long& s1() { return v[1]; }
    template <int M = N>
    typename std::enable_if<(M > 2), unsigned long&>::type s2() { return v[2]; }
    template <int M = N>
    typename std::enable_if<(M > 3), unsigned long&>::type s3() { return v[3]; }
    const unsigned long& s0() const { return v[0]; }
    const unsigned long& s1() const { return v[1]; }
    template <int M = N>
    typename std::enable_if<(M > 2), const unsigned long&>::type s2() const { return v[2]; }
    template <int M = N>
    typename std::enable_if<(M > 3), const unsigned long&>::type s3() const { return v[3]; }
};
using ulong2 = ulongN<2>;
using ulong4 = ulongN<4>;
using ulong8 = ulongN<8>;
using ulong = unsigned long;

// ---- events / queue / handler ------------------------------------------------
struct task_state {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
};

class event {
public:
    event() = default;
    explicit event(std::shared_ptr<task_state> s) : st_(std::move(s)) {}
    void wait() {
        if (!st_) return;
        std::unique_lock<std::mutex> lk(st_->mu);
        st_->cv.wait(lk, [&] { return st_->done; });
    }

private:
    std::shared_ptr<task_state> st_;
};

class handler {
public:
    template <class Name = void, class F>
    void single_task(F f) {
        st_ = std::make_shared<task_state>();
        std::shared_ptr<task_state> st = st_;
        // kernels keep large arrays on their stack (e.g. 7 twiddle tables): give them room
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, (size_t)64 << 20);
        auto* fn = new std::function<void()>([f, st]() mutable {
            f();
            std::lock_guard<std::mutex> lk(st->mu);
            st->done = true;
            st->cv.notify_all();
        });
        pthread_t th;
        pthread_create(&th, &attr, [](void* p) -> void* {
            auto* g = static_cast<std::function<void()>*>(p);
            (*g)();
            delete g;
            return nullptr;
        }, fn);
        pthread_detach(th);
        pthread_attr_destroy(&attr);
    }
    std::shared_ptr<task_state> st_;
};

class queue {
public:
    template <class F>
    event submit(F f) {
        handler h;
        f(h);
        return event(h.st_);
    }
};

// ---- memory views ---------------------------------------------------------------
struct read_only_t {};
struct write_only_t {};
struct no_init_t {};
static constexpr read_only_t read_only{};
static constexpr write_only_t write_only{};
static constexpr no_init_t no_init{};

template <class T>
class buffer {
public:
    buffer(T* p, size_t n) : p_(p), n_(n) {}
    T* data() const { return p_; }
    size_t size() const { return n_; }

private:
    T* p_;
    size_t n_;
};

template <class T>
class accessor {
public:
    template <class... Tags>
    accessor(buffer<T>& b, handler&, Tags...) : p_(b.data()) {}
    T& operator[](size_t i) const { return p_[i]; }
    T* get_pointer() const { return p_; }

private:
    T* p_;
};
template <class T, class... Tags>
accessor(buffer<T>&, handler&, Tags...) -> accessor<T>;

template <class T>
class device_ptr {
public:
    device_ptr(const accessor<T>& a) : p_(a.get_pointer()) {}
    explicit device_ptr(T* p) : p_(p) {}
    T& operator[](size_t i) const { return p_[i]; }
    T& operator*() const { return *p_; }
    explicit operator T*() const { return p_; }

private:
    T* p_;
};
// USM: host and device allocations are both plain host memory here
template <class T>
using host_ptr = device_ptr<T>;

enum class memory_order { relaxed, acquire, release, acq_rel, seq_cst };
enum class memory_scope { work_item, sub_group, work_group, device, system };
inline void atomic_fence(memory_order, memory_scope) { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---- pipes ------------------------------------------------------------------------
namespace ext {
namespace intel {

// One FIFO per (Id, T, depth) instantiation.  A pipe of the bitstream holds at least `depth`
// entries; a deeper FIFO can only remove stalls, never change a result, and the non-blocking
// writers of the twiddle dispatcher need a BOUND (they stream the tables cyclically for as long
// as there is room), so the capacity is max(depth, 64).
template <class Id, class T, size_t depth = 0>
class pipe {
    struct fifo {
        std::mutex mu;
        std::condition_variable not_empty, not_full;
        std::deque<T> q;
    };
    static fifo& f() {
        static fifo inst;
        return inst;
    }
    static constexpr size_t cap = depth > 64 ? depth : 64;

public:
    static T read() {
        fifo& x = f();
        std::unique_lock<std::mutex> lk(x.mu);
        x.not_empty.wait(lk, [&] { return !x.q.empty(); });
        T v = x.q.front();
        x.q.pop_front();
        lk.unlock();
        x.not_full.notify_one();
        return v;
    }
    static T read(bool& ok) {
        fifo& x = f();
        std::unique_lock<std::mutex> lk(x.mu);
        if (x.q.empty()) {
            ok = false;
            lk.unlock();
            std::this_thread::yield();
            return T();
        }
        T v = x.q.front();
        x.q.pop_front();
        ok = true;
        lk.unlock();
        x.not_full.notify_one();
        return v;
    }
    static void write(const T& v) {
        fifo& x = f();
        std::unique_lock<std::mutex> lk(x.mu);
        x.not_full.wait(lk, [&] { return x.q.size() < cap; });
        x.q.push_back(v);
        lk.unlock();
        x.not_empty.notify_one();
    }
    static void write(const T& v, bool& ok) {
        fifo& x = f();
        std::unique_lock<std::mutex> lk(x.mu);
        if (x.q.size() >= cap) {
            ok = false;
            lk.unlock();
            std::this_thread::yield();
            return;
        }
        x.q.push_back(v);
        ok = true;
        lk.unlock();
        x.not_empty.notify_one();
    }
};

}  // namespace intel
}  // namespace ext
}  // namespace sycl

namespace ext = sycl::ext;   // kernel_assert.hpp names ext::oneapi under EMULATOR only (not defined here)
