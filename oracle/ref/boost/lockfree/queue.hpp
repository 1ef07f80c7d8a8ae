/*
 * Minimal stand-in for <boost/lockfree/queue.hpp> (Boost is not installed in this image).
 *
 * TEST INFRASTRUCTURE ONLY: this header exists so that the UNMODIFIED reference source
 * Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp can be compiled where it lies under
 * /root/reference into oracle/_ref/.  It provides exactly the members that file uses
 * (radix_join.cpp:110 type alias, :831 bounded_push, :1209 push, :1261/:1321 pop,
 * :1412-1413 size_type constructor).  The queue carries <= 2^(bits+1) task descriptors per
 * join, so a mutex-protected ring is more than fast enough and does not perturb timing of
 * the partition/build/probe loops themselves.
 */
#pragma once
#include <cstddef>
#include <deque>
#include <mutex>

namespace boost { namespace lockfree {

template <bool B> struct fixed_sized { static constexpr bool value = B; };

template <typename T, typename... Options>
class queue {
public:
    using size_type = std::size_t;
    explicit queue(size_type capacity) : capacity_(capacity) {}
    bool push(const T &t) {
        std::lock_guard<std::mutex> g(m_);
        q_.push_back(t);
        return true;
    }
    bool bounded_push(const T &t) {
        std::lock_guard<std::mutex> g(m_);
        if (q_.size() >= capacity_) return false;
        q_.push_back(t);
        return true;
    }
    bool pop(T &out) {
        std::lock_guard<std::mutex> g(m_);
        if (q_.empty()) return false;
        out = q_.front();
        q_.pop_front();
        return true;
    }
    bool empty() {
        std::lock_guard<std::mutex> g(m_);
        return q_.empty();
    }
private:
    std::mutex m_;
    std::deque<T> q_;
    size_type capacity_;
};

}}  // namespace boost::lockfree
