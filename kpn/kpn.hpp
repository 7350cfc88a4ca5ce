// kpn.hpp -- the Kahn-process-network block/port contract of LibRedio's `kpn` crate, in C++17.
//
// The reference (src/kpn/src/kpn.rs, src/ratpak.rs:31-185): a block is a free function whose first
// arguments are input Receiver<T>(s), then output Sender<U>(s), then scalar parameters; it loops forever
// `v.send(f(u.recv().unwrap())).unwrap()`; every block runs on its own OS thread; ports are UNBOUNDED
// channels with non-blocking send and blocking recv (README.mkd:3); messages are MOVED; a closed port makes
// `.unwrap()` panic, which drops the block's own ports and so tears the graph down (kpn.rs:18-28).
//
// Here: channel<T>() -> {Sender<T>, Receiver<T>} over a mutex/condvar queue; recv()/send() throw
// PortClosed where the Rust would panic; spawn() runs a block on its own thread and swallows PortClosed
// (thread death = teardown signal), exactly the reference's convention.  The CPU-side blocks the OOK chain
// needs (rle, dle, looper, shaper_optional, binconv, eat, b2d, fork ...) are restated below so an
// existing graph can be wired unchanged around the GPU blocks of gpu_blocks.hpp.
#pragma once
#include <condition_variable>
#include <cstdio>
#include <cstddef>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

namespace kpn {

struct PortClosed : std::runtime_error {
    PortClosed() : std::runtime_error("called `Result::unwrap()` on an `Err` value: port closed") {}
};

template <class T>
struct Chan {
    std::mutex m;
    std::condition_variable cv;
    std::deque<T> q;
    size_t senders = 0;
    bool receiver_alive = true;
};

template <class T>
class Sender {
    std::shared_ptr<Chan<T>> c_;
public:
    Sender() = default;
    explicit Sender(std::shared_ptr<Chan<T>> c) : c_(std::move(c)) { std::lock_guard<std::mutex> l(c_->m); ++c_->senders; }
    Sender(const Sender &o) : c_(o.c_) { if (c_) { std::lock_guard<std::mutex> l(c_->m); ++c_->senders; } }   // tx.clone()
    Sender(Sender &&o) noexcept : c_(std::move(o.c_)) {}
    Sender &operator=(Sender o) { std::swap(c_, o.c_); return *this; }
    ~Sender() { drop(); }
    void drop()
    {
        if (!c_) return;
        { std::lock_guard<std::mutex> l(c_->m); --c_->senders; }
        c_->cv.notify_all();
        c_.reset();
    }
    // never blocks (unbounded queue); throws if the receiving end is gone, like send().unwrap()
    void send(T v) const
    {
        if (!c_) throw PortClosed();
        {
            std::lock_guard<std::mutex> l(c_->m);
            if (!c_->receiver_alive) throw PortClosed();
            c_->q.push_back(std::move(v));
        }
        c_->cv.notify_one();
    }
};

template <class T>
class Receiver {
    std::shared_ptr<Chan<T>> c_;
public:
    Receiver() = default;
    explicit Receiver(std::shared_ptr<Chan<T>> c) : c_(std::move(c)) {}
    Receiver(const Receiver &) = delete;                      // a Receiver has one owner
    Receiver(Receiver &&o) noexcept : c_(std::move(o.c_)) {}
    Receiver &operator=(Receiver &&o) noexcept { drop(); c_ = std::move(o.c_); return *this; }
    ~Receiver() { drop(); }
    void drop()
    {
        if (!c_) return;
        { std::lock_guard<std::mutex> l(c_->m); c_->receiver_alive = false; c_->q.clear(); }
        c_.reset();
    }
    // blocks; throws PortClosed once the queue is empty and every Sender is gone (recv().unwrap())
    T recv() const
    {
        if (!c_) throw PortClosed();
        std::unique_lock<std::mutex> l(c_->m);
        c_->cv.wait(l, [&] { return !c_->q.empty() || c_->senders == 0; });
        if (c_->q.empty()) throw PortClosed();
        T v = std::move(c_->q.front());
        c_->q.pop_front();
        return v;
    }
    // Ok(x) / Err(Empty|Disconnected) collapsed to optional; never blocks
    std::optional<T> try_recv() const
    {
        if (!c_) return std::nullopt;
        std::lock_guard<std::mutex> l(c_->m);
        if (c_->q.empty()) return std::nullopt;
        T v = std::move(c_->q.front());
        c_->q.pop_front();
        return v;
    }
    // number of queued messages right now (used by GPU blocks to batch whatever has already arrived)
    size_t pending() const
    {
        if (!c_) return 0;
        std::lock_guard<std::mutex> l(c_->m);
        return c_->q.size();
    }
};

template <class T>
std::pair<Sender<T>, Receiver<T>> channel()
{
    auto c = std::make_shared<Chan<T>>();
    return {Sender<T>(c), Receiver<T>(c)};
}

// one named task per block (ratpak.rs:60-185).  A Rust panic kills only the panicking thread, whose
// ports are then dropped: a block that throws -- on a closed port, a failed assert, or a non-zero status of
// the GPU library -- ends its thread the same way, and the teardown cascades through the dropped ports.
template <class F>
std::thread spawn(F &&f)
{
    return std::thread([fn = std::forward<F>(f)]() mutable {
        try { fn(); }
        catch (const PortClosed &) {}
        catch (const std::exception &e) { std::fprintf(stderr, "kpn: block thread panicked: %s\n", e.what()); }
    });
}

// ---- CPU-side blocks of kpn.rs used around the hot path -------------------------------------------------

// kpn.rs:17-29  run length encoding: a run is emitted when the value changes; the last run never is
template <class T>
void rle(Receiver<T> u, Sender<std::pair<T, size_t>> v)
{
    T x = u.recv();
    size_t i = 1;
    for (;;) {
        T y = u.recv();
        if (y != x) { v.send({x, i}); i = 1; } else { i = i + 1; }
        x = std::move(y);
    }
}

// kpn.rs:32-38  counts -> seconds: ct as f32 / s_rate as f32
template <class T>
void dle(Receiver<std::pair<T, size_t>> u, Sender<std::pair<T, float>> v, size_t s_rate)
{
    for (;;) {
        auto p = u.recv();
        v.send({p.first, (float)p.second / (float)s_rate});
    }
}

// kpn.rs:50-56  run length decoding
template <class T>
void rld(Receiver<std::pair<T, size_t>> u, Sender<T> v)
{
    for (;;) {
        auto p = u.recv();
        for (size_t k = 0; k < p.second; ++k) v.send(p.first);
    }
}

// kpn.rs:111-113  MSB-first binary digits -> unsigned
inline size_t b2d(const std::vector<size_t> &xs)
{
    size_t s = 0;
    for (size_t i = 0; i < xs.size(); ++i) s += ((size_t)1 << (xs.size() - i - 1)) * xs[i];
    return s;
}

// kpn.rs:116-124  split by bit widths (out-of-range slices panic in the reference: throw here)
inline std::vector<size_t> eat(const std::vector<size_t> &x, const std::vector<size_t> &is)
{
    size_t i = 0;
    std::vector<size_t> out;
    for (size_t w : is) {
        if (i + w > x.size()) throw std::out_of_range("eat: slice index out of range");
        out.push_back(b2d(std::vector<size_t>(x.begin() + i, x.begin() + i + w)));
        i += w;
    }
    return out;
}

// kpn.rs:127-131 / :163-167  map a function across a stream
template <class T, class U, class F>
void cross_applicator(Receiver<T> u, Sender<U> v, F f)
{
    for (;;) v.send(f(u.recv()));
}
template <class T, class F>
void applicator(Receiver<T> u, Sender<T> v, F f) { cross_applicator<T, T, F>(std::move(u), std::move(v), f); }

// kpn.rs:170-174  map across Vec chunks
template <class T, class U, class F>
void cross_applicator_vecs(Receiver<std::vector<T>> u, Sender<std::vector<U>> v, F f)
{
    for (;;) {
        std::vector<T> in = u.recv();
        std::vector<U> out;
        out.reserve(in.size());
        for (const T &x : in) out.push_back(f(x));
        v.send(std::move(out));
    }
}

// kpn.rs:148-150  hand the whole input iterator to f.  `next()` yields nullopt at end of stream.
template <class T>
struct Iter {
    Receiver<T> *r;
    std::optional<T> next()
    {
        try { return r->recv(); } catch (const PortClosed &) { return std::nullopt; }
    }
};
template <class T, class U, class F>
void looper(Receiver<T> u, Sender<U> v, F f) { Iter<T> it{&u}; f(it, v); }

// kpn.rs:153-160  drop the Nones
template <class T>
void looper_optional(Receiver<std::optional<T>> u, Sender<T> v)
{
    for (;;) { auto x = u.recv(); if (x) v.send(std::move(*x)); }
}

// kpn.rs:182-189  duplicate a stream
template <class T>
void fork(Receiver<T> u, std::vector<Sender<T>> v)
{
    for (;;) { T x = u.recv(); for (auto &y : v) y.send(x); }
}

// kpn.rs:95-101  Vec<T> -> T
template <class T>
void unpacketizer(Receiver<std::vector<T>> u, Sender<T> v)
{
    for (;;) { for (auto &x : u.recv()) v.send(std::move(x)); }
}

// kpn.rs:266-275  collect Some(y); on None emit iff exactly l collected, then clear
template <class T>
void shaper_optional(Receiver<std::optional<T>> u, Sender<std::vector<T>> v, size_t l)
{
    std::vector<T> x;
    for (;;) {
        auto y = u.recv();
        if (y) x.push_back(std::move(*y));
        else if (x.size() == l) { v.send(x); x.clear(); }
        else x.clear();
    }
}

// kpn.rs:278-282  T -> Vec<T> of length l
template <class T>
void shaper(Receiver<T> u, Sender<std::vector<T>> v, size_t l)
{
    for (;;) {
        std::vector<T> x;
        x.reserve(l);
        for (size_t k = 0; k < l; ++k) x.push_back(u.recv());
        v.send(std::move(x));
    }
}

// kpn.rs:295-299  eat() over a stream of bit vectors
inline void binconv(Receiver<std::vector<size_t>> u, Sender<std::vector<size_t>> v, std::vector<size_t> l)
{
    for (;;) v.send(eat(u.recv(), l));
}

// the two pulse-pair matchers of ratpak.rs:88-97, as looper bodies
using Run = std::pair<size_t, float>;
inline bool in_rng(float d, float lo, float hi) { return d >= lo && d <= hi; }

inline void matcher_a(Iter<Run> &a, Sender<std::optional<size_t>> &b)          // ratpak.rs:88-92
{
    while (auto x = a.next()) {
        std::optional<size_t> r;
        if (x->first == 1 && in_rng(x->second, 2e-4f, 6e-4f)) {
            auto y = a.next();
            if (!y) throw PortClosed();                                        // a.next().unwrap()
            if (y->first == 0 && in_rng(y->second, 1.5e-3f, 2.5e-3f)) r = 0;
            else if (y->first == 0 && in_rng(y->second, 3.5e-3f, 4.5e-3f)) r = 1;
        }
        b.send(r);
    }
}

inline void matcher_b(Iter<Run> &a, Sender<std::optional<size_t>> &b)          // ratpak.rs:93-97
{
    auto hit = [](float d) { return in_rng(d, 125e-6f, 250e-6f) || in_rng(d, 500e-6f, 650e-6f); };
    while (auto x = a.next()) {
        std::optional<size_t> r;
        if (x->first == 1 && hit(x->second)) {
            auto y = a.next();
            if (!y) throw PortClosed();
            if (y->first == 0 && hit(y->second)) r = (x->second > y->second) ? 1 : 0;
        }
        b.send(r);
    }
}

}  // namespace kpn
