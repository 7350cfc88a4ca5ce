// CPU-only checks of the kpn block/port contract (kpn.hpp) against the semantics of src/kpn/src/kpn.rs.
#include <cassert>
#include <chrono>
#include <cstdio>
#include "kpn.hpp"
using namespace kpn;

#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
    // unbounded, non-blocking send; blocking recv; messages are moved
    {
        auto [tx, rx] = channel<std::vector<int>>();
        for (int i = 0; i < 100000; ++i) tx.send(std::vector<int>{i});      // never blocks
        CHECK(rx.pending() == 100000);
        CHECK(rx.recv()[0] == 0 && rx.recv()[0] == 1);
        CHECK(!channel<int>().second.try_recv().has_value());
    }
    // closed port: recv().unwrap() panics once drained and all senders are gone; send().unwrap() panics once
    // the receiver is gone
    {
        auto [tx, rx] = channel<int>();
        tx.send(7);
        tx.drop();
        CHECK(rx.recv() == 7);
        bool threw = false;
        try { rx.recv(); } catch (const PortClosed &) { threw = true; }
        CHECK(threw);
        auto [tx2, rx2] = channel<int>();
        rx2.drop();
        threw = false;
        try { tx2.send(1); } catch (const PortClosed &) { threw = true; }
        CHECK(threw);
    }
    // teardown cascades: closing the source ends every block thread downstream
    {
        auto [t0, r0] = channel<size_t>();
        auto [t1, r1] = channel<std::pair<size_t, size_t>>();
        auto [t2, r2] = channel<std::pair<size_t, float>>();
        std::thread a = spawn([r = std::move(r0), t = std::move(t1)]() mutable { rle<size_t>(std::move(r), std::move(t)); });
        std::thread b = spawn([r = std::move(r1), t = std::move(t2)]() mutable { dle<size_t>(std::move(r), std::move(t), 256000); });
        // 0 0 0 1 1 0 1 : runs (0,3) (1,2) (0,1); the final run (1,..) is never flushed (kpn.rs:17-29)
        for (size_t x : {0, 0, 0, 1, 1, 0, 1}) t0.send(x);
        auto p = r2.recv(); CHECK(p.first == 0 && p.second == 3.0f / 256000.0f);
        p = r2.recv(); CHECK(p.first == 1 && p.second == 2.0f / 256000.0f);
        p = r2.recv(); CHECK(p.first == 0 && p.second == 1.0f / 256000.0f);
        t0.drop();
        a.join(); b.join();
        bool threw = false;
        try { r2.recv(); } catch (const PortClosed &) { threw = true; }
        CHECK(threw);
    }
    // b2d / eat (kpn.rs:111-124) on the two field layouts of ratpak.rs:115,119
    {
        CHECK(b2d({1, 0, 1}) == 5);
        std::vector<size_t> bits = {0,1,0,1, 1,0,0,0,0,1,1,1, 0,1,1,0, 0,0,0,0,0,0,0,0,0,0,0,1, 1,1,1,1,1,1,1,1};
        CHECK((eat(bits, {4, 8, 4, 12, 8}) == std::vector<size_t>{5, 135, 6, 1, 255}));
        CHECK((eat(bits, {4, 8, 2, 10, 12}) == std::vector<size_t>{5, 135, 1, 512, 511}));
        bool threw = false;
        try { eat({1, 0}, {4}); } catch (const std::out_of_range &) { threw = true; }
        CHECK(threw);
    }
    // matchers + shaper_optional on hand-made runs (ratpak.rs:88-110)
    {
        auto run_graph = [](std::vector<Run> runs, bool proto_b, size_t l) {
            auto [t0, r0] = channel<Run>();
            auto [t1, r1] = channel<std::optional<size_t>>();
            auto [t2, r2] = channel<std::vector<size_t>>();
            std::thread a = spawn([r = std::move(r0), t = std::move(t1), proto_b]() mutable {
                looper<Run, std::optional<size_t>>(std::move(r), std::move(t),
                    [proto_b](Iter<Run> &it, Sender<std::optional<size_t>> &s) { proto_b ? matcher_b(it, s) : matcher_a(it, s); });
            });
            std::thread b = spawn([r = std::move(r1), t = std::move(t2), l]() mutable { shaper_optional<size_t>(std::move(r), std::move(t), l); });
            for (auto &x : runs) t0.send(x);
            t0.drop();
            a.join(); b.join();
            std::vector<std::vector<size_t>> out;
            while (auto p = r2.try_recv()) out.push_back(*p);
            return out;
        };
        // proto A, 3-bit "packets": pulse 4e-4 then gap 2e-3 (0) / 4e-3 (1); a None flushes
        std::vector<Run> a = {{0, 1e-2f}, {1, 4e-4f}, {0, 2e-3f}, {1, 4e-4f}, {0, 4e-3f}, {1, 4e-4f}, {0, 2e-3f},
                              {1, 4e-4f}, {0, 9e-3f}, {1, 3e-3f}};
        auto pa = run_graph(a, false, 3);
        CHECK(pa.size() == 1 && (pa[0] == std::vector<size_t>{0, 1, 0}));
        CHECK(run_graph(a, false, 4).empty());                          // wrong length: dropped (kpn.rs:271)
        // inclusive range ends (a...b) and the consumed-next-run rule
        auto edge = run_graph({{1, 2e-4f}, {0, 2.5e-3f}, {1, 6e-4f}, {0, 3.5e-3f}, {0, 1.0f}}, false, 2);
        CHECK(edge.size() == 1 && (edge[0] == std::vector<size_t>{0, 1}));
        // proto B: bit = (high > low)
        std::vector<Run> b = {{1, 6e-4f}, {0, 2e-4f}, {1, 2e-4f}, {0, 6e-4f}, {1, 2e-4f}, {0, 5e-2f}};
        auto pb = run_graph(b, true, 2);
        CHECK(pb.size() == 1 && (pb[0] == std::vector<size_t>{1, 0}));
    }
    // fork duplicates, looper_optional drops Nones
    {
        auto [t0, r0] = channel<int>();
        auto [ta, ra] = channel<int>();
        auto [tb, rb] = channel<int>();
        std::vector<Sender<int>> outs; outs.push_back(std::move(ta)); outs.push_back(std::move(tb));
        std::thread f = spawn([r = std::move(r0), o = std::move(outs)]() mutable { fork<int>(std::move(r), std::move(o)); });
        t0.send(3); t0.send(4); t0.drop();
        f.join();
        CHECK(ra.recv() == 3 && ra.recv() == 4 && rb.recv() == 3 && rb.recv() == 4);
    }
    std::printf("kpn cpu OK\n");
    return 0;
}
