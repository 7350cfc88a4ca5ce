// CPU-only checks of the kpn block/port contract (kpn.hpp) against the semantics of src/kpn/src/kpn.rs.
#include <cassert>
#include <chrono>
#include <cstdio>
#include "kpn.hpp"
#include "sources.hpp"
#include "dsputils.hpp"
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
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
    // wavio sources (wavio.rs:12-46) over a WAV written here: s16 stereo, 8192 frames
    {
        const std::string path = std::string(std::getenv("TMPDIR") ? std::getenv("TMPDIR") : "/tmp") + "/kpn_test_iq.wav";
        const uint32_t frames = 8192, rate = 2400000, ch = 2, bits = 16;
        std::vector<int16_t> pcm(frames * ch);
        for (uint32_t k = 0; k < frames * ch; ++k) pcm[k] = (int16_t)((int)(k * 37u % 65536u) - 32768);
        {
            FILE *f = std::fopen(path.c_str(), "wb");
            CHECK(f != nullptr);
            const uint32_t data_bytes = frames * ch * bits / 8, riff = 36 + data_bytes, fmt_len = 16, byte_rate = rate * ch * bits / 8;
            const uint16_t fmt = 1, chs = (uint16_t)ch, align = (uint16_t)(ch * bits / 8), bps = (uint16_t)bits;
            std::fwrite("RIFF", 1, 4, f); std::fwrite(&riff, 4, 1, f); std::fwrite("WAVE", 1, 4, f);
            std::fwrite("fmt ", 1, 4, f); std::fwrite(&fmt_len, 4, 1, f); std::fwrite(&fmt, 2, 1, f); std::fwrite(&chs, 2, 1, f);
            std::fwrite(&rate, 4, 1, f); std::fwrite(&byte_rate, 4, 1, f); std::fwrite(&align, 2, 1, f); std::fwrite(&bps, 2, 1, f);
            std::fwrite("data", 1, 4, f); std::fwrite(&data_bytes, 4, 1, f); std::fwrite(pcm.data(), 2, pcm.size(), f);
            std::fclose(f);
        }
        auto [tx, rx] = channel<cf32>();
        std::thread t = spawn([s = std::move(tx), path]() mutable { wav_source_complex_f32(std::move(s), path, 2400000); });
        t.join();
        // (frames/2)/1024 = 4 reads of 1024 floats = 2048 complex samples: a quarter of the file (reference quirk)
        size_t n = 0;
        while (auto z = rx.try_recv()) {
            CHECK(z->real() == (float)pcm[2 * n] / 32768.0f && z->imag() == (float)pcm[2 * n + 1] / 32768.0f);
            ++n;
        }
        CHECK(n == 2048);
        // the chunked variant delivers the whole file
        auto [tc, rc] = channel<std::vector<cf32>>();
        std::thread t2 = spawn([s = std::move(tc), path]() mutable { wav_source_complex_chunks(std::move(s), path, 2400000, 3000); });
        t2.join();
        size_t total = 0, msgs = 0;
        while (auto v = rc.try_recv()) { total += v->size(); ++msgs; CHECK((*v)[0].real() == (float)pcm[2 * (total - v->size())] / 32768.0f); }
        CHECK(total == frames && msgs == 3);
        // a rate mismatch is the reference's assert_eq! panic: the block dies, the port closes
        auto [tb, rb] = channel<cf32>();
        std::thread t3 = spawn([s = std::move(tb), path]() mutable { wav_source_complex_f32(std::move(s), path, 48000); });
        t3.join();
        bool closed = false;
        try { rb.recv(); } catch (const PortClosed &) { closed = true; }
        CHECK(closed);
        // raw .iq replay: 512-sample blocks = 1024 bytes, trailing partial block dropped
        const std::string iqp = path + ".iq";
        { FILE *f = std::fopen(iqp.c_str(), "wb"); std::vector<uint8_t> b(1024 * 3 + 100); for (size_t k = 0; k < b.size(); ++k) b[k] = (uint8_t)(k * 7); std::fwrite(b.data(), 1, b.size(), f); std::fclose(f); }
        auto [ti, ri] = channel<std::vector<uint8_t>>();
        std::thread t4 = spawn([s = std::move(ti), iqp]() mutable { iq_file_source_u8(std::move(s), iqp); });
        t4.join();
        size_t blocks = 0;
        while (auto b = ri.try_recv()) { CHECK(b->size() == 1024 && (*b)[5] == (uint8_t)((blocks * 1024 + 5) * 7)); ++blocks; }
        CHECK(blocks == 3);
        std::remove(path.c_str()); std::remove(iqp.c_str());
    }
    // oblw.rs:17-49 run/bit/byte packing
    {
        std::vector<OblwRun> runs = {{1, 3}, {0, 2}, {1, 1}, {0, 4}};
        CHECK((oblw_rld(runs) == std::vector<size_t>{1, 1, 1, 0, 0, 1, 0, 0, 0, 0}));
        auto bytes = oblw_b2B(oblw_r2b(runs));                            // 1110 0100 | 00(pad)
        CHECK(bytes.size() == 2 && bytes[0] == 0xE4 && bytes[1] == 0x00);
        CHECK(oblw_B2b({0xE4})[0] && !oblw_B2b({0xE4})[3] && oblw_B2b({0xE4})[5]);
        uint8_t y[2];
        oblw_assemble_packet(y, bytes.data(), 2, false);
        CHECK(y[0] == 0x1B && y[1] == 0xFF);
        oblw_assemble_packet(y, bytes.data(), 2, true);
        CHECK(y[0] == 0xE4);
    }
    // dsputils tap designers (dsputils.rs:38-94): corrected window by default, the reference's NaN on request
    {
        using namespace dsputils;
        auto gain = [](const std::vector<float> &h, double f) {
            double re = 0, im = 0;
            for (size_t n = 0; n < h.size(); ++n) { re += h[n] * std::cos(2 * M_PI * f * n); im -= h[n] * std::sin(2 * M_PI * f * n); }
            return std::sqrt(re * re + im * im);
        };
        const auto lp = lpf(64, 0.1f), hp = hpf(64, 0.1f), bs = bsf(64, 0.05f, 0.2f), bp = bpf(64, 0.05f, 0.2f);
        CHECK(window(64).size() == 65 && lp.size() == 64 && hp.size() == 64 && bs.size() == 64 && bp.size() == 64);
        for (float v : lp) CHECK(std::isfinite(v));
        CHECK(std::fabs(gain(lp, 0.0) - 1.0) < 2e-3 && gain(lp, 0.3) < 1e-3);          // low-pass
        CHECK(gain(hp, 0.0) < 2e-3 && std::fabs(gain(hp, 0.3) - 1.0) < 2e-3);          // high-pass
        CHECK(hp[31] == -lp[31] + 1.0f && hp[32] == -lp[32]);                           // the 1.0 sits at m/2 - 1 (:77)
        for (size_t i = 0; i < 64; ++i) CHECK(bp[i] == -bs[i]);                         // bpf = -bsf (:91-94)
        CHECK(std::fabs(gain(bs, 0.0) - 1.0) < 5e-3 && std::fabs(gain(bs, 0.35) - 1.0) < 5e-3 && gain(bs, 0.1) < 0.7);
        bool nan = false;
        for (float v : lpf(64, 0.1f, true)) nan = nan || std::isnan(v);
        CHECK(nan);                                                                      // the reference's window bug (:49)
        bool threw = false;
        try { sinc(8, 0.5f); } catch (const std::logic_error &) { threw = true; }
        CHECK(threw);                                                                    // assert!(fc < 0.5) :55
    }
    std::printf("kpn cpu OK\n");
    return 0;
}
