// GPU blocks wired into a small KPN graph; results checked against direct f64 evaluation in this file.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "gpu_blocks.hpp"
using namespace kpn;
using kpn_gpu::cf32;

#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
    kpn_gpu::Gpu gpu(0);
    std::mt19937 rng(1);
    std::normal_distribution<float> nd;
    // --- kissfft::fft block: 37 frames of 256, batched transparently --------------------------------
    {
        const int N = 256, F = 37;
        auto [tx, rx] = channel<std::vector<cf32>>();
        auto [ty, ry] = channel<std::vector<cf32>>();
        std::thread t = spawn([&gpu, r = std::move(rx), s = std::move(ty)]() mutable { kpn_gpu::fft(gpu, std::move(r), std::move(s), 256, 0); });
        std::vector<std::vector<cf32>> frames(F, std::vector<cf32>(N));
        for (auto &f : frames) for (auto &x : f) x = cf32(nd(rng), nd(rng));
        for (auto &f : frames) tx.send(f);
        double worst = 0, rms = 0;
        for (int k = 0; k < F; ++k) {
            std::vector<cf32> y = ry.recv();
            CHECK((int)y.size() == N);
            for (int b = 0; b < N; b += 17) {
                std::complex<double> acc = 0;
                for (int n = 0; n < N; ++n)
                    acc += std::complex<double>(frames[k][n]) * std::polar(1.0, -2.0 * M_PI * b * n / N);
                worst = std::max(worst, std::abs(acc - std::complex<double>(y[b])));
                rms += std::norm(acc);
            }
        }
        rms = std::sqrt(rms / (F * (N / 17 + 1)));
        CHECK(worst <= 1e-4 * rms);
        // a frame of the wrong length kills the block like the reference's assert (kissfft.rs:24)
        tx.send(std::vector<cf32>(100));
        t.join();
        bool closed = false;
        try { ry.recv(); } catch (const PortClosed &) { closed = true; }
        CHECK(closed);
    }
    // --- channel ring: 6 channels x 5 chunks through fir_decimate_multi --------------------------------
    {
        const size_t n_ch = 6, chunk = 4000, n_chunks = 5, ntaps = 64, decim = 10;
        std::vector<float> taps(ntaps);
        for (auto &h : taps) h = nd(rng) / 8;
        std::vector<Receiver<std::vector<cf32>>> ins; std::vector<Sender<std::vector<cf32>>> outs;
        std::vector<Sender<std::vector<cf32>>> src; std::vector<Receiver<std::vector<cf32>>> sink;
        for (size_t c = 0; c < n_ch; ++c) {
            auto [a, b] = channel<std::vector<cf32>>(); src.push_back(std::move(a)); ins.push_back(std::move(b));
            auto [d, e] = channel<std::vector<cf32>>(); outs.push_back(std::move(d)); sink.push_back(std::move(e));
        }
        std::thread t = spawn([&gpu, i = std::move(ins), o = std::move(outs), taps]() mutable {
            kpn_gpu::fir_decimate_multi(gpu, std::move(i), std::move(o), taps, decim, chunk); });
        std::vector<std::vector<cf32>> x(n_ch, std::vector<cf32>(chunk * n_chunks));
        for (auto &ch : x) for (auto &s : ch) s = cf32(nd(rng), nd(rng));
        for (size_t k = 0; k < n_chunks; ++k)
            for (size_t c = 0; c < n_ch; ++c) src[c].send(std::vector<cf32>(x[c].begin() + k * chunk, x[c].begin() + (k + 1) * chunk));
        for (auto &s : src) s.drop();
        t.join();
        const size_t n_out = (chunk * n_chunks - ntaps) / decim + 1;
        for (size_t c = 0; c < n_ch; ++c) {
            std::vector<cf32> y;
            while (auto p = sink[c].try_recv()) y.insert(y.end(), p->begin(), p->end());
            CHECK(y.size() == n_out);                                      // seam-exact: nothing lost at chunk seams
            double worst = 0, rms = 0;
            for (size_t k = 0; k < n_out; k += 7) {
                std::complex<double> acc = 0;
                for (size_t j = 0; j < ntaps; ++j) acc += std::complex<double>(x[c][k * decim + j]) * (double)taps[j];
                worst = std::max(worst, std::abs(acc - std::complex<double>(y[k])));
                rms += std::norm(acc);
            }
            CHECK(worst <= 1e-4 * std::sqrt(rms / (n_out / 7 + 1)));
        }
    }
    // --- data_to_samples -> fm_demod -> resample(0.2) pipeline on threads ----------------------------
    {
        auto [t0, r0] = channel<std::vector<uint8_t>>();
        auto [t1, r1] = channel<std::vector<cf32>>();
        auto [t2, r2] = channel<std::vector<float>>();
        auto [t3, r3] = channel<std::vector<float>>();
        std::thread a = spawn([&gpu, r = std::move(r0), s = std::move(t1)]() mutable { kpn_gpu::data_to_samples(gpu, std::move(r), std::move(s)); });
        std::thread b = spawn([&gpu, r = std::move(r1), s = std::move(t2)]() mutable { kpn_gpu::fm_demod(gpu, std::move(r), std::move(s)); });
        std::thread c = spawn([&gpu, r = std::move(r2), s = std::move(t3)]() mutable { kpn_gpu::resample(gpu, std::move(r), std::move(s), 0.2); });
        std::vector<uint8_t> iq(2 * 5000);
        for (size_t n = 0; n < 5000; ++n) {                                // constant rotation 0.1 rad/sample
            iq[2 * n] = (uint8_t)std::lround(127 + 100 * std::cos(0.1 * n));
            iq[2 * n + 1] = (uint8_t)std::lround(127 + 100 * std::sin(0.1 * n));
        }
        t0.send(iq);
        std::vector<float> y = r3.recv();
        CHECK(y.size() == 1000);
        for (size_t k = 400; k < 1000; ++k) CHECK(std::fabs(y[k] - 0.1f) < 5e-3f);   // demodulated DC = 0.1 rad/sample
        t0.drop(); a.join(); b.join(); c.join();
    }
    // --- config-3 receiver ring (lrc_fmrx), 3 channels; chunked == one big chunk, bit for bit ------------------------
    {
        const size_t n_ch = 3, n_samp = 60000, ntaps = 64, decim = 10;
        std::vector<float> taps(ntaps);
        for (size_t j = 0; j < ntaps; ++j) {                               // windowed sinc, cutoff 0.04, unity DC gain
            const double m = (double)j - 31.5, w = 0.5 - 0.5 * std::cos(2.0 * M_PI * (j + 0.5) / ntaps);
            taps[j] = (float)(w * std::sin(2.0 * M_PI * 0.04 * m) / (M_PI * m));
        }
        double sum = 0; for (float h : taps) sum += h;
        for (float &h : taps) h = (float)(h / sum);
        std::vector<std::vector<uint8_t>> iq(n_ch, std::vector<uint8_t>(2 * n_samp));
        for (size_t c = 0; c < n_ch; ++c)
            for (size_t n = 0; n < n_samp; ++n) {                          // constant rotation: 0.004 (c+1) rad/sample
                const double ph = 0.004 * (double)(c + 1) * (double)n;
                iq[c][2 * n] = (uint8_t)std::lround(127 + 100 * std::cos(ph));
                iq[c][2 * n + 1] = (uint8_t)std::lround(127 + 100 * std::sin(ph));
            }
        auto run = [&](size_t chunk_bytes) {
            std::vector<Receiver<std::vector<uint8_t>>> ins; std::vector<Sender<std::vector<float>>> outs;
            std::vector<Sender<std::vector<uint8_t>>> src; std::vector<Receiver<std::vector<float>>> sink;
            for (size_t c = 0; c < n_ch; ++c) {
                auto [a, b] = channel<std::vector<uint8_t>>(); src.push_back(std::move(a)); ins.push_back(std::move(b));
                auto [d, f] = channel<std::vector<float>>(); outs.push_back(std::move(d)); sink.push_back(std::move(f));
            }
            std::thread t = spawn([&gpu, i = std::move(ins), o = std::move(outs), taps, chunk_bytes]() mutable {
                kpn_gpu::fm_receiver_multi(gpu, std::move(i), std::move(o), taps, decim, 0.2, chunk_bytes); });
            for (size_t k = 0; k < 2 * n_samp / chunk_bytes; ++k)
                for (size_t c = 0; c < n_ch; ++c)
                    src[c].send(std::vector<uint8_t>(iq[c].begin() + k * chunk_bytes, iq[c].begin() + (k + 1) * chunk_bytes));
            for (auto &s : src) s.drop();
            t.join();
            std::vector<std::vector<float>> y(n_ch);
            for (size_t c = 0; c < n_ch; ++c)
                while (auto p = sink[c].try_recv()) y[c].insert(y[c].end(), p->begin(), p->end());
            return y;
        };
        const auto whole = run(2 * n_samp), parts = run(2 * n_samp / 6);
        const size_t n_bb = (n_samp - ntaps) / decim + 1, n_au = (n_bb - 1) / 5 + 1;
        for (size_t c = 0; c < n_ch; ++c) {
            CHECK(whole[c].size() == n_au && parts[c].size() == n_au);
            CHECK(std::memcmp(whole[c].data(), parts[c].data(), n_au * sizeof(float)) == 0);   // seam-exact through all three stages
            for (size_t k = 300; k < n_au; ++k) CHECK(std::fabs(whole[c][k] - 0.04f * (float)(c + 1)) < 5e-3f);
        }
        std::printf("kpn gpu fm_receiver_multi OK\n");
    }
    std::printf("kpn gpu OK\n");
    return 0;
}
