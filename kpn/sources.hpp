// sources.hpp -- the wire formats either side of the hot path, host side only:
//
//   wavio::wav_source_f32 / wav_source_complex_f32   src/wavio/src/wavio.rs:12-46   (libsndfile there;
//       a self-contained RIFF/WAVE reader here: PCM u8/s16/s24/s32 and IEEE f32, scaled the way
//       libsndfile's sf_read_float scales them -- u8 (x-128)/128, s16 x/32768, s24 x/8388608, s32 x/2^31)
//   raw rtl_sdr captures (.iq, interleaved u8 I,Q)   the bytes rtlsdr::read_async delivers, block_size
//       samples per message (src/bitfount/src/bitfount.rs:16-34, src/rtlsdr/src/rtlsdr.rs:112-126)
//   oblw run/bit packing (TX direction)              src/oblw/src/oblw.rs:11-47
//
// The *_chunks variants hand whole Vec chunks to the batched GPU blocks of gpu_blocks.hpp instead of one
// message per sample.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include "kpn.hpp"

namespace kpn {

using cf32 = std::complex<float>;

struct WavInfo {
    uint32_t samplerate = 0, channels = 0, bits = 0, format = 0;   // format 1 = PCM, 3 = IEEE float
    uint64_t frames = 0;
    long data_off = 0;
};

class WavReader {
    FILE *f_ = nullptr;
    WavInfo info_;
    uint64_t left_ = 0;   // sample values (not frames) still unread
    static uint32_t u32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
    static uint16_t u16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
public:
    explicit WavReader(const std::string &fname)
    {
        f_ = std::fopen(fname.c_str(), "rb");
        // SndFile::new(fname, Read).unwrap()  wavio.rs:13,31
        if (!f_) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: cannot open " + fname);
        unsigned char h[12];
        if (std::fread(h, 1, 12, f_) != 12 || std::memcmp(h, "RIFF", 4) || std::memcmp(h + 8, "WAVE", 4))
            throw std::runtime_error("wav: not a RIFF/WAVE file: " + fname);
        bool have_fmt = false;
        for (;;) {
            unsigned char ch[8];
            if (std::fread(ch, 1, 8, f_) != 8) throw std::runtime_error("wav: no data chunk in " + fname);
            const uint32_t len = u32(ch + 4);
            if (!std::memcmp(ch, "fmt ", 4)) {
                unsigned char fm[40] = {0};
                const uint32_t take = len < 40 ? len : 40;
                if (std::fread(fm, 1, take, f_) != take) throw std::runtime_error("wav: short fmt chunk");
                info_.format = u16(fm); info_.channels = u16(fm + 2); info_.samplerate = u32(fm + 4); info_.bits = u16(fm + 14);
                if (info_.format == 0xfffe && take >= 26) info_.format = u16(fm + 24);     // WAVE_FORMAT_EXTENSIBLE
                std::fseek(f_, (long)(len - take) + (long)(len & 1), SEEK_CUR);
                have_fmt = true;
            } else if (!std::memcmp(ch, "data", 4)) {
                if (!have_fmt) throw std::runtime_error("wav: data before fmt");
                info_.data_off = std::ftell(f_);
                const uint32_t bps = info_.bits / 8;
                if (!(info_.format == 1 && (bps >= 1 && bps <= 4)) && !(info_.format == 3 && bps == 4))
                    throw std::runtime_error("wav: unsupported sample format");
                if (info_.channels == 0) throw std::runtime_error("wav: zero channels");
                left_ = len / bps;
                info_.frames = left_ / info_.channels;
                break;
            } else {
                std::fseek(f_, (long)len + (long)(len & 1), SEEK_CUR);
            }
        }
    }
    WavReader(const WavReader &) = delete;
    ~WavReader() { if (f_) std::fclose(f_); }
    const WavInfo &info() const { return info_; }
    // sf_read_float: up to n sample values (all channels interleaved) as f32; returns how many were read
    size_t read_f32(float *dst, size_t n)
    {
        if (n > left_) n = (size_t)left_;
        const uint32_t bps = info_.bits / 8;
        std::vector<unsigned char> raw(n * bps);
        const size_t got = std::fread(raw.data(), bps, n, f_);
        for (size_t i = 0; i < got; ++i) {
            const unsigned char *p = raw.data() + i * bps;
            if (info_.format == 3) { std::memcpy(dst + i, p, 4); continue; }
            switch (bps) {
                case 1: dst[i] = ((float)p[0] - 128.0f) / 128.0f; break;
                case 2: dst[i] = (float)(int16_t)u16(p) / 32768.0f; break;
                case 3: dst[i] = (float)(((int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24)) >> 8) / 8388608.0f; break;
                default: dst[i] = (float)((double)(int32_t)u32(p) / 2147483648.0); break;
            }
        }
        left_ -= got;
        return got;
    }
};

inline void wav_assert_eq(uint32_t a, uint32_t b, const char *what)
{
    // assert_eq!(info.samplerate as u32, s_rate) / assert_eq!(info.channels as u32, N)   wavio.rs:15-16,33-34
    if (a != b) throw std::logic_error(std::string("assertion failed: `(left == right)` ") + what + ": " +
                                       std::to_string(a) + " != " + std::to_string(b));
}

// wavio.rs:12-28.  One f32 per message.  The reference reads (frames/2)/1024 buffers of 1024 values --
// half the file for a mono WAV -- and then parks forever on a private channel (:25-27) so that its port
// stays open; `park` reproduces that, the default returns (dropping the port tears the graph down).
inline void wav_source_f32(Sender<float> u, const std::string &fname, uint32_t s_rate, bool park = false)
{
    WavReader w(fname);
    wav_assert_eq(w.info().samplerate, s_rate, "samplerate");
    wav_assert_eq(w.info().channels, 1, "channels");
    std::vector<float> x(1024, 0.0f);
    for (uint64_t it = 0; it < (w.info().frames / 2) / 1024; ++it) {
        w.read_f32(x.data(), 1024);
        for (float z : x) u.send(z);
    }
    if (park) { auto [c, p] = channel<int>(); p.recv(); }
}

// wavio.rs:30-46.  1024 floats = 512 I/Q frames per read, one Complex<f32> per message; the loop bound
// (frames/2)/1024 consumes a quarter of a stereo file (reference quirk kept).
inline void wav_source_complex_f32(Sender<cf32> u, const std::string &fname, uint32_t s_rate, bool park = false)
{
    WavReader w(fname);
    wav_assert_eq(w.info().samplerate, s_rate, "samplerate");
    wav_assert_eq(w.info().channels, 2, "channels");
    std::vector<float> x(1024, 0.0f);
    for (uint64_t it = 0; it < (w.info().frames / 2) / 1024; ++it) {
        w.read_f32(x.data(), 1024);
        for (size_t k = 0; k + 1 < x.size(); k += 2) u.send(cf32(x[k], x[k + 1]));
    }
    if (park) { auto [c, p] = channel<int>(); p.recv(); }
}

// the same file as Vec chunks for the batched GPU blocks: the WHOLE file (no quarter-file quirk), `chunk`
// frames per message, last message shorter
inline void wav_source_complex_chunks(Sender<std::vector<cf32>> u, const std::string &fname, uint32_t s_rate, size_t chunk)
{
    WavReader w(fname);
    wav_assert_eq(w.info().samplerate, s_rate, "samplerate");
    wav_assert_eq(w.info().channels, 2, "channels");
    std::vector<float> x(2 * chunk);
    for (;;) {
        const size_t got = w.read_f32(x.data(), 2 * chunk) / 2;
        if (got == 0) break;
        std::vector<cf32> out(got);
        for (size_t k = 0; k < got; ++k) out[k] = cf32(x[2 * k], x[2 * k + 1]);
        u.send(std::move(out));
    }
}

// raw rtl_sdr capture replay: block_size samples = 2*block_size bytes per message, exactly what
// read_async(block_size) hands to data_to_samples (bitfount.rs:17,24-28).  A trailing partial block is
// dropped (librtlsdr only ever delivers whole buffers).
inline void iq_file_source_u8(Sender<std::vector<uint8_t>> u, const std::string &fname, size_t block_size = 512)
{
    FILE *f = std::fopen(fname.c_str(), "rb");
    if (!f) throw std::runtime_error("iq_file_source_u8: cannot open " + fname);
    std::vector<uint8_t> buf(2 * block_size);
    try {
        while (std::fread(buf.data(), 1, buf.size(), f) == buf.size()) u.send(buf);
    } catch (...) { std::fclose(f); throw; }
    std::fclose(f);
}

// ---- oblw.rs:11-47, transmit direction ------------------------------------------------------------------
struct OblwRun { size_t v, ct; };

inline std::vector<size_t> oblw_rld(const std::vector<OblwRun> &input)                 // oblw.rs:17-25
{
    std::vector<size_t> out;
    for (const auto &i : input) for (size_t a = 0; a < i.ct; ++a) out.push_back(i.v);
    return out;
}
inline std::vector<bool> oblw_v2b(const std::vector<size_t> &usizes)                   // :31-34  x == 1
{
    std::vector<bool> y(usizes.size());
    for (size_t k = 0; k < usizes.size(); ++k) y[k] = usizes[k] == 1;
    return y;
}
inline std::vector<bool> oblw_B2b(const std::vector<uint8_t> &bytes)                   // :27-29  Bitv::from_bytes, MSB first
{
    std::vector<bool> b(bytes.size() * 8);
    for (size_t k = 0; k < b.size(); ++k) b[k] = (bytes[k / 8] >> (7 - k % 8)) & 1;
    return b;
}
inline std::vector<uint8_t> oblw_b2B(const std::vector<bool> &bits)                    // :36-38  Bitv::to_bytes, zero padded
{
    std::vector<uint8_t> out((bits.size() + 7) / 8, 0);
    for (size_t k = 0; k < bits.size(); ++k) if (bits[k]) out[k / 8] |= (uint8_t)(0x80u >> (k % 8));
    return out;
}
inline std::vector<bool> oblw_r2b(const std::vector<OblwRun> &runs) { return oblw_v2b(oblw_rld(runs)); }   // :40-42
inline void oblw_assemble_packet(uint8_t *y, const uint8_t *x, size_t n, bool norm)    // :44-49
{
    for (size_t i = 0; i < n; ++i) y[i] = norm ? x[i] : (uint8_t)(x[i] ^ 255u);
}

}  // namespace kpn
