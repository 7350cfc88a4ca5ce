// dsputils.hpp -- host-side tap designers with the names and argument meaning of the reference's dsputils crate
// (src/dsputils/src/dsputils.rs:38-94): window, sinc, lpf, hpf, bsf, bpf.  f32 arithmetic like the reference.
//
// DOCUMENTED DEVIATION: dsputils::window evaluates cos(2 pi m / (x - 1)) -- arguments swapped, :49 -- which is NaN at
// x = 1, so every designer built on it returns NaN taps in the reference.  `faithful = true` reproduces that literally;
// the default evaluates the evidently intended Blackman-Nuttall window a0 - a1 cos(2 pi x/m) + a2 cos(4 pi x/m)
// - a3 cos(6 pi x/m) with the same coefficients (:42).  Everything else is kept as written, including two quirks that
// are not NaN bugs: hpf adds its 1.0 at index m/2 - 1 although the sinc peaks at m/2 (:77), and bpf = -bsf (:91-94) is a
// sign-flipped band-stop rather than a band-pass.  The GPU FIR (lrc_fir_create) rejects non-finite taps.
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace dsputils {

inline std::vector<float> window(std::size_t m, bool faithful = false)
{
    const float a0 = 0.3635819f, a1 = 0.4891775f, a2 = 0.1365995f, a3 = 0.0106411f;   // blackman-nuttall, :42
    const float pi = 3.14159265358979323846f, n = (float)m;
    std::vector<float> w(m + 1);                                                        // (0..m + 1), :48
    for (std::size_t x = 0; x < m + 1; ++x) {
        const float nn = (float)x;
        if (faithful)
            w[x] = a0 - a1 * std::cos(2.0f * pi * n / (nn - 1.0f)) + a2 * std::cos(4.0f * pi * n / (nn - 1.0f))
                   - a3 * (6.0f * pi * n / std::cos(nn - 1.0f));                        // :49, as written
        else
            w[x] = a0 - a1 * std::cos(2.0f * pi * nn / n) + a2 * std::cos(4.0f * pi * nn / n) - a3 * std::cos(6.0f * pi * nn / n);
    }
    return w;
}

inline std::vector<float> sinc(std::size_t m, float fc)
{
    if (!(fc < 0.5f)) throw std::logic_error("assertion failed: fc < 0.5");             // assert!(fc < 0.5), :55
    const float pi = 3.14159265358979323846f;
    std::vector<float> s(m);
    for (std::size_t x = 0; x < m; ++x) {
        const float n = (float)x - (float)m / 2.0f;
        float r = 2.0f * fc;
        if (n != 0.0f) r = std::sin(2.0f * pi * fc * n) / (pi * n);
        s[x] = r;
    }
    return s;
}

// low-pass: zip(window(m), sinc(m, fc)) -> m products, :66-71
inline std::vector<float> lpf(std::size_t m, float fc, bool faithful = false)
{
    const std::vector<float> w = window(m, faithful), s = sinc(m, fc);
    std::vector<float> h(m);
    for (std::size_t i = 0; i < m; ++i) h[i] = w[i] * s[i];
    return h;
}

// high-pass: -lpf, then += 1.0 at m/2 - 1, :74-79
inline std::vector<float> hpf(std::size_t m, float fc, bool faithful = false)
{
    std::vector<float> h = lpf(m, fc, faithful);
    for (float &x : h) x = -x;
    h.at(m / 2 - 1) += 1.0f;                                                            // get_mut(m/2-1).unwrap()
    return h;
}

// band-stop: lpf(fc1) + hpf(fc2), :82-88
inline std::vector<float> bsf(std::size_t m, float fc1, float fc2, bool faithful = false)
{
    const std::vector<float> lp = lpf(m, fc1, faithful), hp = hpf(m, fc2, faithful);
    std::vector<float> h(m);
    for (std::size_t i = 0; i < m; ++i) h[i] = lp[i] + hp[i];
    h.at(m / 2 - 1) -= 0.0f;
    return h;
}

// "bandpass": -bsf, :91-94
inline std::vector<float> bpf(std::size_t m, float fc1, float fc2, bool faithful = false)
{
    std::vector<float> h = bsf(m, fc1, fc2, faithful);
    for (float &x : h) x = -x;
    return h;
}

}  // namespace dsputils
