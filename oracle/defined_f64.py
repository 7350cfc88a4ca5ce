"""oracle/defined_f64.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

f64 numpy definitions of the stages BASELINE.json's north-star names but the reference does not
contain: periodic Hann window, |X|^2 frame averaging, the quadrature FM
discriminator and the rational polyphase resampler.

PARITY UNPINNED for the resampler: the reference's arithmetic is libsamplerate
(src/samplerate/src/samplerate.rs:32-42, `src_new(1, 1)` = SRC_SINC_MEDIUM_QUALITY), an un-vendored,
un-pinned system library that is not installed here and has no test in the reference.  The definition
below (Kaiser-windowed sinc, streaming-causal polyphase) is OURS; the CUDA path is held to >= 100 dB SNR
against it and no libsamplerate bit/near parity is claimed.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np


# ---- window + |X|^2 averaging (nearest reference: tools/psdpng.c:157-178, which has no window) -------
def hann_periodic(n: int) -> np.ndarray:
    """w[k] = 0.5 - 0.5 cos(2 pi k / n), evaluated in f64 and rounded once to f32."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def psd_rows(x: np.ndarray, nfft: int, k_avg: int, window: np.ndarray | None) -> np.ndarray:
    """rows[r, b] = (1/K) sum_{f<K} |FFT(w * x_frame(rK+f))[b]|^2, all in f64.
    Trailing samples that do not fill a row of K frames are dropped."""
    x = np.asarray(x, dtype=np.complex128)
    nframes = x.size // nfft
    nrows = nframes // k_avg
    fr = x[: nrows * k_avg * nfft].reshape(nrows, k_avg, nfft)
    if window is not None:
        fr = fr * np.asarray(window, dtype=np.float64)[None, None, :]
    spec = np.fft.fft(fr, axis=-1)
    return (spec.real ** 2 + spec.imag ** 2).mean(axis=1)


# ---- FM discriminator (absent from the reference; north-star defined) ------------------------------
def fm_discriminator(x: np.ndarray, prev: complex = 0j) -> np.ndarray:
    """d[n] = atan2(Im z, Re z), z = x[n] * conj(x[n-1]); x[-1] = prev (0 at stream start, so d[0] = 0)."""
    x = np.asarray(x, dtype=np.complex128)
    xm1 = np.concatenate([[np.complex128(prev)], x[:-1]])
    z = x * np.conj(xm1)
    return np.arctan2(z.imag, z.real)


# ---- rational polyphase resampler (ours; libsamplerate parity unpinned) ------------------------------
RESAMPLER_ZERO_CROSSINGS = 32     # one-sided sinc zero crossings at the narrower of the two rates
RESAMPLER_KAISER_BETA = 12.0      # ~ -120 dB side lobes
RESAMPLER_BANDWIDTH = 0.9         # -6 dB point as a fraction of the narrower Nyquist (libsamplerate class)
RESAMPLER_MAX_DEN = 4096


def resampler_ratio(ratio: float) -> tuple[int, int]:
    """ratio = L/M in lowest terms; the ratio must be representable with L, M <= 4096."""
    fr = Fraction(ratio).limit_denominator(RESAMPLER_MAX_DEN)
    if fr.numerator < 1 or fr.numerator > RESAMPLER_MAX_DEN or abs(float(fr) - ratio) > 1e-12 * ratio:
        raise ValueError(f"ratio {ratio} is not L/M with L, M <= {RESAMPLER_MAX_DEN}")
    return fr.numerator, fr.denominator


def resampler_taps(L: int, M: int) -> np.ndarray:
    """Prototype low-pass at the L-times-upsampled rate, f64.
    ntaps = 2*Z*max(L, M) + 1, h[i] = sinc(2 fc (i - c)) * kaiser(beta), fc = 0.5*BW/max(L, M),
    normalised so that sum(h) = L (unity DC gain through upsample-by-L)."""
    q = max(L, M)
    ntaps = 2 * RESAMPLER_ZERO_CROSSINGS * q + 1
    c = (ntaps - 1) / 2.0
    i = np.arange(ntaps, dtype=np.float64)
    fc = 0.5 * RESAMPLER_BANDWIDTH / q
    h = np.sinc(2.0 * fc * (i - c)) * np.kaiser(ntaps, RESAMPLER_KAISER_BETA)
    return h * (L / h.sum())


def resample(x: np.ndarray, ratio: float, taps: np.ndarray | None = None) -> np.ndarray:
    """Streaming-causal polyphase resampling of a stream that is preceded by silence:
        y[m] = sum_j h[(m M mod L) + j L] * x[floor(m M / L) - j],   x[<0] = 0,
    for every m whose newest input floor(mM/L) has arrived: m in [0, floor((N L - 1)/M)].
    `taps` lets a test feed the exact (e.g. f32-rounded) coefficients the device used."""
    L, M = resampler_ratio(ratio)
    h = resampler_taps(L, M) if taps is None else np.asarray(taps, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n == 0:
        return np.empty(0)
    nout = (n * L - 1) // M + 1
    tpp = -(-h.size // L)                              # taps per phase
    hp = np.zeros(tpp * L)
    hp[: h.size] = h
    hp = hp.reshape(tpp, L)                            # hp[j, phase] = h[phase + j L]
    xp = np.concatenate([np.zeros(tpp - 1), x])        # xp[i + tpp - 1] = x[i]
    m = np.arange(nout, dtype=np.int64)
    base = (m * M) // L
    phase = (m * M) % L
    y = np.zeros(nout)
    for j in range(tpp):
        y += hp[j, phase] * xp[base - j + tpp - 1]
    return y


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.complex128 if np.iscomplexobj(ref) else np.float64)
    err = np.asarray(got) - ref
    p_sig = float(np.sum(np.abs(ref) ** 2))
    p_err = float(np.sum(np.abs(err) ** 2))
    if p_err == 0.0:
        return float("inf")
    return 10.0 * np.log10(p_sig / p_err)
