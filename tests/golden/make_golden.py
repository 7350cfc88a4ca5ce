"""Generate tests/golden/kissfft_golden.npz from the UNMODIFIED vendored reference (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures let the GPU box and any machine without /root/reference check both the restated oracle and
the CUDA path against outputs of the reference itself.  Inputs are regenerated from the seeds stored in
the file, outputs are what kiss_fft()/kiss_fastfir() of the reference build returned.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

FFT_SIZES = [2, 4, 8, 16, 32, 64, 128, 240, 256, 512, 1000, 1024, 2048, 4096, 8192]
# sizes that are not powers of two: kissfft's radix 3/5 and generic (odd prime) butterflies, kiss_fft.c:92-235
MIXED_SIZES = [3, 5, 6, 7, 15, 30, 45, 97, 100, 243, 625, 1001, 1536, 3125, 6000, 7919]
# real-input transform pair (tools/kiss_fftr.c), even sizes
RFFT_SIZES = [4, 8, 30, 256, 1000, 1024, 4096]
SEED = 20261017


def fft_input(n: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed + n)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def rfft_input(n: int, seed: int = SEED) -> np.ndarray:
    return np.random.default_rng(seed + 7 * n + 1).standard_normal(n).astype(np.float32)


def psdpng_input(stereo: bool, seed: int = SEED) -> np.ndarray:
    """16-bit PCM: two tones + noise + a DC offset, 7 frames of 256 (the last row stays incomplete at navg 3)."""
    rng = np.random.default_rng(seed + (99 if stereo else 98))
    n = 7 * 256 + 100
    t = np.arange(n)
    x = 6000 * np.sin(2 * np.pi * 0.05 * t) + 2500 * np.sin(2 * np.pi * 0.21 * t) + 300 * rng.standard_normal(n) + 1234
    if stereo:
        y = 4000 * np.sin(2 * np.pi * 0.11 * t) + 300 * rng.standard_normal(n) - 500
        return np.stack([x, y], axis=1).round().astype(np.int16).reshape(-1)
    return x.round().astype(np.int16)


def fastfir_case(seed: int = SEED):
    rng = np.random.default_rng(seed)
    h = (rng.standard_normal(300) + 1j * rng.standard_normal(300)).astype(np.complex64) / 16
    x = (rng.standard_normal(6000) + 1j * rng.standard_normal(6000)).astype(np.complex64)
    return h, x


def main() -> None:
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {"seed": np.int64(SEED), "fft_sizes": np.array(FFT_SIZES)}
    for n in FFT_SIZES:
        x = fft_input(n)
        out[f"fwd_{n}"] = oracle.ref_kissfft(x, False)
        out[f"inv_{n}"] = oracle.ref_kissfft(x, True)
    out["mixed_sizes"] = np.array(MIXED_SIZES)
    for n in MIXED_SIZES:
        x = fft_input(n)
        out[f"fwd_{n}"] = oracle.ref_kissfft(x, False)
        out[f"inv_{n}"] = oracle.ref_kissfft(x, True)
    out["rfft_sizes"] = np.array(RFFT_SIZES)
    for n in RFFT_SIZES:
        t = rfft_input(n)
        F = oracle.ref_fftr(t)
        out[f"rfft_{n}"] = F
        out[f"irfft_{n}"] = oracle.ref_fftri(F)
    # tools/psdpng.c rows (restated loop around the vendored kiss_fftr): mono, stereo, with and without -a
    for stereo in (False, True):
        for dc in (False, True):
            out[f"psdpng_{int(stereo)}{int(dc)}"] = oracle.psdpng_rows(psdpng_input(stereo), 256, 3, dc, stereo,
                                                                        fftr=oracle.ref_fftr)
    h, x = fastfir_case()
    out["fastfir_noflush"] = oracle.ref_fastfir(h, x, 0, False)     # nfft auto = 1024
    out["fastfir_flush"] = oracle.ref_fastfir(h, x, 0, True)
    # the one hard-coded golden vector of the reference tree: test/fft.py:95-98 (tolerance 1e-5, :104)
    out["fftpy_tvec"] = np.array([0.309655, 0.815653, 0.768570, 0.591841, 0.404767, 0.637617, 0.007803, 0.012665],
                                 dtype=np.float32)
    out["fftpy_Ftvec"] = (np.array([3.548571, -0.378761, -0.061950, 0.188537, -0.566981, 0.188537, -0.061950, -0.378761])
                          + 1j * np.array([0.0, -1.296198, -0.848764, 0.225337, 0.0, -0.225337, 0.848764, 1.296198]))
    path = os.path.join(os.path.dirname(__file__), "kissfft_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
