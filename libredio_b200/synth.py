"""Seeded synthetic inputs for BASELINE.json's configs and the host-side tap designer.

Host-side helpers only (numpy); nothing here is on the device hot path.
"""
from __future__ import annotations

import numpy as np


def lpf_taps(m: int = 64, fc: float = 0.04) -> np.ndarray:
    """The low-pass dsputils::lpf(m, fc) evidently intends (src/dsputils/src/dsputils.rs:38-71):
    Blackman-Nuttall window (coefficients of :42) x sinc(2 fc) centred at m/2, f32.

    DOCUMENTED DEVIATION: the reference's `window` swaps its arguments and returns NaN at index 1
    (dsputils.rs:49), so `lpf` yields NaN taps there; this designer evaluates cos(2 pi k x / m) as
    intended.  FIR parity is on `convolve` with supplied finite taps, never on the designer."""
    a = np.array([0.3635819, 0.4891775, 0.1365995, 0.0106411], dtype=np.float64)
    x = np.arange(m, dtype=np.float64)
    w = a[0] - a[1] * np.cos(2 * np.pi * x / m) + a[2] * np.cos(4 * np.pi * x / m) - a[3] * np.cos(6 * np.pi * x / m)
    n = x - m / 2.0
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.where(n == 0.0, 2.0 * fc, np.sin(2 * np.pi * fc * n) / (np.pi * n))
    return (w * s).astype(np.float32)


def hpf_taps(m: int = 64, fc: float = 0.04) -> np.ndarray:
    """dsputils::hpf(m, fc) (dsputils.rs:74-79) on the corrected window: -lpf with 1.0 added at index m/2 - 1 --
    the reference's own position, one tap before the sinc's peak at m/2."""
    h = (-lpf_taps(m, fc)).astype(np.float32)
    h[m // 2 - 1] = np.float32(h[m // 2 - 1] + np.float32(1.0))
    return h


def bsf_taps(m: int, fc1: float, fc2: float) -> np.ndarray:
    """dsputils::bsf(m, fc1, fc2) (dsputils.rs:82-88): lpf(fc1) + hpf(fc2)."""
    return (lpf_taps(m, fc1) + hpf_taps(m, fc2)).astype(np.float32)


def bpf_taps(m: int, fc1: float, fc2: float) -> np.ndarray:
    """dsputils::bpf(m, fc1, fc2) (dsputils.rs:91-94): -bsf."""
    return (-bsf_taps(m, fc1, fc2)).astype(np.float32)


def hann_periodic(n: int) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def iq_tone_noise_u8(n: int, seed: int = 1, amp: float = 0.6, f0: float = 0.013) -> np.ndarray:
    """Config 1: u8 IQ = clip(round(127.5 + 127.5*A*(tone + noise))), interleaved I,Q (2n bytes)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    z = np.exp(2j * np.pi * f0 * t) + 0.15 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype=np.float64)
    iq[0::2] = z.real
    iq[1::2] = z.imag
    return np.clip(np.rint(127.5 + 127.5 * amp * iq / 1.6), 0, 255).astype(np.uint8)


def cf32_noise_tones(n: int, seed: int = 2) -> np.ndarray:
    """Config 2: standard-normal re/im + 3 tones, complex64."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n, dtype=np.float32) + 1j * rng.standard_normal(n, dtype=np.float32)).astype(np.complex64)
    t = np.arange(n, dtype=np.float64)
    for f, a in ((0.011, 2.0), (-0.0273, 1.0), (0.0402, 0.5)):
        x += (a * np.exp(2j * np.pi * f * t)).astype(np.complex64)
    return x


def fm_iq_u8(n: int, seed: int = 3, fs: float = 2.4e6, dev: float = 75e3) -> np.ndarray:
    """Config 3: FM-modulated 1 kHz tone + slow chirp, deviation 75 kHz, as u8 IQ."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    audio = 0.6 * np.sin(2 * np.pi * 1e3 * t) + 0.3 * np.sin(2 * np.pi * (2e3 * t + 4e5 * t * t))
    ph = 2 * np.pi * dev * np.cumsum(audio) / fs
    z = 0.8 * np.exp(1j * ph) + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty(2 * n, dtype=np.float64)
    iq[0::2] = z.real
    iq[1::2] = z.imag
    return np.clip(np.rint(127.5 + 127.5 * iq), 0, 255).astype(np.uint8)


# ---- config 4: OOK captures ---------------------------------------------------------------------------
OOK_RATE = 256000          # ratpak.rs:62
OOK_BLOCK = 512            # bitfount.rs:17


def _ook_runs_proto_a(bits, rng):
    """pulse 2e-4..6e-4 s high then gap 1.5e-3..2.5e-3 (0) or 3.5e-3..4.5e-3 (1), ratpak.rs:88-92."""
    runs = []
    for b in bits:
        hi = rng.uniform(2.6e-4, 5.4e-4)
        lo = rng.uniform(1.65e-3, 2.35e-3) if b == 0 else rng.uniform(3.65e-3, 4.35e-3)
        runs += [(1, hi), (0, lo)]
    return runs


def _ook_runs_proto_b(bits, rng):
    """short 125..250 us / long 500..650 us; bit = (high > low), ratpak.rs:93-97."""
    runs = []
    for b in bits:
        s, l = rng.uniform(150e-6, 225e-6), rng.uniform(525e-6, 625e-6)
        runs += [(1, l), (0, s)] if b else [(1, s), (0, l)]
    return runs


def ook_capture_u8(n_blocks: int, seed: int = 4, n_packets: int = 2, noise_lsb: float = 1.0, amp: float = 0.7):
    """One stream: noise floor + OOK bursts carrying proto-A (36-bit) / proto-B (24-bit) packets.
    Returns (iq u8 [n_blocks*1024], list of (proto, bits))."""
    rng = np.random.default_rng(seed)
    n = n_blocks * OOK_BLOCK
    env = np.zeros(n, dtype=np.float64)
    sent = []
    # leave ~60 quiet blocks first so the trigger threshold settles (bitfount.rs:57-65)
    pos = 60 * OOK_BLOCK + int(rng.integers(0, 4000))
    for k in range(n_packets):
        proto = int(rng.integers(0, 2))
        nb = 36 if proto == 0 else 24
        bits = rng.integers(0, 2, nb).astype(np.uint8)
        runs = _ook_runs_proto_a(bits, rng) if proto == 0 else _ook_runs_proto_b(bits, rng)
        # a closing pulse + long gap terminates the packet (the matcher emits None)
        runs += [(1, 3.0e-4 if proto == 0 else 2.0e-4), (0, 8e-3)]
        need = int(sum(d for _, d in runs) * OOK_RATE) + 60 * OOK_BLOCK
        if pos + need >= n:
            break
        p = pos
        for v, d in runs:
            m = max(1, int(round(d * OOK_RATE)))
            if v:
                env[p:p + m] = amp + 0.05 * rng.standard_normal()
            p += m
        sent.append((proto, bits))
        pos = p + 70 * OOK_BLOCK + int(rng.integers(0, 20000))
    # kpn::rle never flushes the final run (kpn.rs:17-29), so the last packet only comes out once a LATER
    # burst changes the bit value again: end the capture with a lone terminator pulse
    if pos + 64 * OOK_BLOCK < n:
        env[pos:pos + 300] = amp
    ph = rng.uniform(0, 2 * np.pi, n)
    z = env * np.exp(1j * ph)
    i = 127.0 + 127.0 * z.real + noise_lsb * rng.standard_normal(n)
    q = 127.0 + 127.0 * z.imag + noise_lsb * rng.standard_normal(n)
    iq = np.empty(2 * n, dtype=np.float64)
    iq[0::2] = i
    iq[1::2] = q
    return np.clip(np.rint(iq), 0, 255).astype(np.uint8), sent


def ook_guard_capture_u8(loud_blocks: int, seed: int = 0, before: bool = True, after: bool = True, tail_quiet: int = 80):
    """A capture that drives bitfount::trigger's OOM guard (bitfount.rs:52-54) once its constant is shrunk by the test hooks
    (oracle.set_trigger_guard_blocks / LRC_OOK_TEST_GUARD_BLOCKS): a quiet floor, optionally a short loud burst (3 blocks,
    sent 48 blocks later), then `loud_blocks` blocks that every one re-fire the trigger (the threshold is frozen while
    triggered, :62-65), optionally another short burst, and a quiet tail.  The long stretch collects loud_blocks + 48 blocks:
    with a guard of G blocks it is cut into abandoned pieces of G + 1 blocks, and when loud_blocks + 48 == G exactly the guard
    fires on the block where the counter stands at 1 -- nothing is pushed and the NEXT block sends the reset buffer [0.0]."""
    rng = np.random.default_rng(seed)

    def quiet(n):
        return np.clip(np.rint(127 + 1.5 * rng.standard_normal(n * 2 * OOK_BLOCK)), 0, 255).astype(np.uint8)

    def loud(n):
        # pulsed inside every block, so the slicer has transitions to place after the guard moved the bit positions
        x = 127 + 60 * rng.standard_normal(n * 2 * OOK_BLOCK)
        gate = (np.arange(n * 2 * OOK_BLOCK) // 128) % 2 == 0
        return np.clip(np.rint(np.where(gate, x, 127 + 1.5 * rng.standard_normal(n * 2 * OOK_BLOCK))), 0, 255).astype(np.uint8)

    parts = [quiet(100)]
    if before:
        parts += [loud(3), quiet(70)]
    parts += [loud(loud_blocks), quiet(70)]
    if after:
        parts += [loud(3), quiet(70)]
    parts += [quiet(tail_quiet)]
    return np.concatenate(parts)

