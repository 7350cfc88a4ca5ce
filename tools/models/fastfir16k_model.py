#!/usr/bin/env python
"""Numpy model of the data flow planned for an nfft = 16384 overlap-save kernel.

    16384 = 16 x 1024,  n = 1024 n1 + n2,  k = k1 + 16 k2
    P1   for every column n2: DFT16 over n1 (inputs straight from global memory, stride 1024), times W_16384^(n2 k1),
         stored to shared row k1 at padded position n2 + n2/32 (pitch 33 per 32 entries)
    P2   warp k1 owns row k1: 1024-point transform as 32 x 32 (lane n2' holds y[lane + 32 e]; DFT32 over e, twiddle
         W_1024^(lane r), exchange through the row buffer itself as a [32][33] tile, DFT32 over the lanes) ->
         X[k1 + 16 (lane + 32 e)] in the same "lane + 32 e" layout
    H    multiply by Hp[k1][k2] = H[k1 + 16 k2]     (H = FFT(rotated taps) / nfft, kiss_fastfir.c:148-169)
    P2'  the same warp transform with conjugated twiddles, back into the row
    P1'  for every column n2: times conj(W_16384^(n2 k1)), IDFT16 over k1, outputs 1024 n1 + n2 < ngood to global memory

The model follows the per-thread / per-lane index arithmetic of that plan (including the padded shared-memory
addresses and the factorised P1 twiddle W^(n2 k1) = W^(32 a k1) W^(b k1), n2 = 32 a + b) and is checked against a
direct overlap-save block by tests/test_cpu_oracle.py.  It is a design aid, not a product path."""
import numpy as np

N, N1, N2 = 16384, 16, 1024
PITCH = 33 * 32                    # padded row length (complex entries)


def pad(n2):
    return n2 + (n2 >> 5)


def warp_fft1024(row, inverse):
    """row: 1056 padded entries holding y[n2] at pad(n2).  Returns the transform in registers v[lane][e] =
    Y[lane + 32 e], using the row buffer as the exchange tile exactly like WarpFFT1024 (fft_core.cuh)."""
    sgn = 1.0 if inverse else -1.0
    lane = np.arange(32)
    v = np.empty((32, 32), dtype=np.complex128)                       # v[lane][e]
    for e in range(32):
        v[:, e] = row[lane + 33 * e]                                  # pad(lane + 32 e) = lane + 33 e
    # pass 1: DFT32 over e in registers -> A[n2' = lane][k1' = r]
    w32 = np.exp(sgn * 2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    a = v @ w32                                                       # a[lane][r] = sum_e v[lane][e] W32^(e r)
    a *= np.exp(sgn * 2j * np.pi * np.outer(lane, np.arange(32)) / 1024)     # W_1024^(lane r)
    # ONE exchange through the row buffer as a [32][33] tile: write tile[lane][r], read tile[e][lane]
    tile = np.zeros(PITCH, dtype=np.complex128)
    for ln in range(32):
        tile[33 * ln + np.arange(32)] = a[ln]
    b = np.empty((32, 32), dtype=np.complex128)
    for e in range(32):
        b[:, e] = tile[33 * e + lane]                                 # b[lane = k1'][e = n2']
    # pass 2: DFT32 over n2' -> Y[k1' + 32 k2'] with lane = k1', register = k2'
    return b @ w32


def block(x, Hp):
    """One overlap-save block: x (16384 inputs) -> 16384 circular-convolution outputs (caller keeps ngood)."""
    sm = np.zeros((N1, PITCH), dtype=np.complex128)
    w16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    # factorised P1 twiddles: n2 = 32 a + b
    t1 = np.exp(-2j * np.pi * 32 * np.outer(np.arange(16), np.arange(32)) / N)    # [k1][a] = W^(32 a k1)
    t2 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(32)) / N)         # [k1][b] = W^(b k1)
    for t in range(512):                                              # thread t owns columns t and t + 512
        for n2 in (t, t + 512):
            a_, b_ = n2 >> 5, n2 & 31
            v = x[n2 + 1024 * np.arange(16)]
            V = w16 @ v                                               # natural order k1
            sm[:, pad(n2)] = V * t1[:, a_] * t2[:, b_]
    out_rows = np.zeros_like(sm)
    for k1 in range(16):                                              # warp k1
        X = warp_fft1024(sm[k1], inverse=False)                       # X[lane][e] = X[k1 + 16 (lane + 32 e)]
        X = X * Hp[k1].reshape(32, 32).T                              # Hp[k1][lane + 32 e]
        # inverse warp transform needs its input in the row ("lane + 32 e" at padded positions)
        row = np.zeros(PITCH, dtype=np.complex128)
        for e in range(32):
            row[np.arange(32) + 33 * e] = X[:, e]
        Y = warp_fft1024(row, inverse=True)
        for e in range(32):
            out_rows[k1, np.arange(32) + 33 * e] = Y[:, e]
    y = np.zeros(N, dtype=np.complex128)
    iw16 = np.conj(w16)
    for t in range(512):
        for n2 in (t, t + 512):
            a_, b_ = n2 >> 5, n2 & 31
            v = out_rows[:, pad(n2)] * np.conj(t1[:, a_] * t2[:, b_])
            y[n2 + 1024 * np.arange(16)] = iw16 @ v
    return y


def permute_H(H):
    """Hp[k1][k2] = H[k1 + 16 k2]: what warp k1 reads, coalesced, in the transform's output layout."""
    return np.ascontiguousarray(H.reshape(N2, N1).T)


def fastfir(h, x):
    """kiss_fastfir semantics (tools/kiss_fastfir.c:65-245, full blocks only): y[k] = sum_j h[j] x[k + nh - 1 - j]."""
    nh = h.size
    ngood = N - nh + 1
    rot = np.zeros(N, dtype=np.complex128)
    rot[0] = h[nh - 1]
    rot[N - nh + 1:] = h[: nh - 1]
    Hp = permute_H(np.fft.fft(rot) / N)
    out = []
    s = 0
    while s + N <= x.size:
        out.append(block(x[s:s + N], Hp)[:ngood])
        s += ngood
    return np.concatenate(out) if out else np.empty(0, dtype=np.complex128)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    h = (rng.standard_normal(4096) + 1j * rng.standard_normal(4096)) / 64
    x = rng.standard_normal(N + 12289) + 1j * rng.standard_normal(N + 12289)
    y = fastfir(h, x)
    ref = np.convolve(x, h)[h.size - 1: h.size - 1 + y.size]
    print(y.size, np.max(np.abs(y - ref)) / np.sqrt(np.mean(np.abs(ref) ** 2)))
