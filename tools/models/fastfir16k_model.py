#!/usr/bin/env python
"""Numpy model of the data flow of the nfft = 16384 overlap-save kernel (libredio_b200/csrc/k_fastfir16k.cu).

    16384 = 16 x 1024,  n = 1024 n1 + n2,  k = k1 + 16 k2
    P1   thread (lane l, warp j) owns the column pair n2 = l + 64 j and n2 + 32: DFT16 over n1 (inputs straight from
         global memory, stride 1024), times W_16384^(n2 k1) -- the second column's twiddle is W_512^k1 times the first --
         stored to shared row k1 as ONE 16-byte cell at l * 34 + 2 j (row element n2 = l + 32 e lives at l * 34 + e)
    P2   warp k1 owns row k1: 1024-point transform as 32 x 32 (lane l holds y[l + 32 e] = its own 32-entry slab, read
         as 16-byte words; DFT32 over e, twiddle W_1024^(l r), exchange through the row buffer itself -- slab writes,
         column reads -- DFT32 over the lanes) -> X[k1 + 16 (l + 32 e)] in the same layout
    H    multiply by Hq[k1][i][l][c] = H[k1 + 16 (l + 32 (2 i + c))]   (H = FFT(rotated taps) / nfft, kiss_fastfir.c:148-169)
    P2'  the same warp transform with conjugated twiddles, back into the slabs
    P1'  the column pair: times conj(W_16384^(n2 k1)), IDFT16 over k1, outputs 1024 n1 + n2 < ngood to global memory

The model follows the per-thread / per-lane index arithmetic of the kernel (including the padded shared-memory
addresses) and is checked against a direct overlap-save block by tests/test_cpu_oracle.py.  It is a design aid, not a
product path."""
import numpy as np

N, N1, N2 = 16384, 16, 1024
LP = 34                            # entries per lane slab (32 used)
RP = 32 * LP                       # padded row length (complex entries)


def pos(n2):
    return (n2 & 31) * LP + (n2 >> 5)


def warp_fft1024(v, row, inverse):
    """v[lane][e] = y[lane + 32 e] in registers.  Returns Y[lane + 32 e]; uses `row` as the exchange tile exactly
    like ff16k::warp_fft1024: slab writes tile[lane][r], column reads tile[e][lane]."""
    sgn = 1.0 if inverse else -1.0
    lane = np.arange(32)
    w32 = np.exp(sgn * 2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    a = v @ w32                                                       # a[lane][r] = sum_e v[lane][e] W32^(e r)
    a = a * np.exp(sgn * 2j * np.pi * np.outer(lane, np.arange(32)) / 1024)   # W_1024^(lane r)
    for ln in range(32):
        row[ln * LP + np.arange(32)] = a[ln]
    b = np.empty((32, 32), dtype=np.complex128)
    for e in range(32):
        b[:, e] = row[e * LP + lane]                                  # b[lane = k1'][e = n2']
    return b @ w32


def block(x, Hq):
    """One overlap-save block: x (16384 inputs) -> 16384 circular-convolution outputs (caller keeps ngood)."""
    sm = np.zeros((N1, RP), dtype=np.complex128)
    w16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    w512 = np.exp(-2j * np.pi * np.arange(16) / 512)
    for t in range(512):                                              # thread t: lane l, warp j
        l, j = t & 31, t >> 5
        ca = l + 64 * j
        twa = np.exp(-2j * np.pi * ca * np.arange(16) / N)            # tw1[t][q]
        va = w16 @ x[ca + 1024 * np.arange(16)]
        vb = w16 @ x[ca + 32 + 1024 * np.arange(16)]
        cell = l * LP + 2 * j
        assert cell == pos(ca) and cell + 1 == pos(ca + 32)
        sm[:, cell] = va * twa
        sm[:, cell + 1] = vb * w512 * twa
    for k1 in range(16):                                              # warp k1
        row = sm[k1]
        v = np.stack([row[ln * LP: ln * LP + 32] for ln in range(32)])    # v[lane][e]
        X = warp_fft1024(v, row, inverse=False)                       # X[lane][e] = X[k1 + 16 (lane + 32 e)]
        X = X * Hq[k1].transpose(1, 0, 2).reshape(32, 32)             # Hq[k1][i][lane][c] -> [lane][2 i + c]
        Y = warp_fft1024(X, row, inverse=True)
        for ln in range(32):
            row[ln * LP: ln * LP + 32] = Y[ln]
    y = np.zeros(N, dtype=np.complex128)
    iw16 = np.conj(w16)
    for t in range(512):
        l, j = t & 31, t >> 5
        ca = l + 64 * j
        twa = np.exp(-2j * np.pi * ca * np.arange(16) / N)
        cell = l * LP + 2 * j
        y[ca + 1024 * np.arange(16)] = iw16 @ (sm[:, cell] * np.conj(twa))
        y[ca + 32 + 1024 * np.arange(16)] = iw16 @ (sm[:, cell + 1] * np.conj(w512 * twa))
    return y


def permute_H(H):
    """Hq[k1][i][lane][c] = H[k1 + 16 (lane + 32 (2 i + c))]: lrc_fastfir16k_permute_H."""
    Hq = np.empty((16, 16, 32, 2), dtype=H.dtype)
    for k1 in range(16):
        for i in range(16):
            for c in range(2):
                Hq[k1, i, :, c] = H[k1 + 16 * (np.arange(32) + 32 * (2 * i + c))]
    return Hq


def fastfir(h, x):
    """kiss_fastfir semantics (tools/kiss_fastfir.c:65-245, full blocks only): y[k] = sum_j h[j] x[k + nh - 1 - j]."""
    nh = h.size
    ngood = N - nh + 1
    rot = np.zeros(N, dtype=np.complex128)
    rot[0] = h[nh - 1]
    rot[N - nh + 1:] = h[: nh - 1]
    Hq = permute_H(np.fft.fft(rot) / N)
    out = []
    s = 0
    while s + N <= x.size:
        out.append(block(x[s:s + N], Hq)[:ngood])
        s += ngood
    return np.concatenate(out) if out else np.empty(0, dtype=np.complex128)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    h = (rng.standard_normal(4096) + 1j * rng.standard_normal(4096)) / 64
    x = rng.standard_normal(N + 12289) + 1j * rng.standard_normal(N + 12289)
    y = fastfir(h, x)
    ref = np.convolve(x, h)[h.size - 1: h.size - 1 + y.size]
    print(y.size, np.max(np.abs(y - ref)) / np.sqrt(np.mean(np.abs(ref) ** 2)))
