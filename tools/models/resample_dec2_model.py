#!/usr/bin/env python
"""Numpy model of the index arithmetic of resample_dec2_kernel + lrc_resampler_process (k_fm_resample.cu), L = 1.

The GPU tests exercise a handful of shapes; this model follows the host and device index logic line by line --
the virtual row [carry (TPP-1) | chunk], s0 = m_next*M - n_total, tile pairing (tiles 2p, 2p+1 of the flattened
(channel, tile) list), the three fill paths (whole window inside the row / past the carry with a bound / generic
addr()), the per-thread window sx = tile + t*STEP with accumulator r reading sx[k + r*M], the store predicates and
the carry update -- so that tests/test_cpu_oracle.py can sweep many (M, R, NT, n_ch, chunking) shapes on the CPU
against the f64 definition (oracle/defined_f64.py).  Arithmetic is f64 here: only the indexing is under test."""
import numpy as np

ZC = 32                                    # RS_ZERO_CROSSINGS


class Dec2Model:
    def __init__(self, M, R, NT, n_ch, h):
        self.M, self.R, self.NT, self.n_ch = M, R, NT, n_ch
        self.TPP = 2 * ZC * M + 1
        assert h.size == self.TPP
        self.g = h[::-1].copy()            # taps.g[i] = h[TPP-1-i]: correlation form
        self.WIN = (R - 1) * M + self.TPP
        self.WIN2 = (self.WIN + 1) // 2 * 2
        self.STEP = R * M
        self.TILE_OUT = R * NT
        self.TILE_IN = (NT - 1) * self.STEP + self.WIN2
        self.carry = np.zeros((n_ch, self.TPP))          # d_carry: TPP floats per channel, the first TPP-1 are history
        self.n_total = 0
        self.m_next = 0
        self.paths = {"full": 0, "fast": 0, "generic": 0}

    def next_out_len(self, n_in):
        N = self.n_total + n_in
        return (N - 1) // self.M + 1 - self.m_next       # L = 1:  (N L - 1) / M + 1 - m_next

    def row_at(self, c, chunk, v):
        hist = self.TPP - 1
        if v < hist:
            return self.carry[c, v]
        k = v - hist
        return chunk[c, k] if k < chunk.shape[1] else 0.0

    def process(self, chunk):
        n_ch, n_in = chunk.shape
        M, R, NT, TPP = self.M, self.R, self.NT, self.TPP
        hist = TPP - 1
        no = self.next_out_len(n_in)
        out = np.full((n_ch, max(no, 0)), np.nan)
        if no > 0:
            s0 = self.m_next * M - self.n_total
            assert s0 >= 0
            tiles_per_ch = -(-no // self.TILE_OUT)
            n_tiles = tiles_per_ch * n_ch
            for pair in range((n_tiles + 1) // 2):
                tiles = []
                for w in (2 * pair, 2 * pair + 1):
                    active = w < n_tiles
                    c = w // tiles_per_ch if active else 0
                    o0 = (w % tiles_per_ch) * self.TILE_OUT if active else 0
                    tiles.append((active, c, o0, s0 + o0 * M))
                (actA, cA, oA, fA), (actB, cB, oB, fB) = tiles
                buf = np.zeros((self.TILE_IN, 2))
                if fA >= hist and actB and fB >= hist:
                    na, nb = n_in - (fA - hist), n_in - (fB - hist)
                    assert na >= 1 and nb >= 1                     # no size_t underflow on the device
                    if na >= self.TILE_IN and nb >= self.TILE_IN:
                        self.paths["full"] += 1
                        buf[:, 0] = chunk[cA, fA - hist: fA - hist + self.TILE_IN]
                        buf[:, 1] = chunk[cB, fB - hist: fB - hist + self.TILE_IN]
                    else:
                        self.paths["fast"] += 1
                        for i in range(self.TILE_IN):
                            buf[i, 0] = chunk[cA, fA - hist + i] if i < na else 0.0
                            buf[i, 1] = chunk[cB, fB - hist + i] if i < nb else 0.0
                else:
                    self.paths["generic"] += 1
                    for i in range(self.TILE_IN):
                        buf[i, 0] = self.row_at(cA, chunk, fA + i) if actA else 0.0
                        buf[i, 1] = self.row_at(cB, chunk, fB + i) if actB else 0.0
                for t in range(NT):
                    sx = t * self.STEP
                    assert sx + self.WIN2 <= self.TILE_IN
                    for r in range(R):
                        acc = buf[sx + r * M: sx + r * M + TPP].T @ self.g      # k = j - r M in [0, TPP)
                        o = t * R + r
                        if o < no - oA:
                            assert np.isnan(out[cA, oA + o])
                            out[cA, oA + o] = acc[0]
                        if actB and o < no - oB:
                            assert np.isnan(out[cB, oB + o])
                            out[cB, oB + o] = acc[1]
            assert not np.isnan(out).any()                      # every output written exactly once
        # rs_carry_kernel: next history = the last TPP-1 samples of [history | chunk]
        nxt = np.zeros_like(self.carry)
        for c in range(n_ch):
            for k in range(hist):
                nxt[c, k] = self.row_at(c, chunk, n_in + k)
        self.carry = nxt
        self.n_total += n_in
        self.m_next += max(no, 0)
        return out
