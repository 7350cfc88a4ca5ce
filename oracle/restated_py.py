"""TEST INFRASTRUCTURE ONLY -- a second, independent restatement of the stages the reference has no test for (unpack,
convolve, the bit-exact OOK chain), in plain Python.

oracle/restated.c is the oracle the GPU is held to; the reference holds no test, fixture or golden vector for this
chain, so the C restatement is pinned by hand-computed micro-cases only.  This module restates the same
reference code a second time, written straight from the Rust as a network of the reference's own blocks (one Python
generator per `kpn` block, messages flowing between them exactly as through the mpsc channels), with every f32
operation an explicit numpy.float32 scalar operation.  tests/test_cpu_oracle.py requires the two restatements to
agree BIT FOR BIT on seeded captures: block sums, burst count, the discretised bit stream, every (value, run) pair and
the decoded packets.  Pure-Python loops: small captures only.

Reference (paths relative to the LibRedio tree):
    i2f, data_to_samples      src/rtlsdr/src/rtlsdr.rs:159-162
    convolve                  src/dsputils/src/dsputils.rs:30-32
    x.norm()                  src/ratpak.rs:64-68 (num 0.1.22 Complex::norm = hypot)
    trigger                   src/bitfount/src/bitfount.rs:36-85
    discretize                src/bitfount/src/bitfount.rs:87-96
    rle, dle                  src/kpn/src/kpn.rs:17-38
    matchers A / B            src/ratpak.rs:88-97
    shaper_optional           src/kpn/src/kpn.rs:266-275
    b2d, eat                  src/kpn/src/kpn.rs:111-124
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
BLOCK = 512                      # samples per message (bitfount.rs:17,24: 1024 bytes)


def i2f(b: int) -> np.float32:
    """rtlsdr.rs:159: i as f32 / 127.0 - 1.0"""
    return F(F(F(b) / F(127.0)) - F(1.0))


def data_to_samples(data) -> list:
    """rtlsdr.rs:160-162: data.chunks(2).map(|i| Complex{re: i2f(i[0]), im: i2f(i[1])}); an odd length indexes
    i[1] out of bounds and panics"""
    out = []
    for k in range(0, len(data), 2):
        pair = data[k:k + 2]
        if len(pair) < 2:
            raise IndexError("index out of bounds: the len is 1 but the index is 1")      # the reference's panic
        out.append((i2f(int(pair[0])), i2f(int(pair[1]))))
    return out


def convolve(u, v) -> list:
    """dsputils.rs:30-32: u.windows(v.len()).map(|x| x.iter().zip(v.iter()).map(|(&x,&y)| x*y)
    .fold(Float::zero(), |a, b| a + b)) -- valid-mode correlation (taps not reversed), every product rounded to f32,
    summed left to right from 0.0 in f32"""
    out = []
    for i in range(len(u) - len(v) + 1):
        acc = F(0.0)
        for j in range(len(v)):
            acc = F(acc + F(F(u[i + j]) * F(v[j])))
        out.append(acc)
    return out


def rtl_source_cmplx(iq: np.ndarray):
    """bitfount.rs:16-34 + rtlsdr.rs:160-162: one Vec<Complex<f32>> of 512 samples per 1024 bytes"""
    lut = [i2f(b) for b in range(256)]
    for k in range(0, len(iq), 2 * BLOCK):
        chunk = iq[k:k + 2 * BLOCK]
        yield [(lut[int(chunk[2 * j])], lut[int(chunk[2 * j + 1])]) for j in range(len(chunk) // 2)]


def norm(re: np.float32, im: np.float32) -> np.float32:
    """num::Complex::norm = re.hypot(im); both restatements define hypot as sqrt in double of the exact double sum
    of squares, rounded once to f32."""
    return F(math.sqrt(float(re) * float(re) + float(im) * float(im)))


def cross_applicator_vecs(src, f):
    """kpn.rs:170-174: map f over every Vec"""
    for v in src:
        yield [f(*x) for x in v]


TEST_GUARD_SAMPLES = 0      # test hook: non-zero replaces the OOM guard's 1000*trigger_duration*BLOCK (bitfount.rs:52) on this side


def trigger(src):
    """bitfount.rs:36-85, statement by statement"""
    trigger_duration = 50
    guard = TEST_GUARD_SAMPLES if TEST_GUARD_SAMPLES else 1000 * trigger_duration * BLOCK
    trig = 0
    sample_buffer = [F(0.0)]
    threshold = F(0.0)
    for samples in src:
        trig -= 1
        s = F(0.0)
        for x in samples:                                   # iter().sum(): left fold from 0.0
            s = F(s + x)
        if len(sample_buffer) > guard:
            sample_buffer = [F(0.0)]
        if threshold == F(0.0):
            threshold = s
        if trig < 0:
            threshold = F(threshold + F(s / F(1000.0)))
            threshold = F(threshold - F(threshold * F(0.002)))
        if s > F(threshold * F(4.0)):
            trig = trigger_duration
        if trig > 1:
            sample_buffer.extend(samples)
        if trig == 0:
            yield sample_buffer
            sample_buffer = []


def discretize(src):
    """bitfount.rs:87-96: one message per sample"""
    for buf in src:
        mx = F(0.0)
        for y in buf:
            mx = y if y > mx else mx                        # f32::max with no NaNs in play
        half = F(mx / F(2.0))
        for x in buf:
            yield 1 if x > half else 0


def rle(src):
    """kpn.rs:17-29: emits the previous run on change, never flushes the last one"""
    it = iter(src)
    try:
        x = next(it)
    except StopIteration:
        return
    i = 1
    for y in it:
        if y != x:
            yield (x, i)
            i = 1
        else:
            i += 1
        x = y


def dle(src, s_rate: int):
    """kpn.rs:32-38: ct as f32 / s_rate as f32"""
    for x, ct in src:
        yield (x, F(F(ct) / F(s_rate)))


def _in(d: np.float32, lo: float, hi: float) -> bool:
    """an f32 range pattern lo...hi (inclusive, literals are f32)"""
    return F(lo) <= d <= F(hi)


def matcher_a(src):
    """ratpak.rs:88-92: the first run decides whether the NEXT run is consumed"""
    it = iter(src)
    for v, d in it:
        if v == 1 and _in(d, 2e-4, 6e-4):
            try:
                v2, d2 = next(it)                           # a.next().unwrap(): a closed port ends the block
            except StopIteration:
                return
            if v2 == 0 and _in(d2, 1.5e-3, 2.5e-3):
                yield 0
            elif v2 == 0 and _in(d2, 3.5e-3, 4.5e-3):
                yield 1
            else:
                yield None
        else:
            yield None


def matcher_b(src):
    """ratpak.rs:93-97"""
    def ok(d):
        return _in(d, 125e-6, 250e-6) or _in(d, 500e-6, 650e-6)
    it = iter(src)
    for v, d in it:
        if v == 1 and ok(d):
            try:
                v2, e = next(it)
            except StopIteration:
                return
            if v2 == 0 and ok(e):
                yield 1 if d > e else 0
            else:
                yield None
        else:
            yield None


def shaper_optional(src, l: int):
    """kpn.rs:266-275"""
    x = []
    for y in src:
        if y is not None:
            x.append(y)
        elif len(x) == l:
            yield list(x)
            x = []
        else:
            x = []


def b2d(xs) -> int:
    """kpn.rs:111-113"""
    return sum((1 << (len(xs) - i - 1)) * xs[i] for i in range(len(xs)))


def eat(x, widths):
    """kpn.rs:116-124"""
    i, out = 0, []
    for w in widths:
        out.append(b2d(x[i:i + w]))
        i += w
    return out


def ook_decode(iq: np.ndarray, s_rate: int = 256000) -> dict:
    """The graph of ratpak.rs:60-111 on one finite capture; same result layout as oracle.ook_decode."""
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    assert iq.size % (2 * BLOCK) == 0
    env_blocks = list(cross_applicator_vecs(rtl_source_cmplx(iq), norm))
    block_sums = []
    for blk in env_blocks:
        s = F(0.0)
        for x in blk:
            s = F(s + x)
        block_sums.append(s)
    bursts = list(trigger(iter(env_blocks)))
    bits = list(discretize(iter(bursts)))
    runs = list(rle(iter(bits)))
    durations = list(dle(iter(runs), s_rate))               # fork: both matchers see every duration (kpn.rs:182)
    a = list(shaper_optional(matcher_a(iter(durations)), 36))
    b = list(shaper_optional(matcher_b(iter(durations)), 24))
    return {
        "a_packets": np.array(a, dtype=np.uint8).reshape(-1, 36),
        "b_packets": np.array(b, dtype=np.uint8).reshape(-1, 24),
        "block_sums": np.array(block_sums, dtype=np.float32),
        "bits": np.array(bits, dtype=np.uint8),
        "run_val": np.array([v for v, _ in runs], dtype=np.uint32),
        "run_len": np.array([n for _, n in runs], dtype=np.uint32),
        "n_bursts": len(bursts),
    }
