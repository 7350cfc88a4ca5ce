"""oracle -- CPU checkers for the LibRedio hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  It wraps

* ``oracle/liboracle.so``   -- our strict-f32 C restatement (``restated.c``; every function cites the
  reference ``file:line`` it follows), and
* ``oracle/_ref/*.so``      -- the UNMODIFIED vendored kissfft C sources of the reference, compiled where
  they lie under ``/root/reference`` by ``oracle/Makefile`` (``make ref``).  Built in the build container;
  the binaries travel to the GPU box, the sources never enter this repo.

Definitions the reference does not contain (Hann, |X|^2 averaging, FM discriminator, resampler) live in
``oracle/defined_f64.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, "_ref")


def build(ref: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is mounted, oracle/_ref/*.so."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/src/kissfft/libkissfft"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class _Cpx(C.Structure):
    _fields_ = [("r", C.c_float), ("i", C.c_float)]


class _OokResult(C.Structure):
    _fields_ = [
        ("a_bits", C.POINTER(C.c_uint8)), ("a_count", C.c_size_t),
        ("b_bits", C.POINTER(C.c_uint8)), ("b_count", C.c_size_t),
        ("block_sums", C.POINTER(C.c_float)), ("n_blocks", C.c_size_t),
        ("bits", C.POINTER(C.c_uint8)), ("n_bits", C.c_size_t),
        ("run_val", C.POINTER(C.c_uint32)), ("run_len", C.POINTER(C.c_uint32)), ("n_runs", C.c_size_t),
        ("n_bursts", C.c_size_t),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "restated.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build(ref=False)
        L = C.CDLL(path)
        vp, sz, fp = C.c_void_p, C.c_size_t, C.POINTER(C.c_float)
        L.orc_i2f.restype = C.c_float
        L.orc_i2f.argtypes = [C.c_uint8]
        L.orc_data_to_samples.restype = sz
        L.orc_data_to_samples.argtypes = [vp, sz, vp]
        L.orc_convolve_f32.restype = sz
        L.orc_convolve_f32.argtypes = [vp, sz, vp, sz, vp]
        L.orc_fir_decimate_cf32.restype = sz
        L.orc_fir_decimate_cf32.argtypes = [vp, sz, vp, sz, sz, vp, C.c_int]
        L.orc_window.argtypes = [sz, vp, C.c_int]
        L.orc_sinc.argtypes = [sz, C.c_float, vp]
        L.orc_lpf.argtypes = [sz, C.c_float, vp, C.c_int]
        L.orc_hpf.argtypes = [sz, C.c_float, vp, C.c_int]
        L.orc_bsf.argtypes = [sz, C.c_float, C.c_float, vp, C.c_int]
        L.orc_bpf.argtypes = [sz, C.c_float, C.c_float, vp, C.c_int]
        L.orc_fft.restype = C.c_int
        L.orc_fft.argtypes = [C.c_int, C.c_int, vp, vp]
        L.orc_fastfir.restype = sz
        L.orc_fastfir.argtypes = [vp, sz, sz, vp, sz, vp, C.c_int]
        L.orc_norm.restype = C.c_float
        L.orc_norm.argtypes = [C.c_float, C.c_float]
        L.orc_discretize.argtypes = [vp, sz, vp]
        L.orc_ook_decode.restype = C.c_int
        L.orc_ook_decode.argtypes = [vp, sz, C.c_uint, C.POINTER(_OokResult)]
        L.orc_ook_free.argtypes = [C.POINTER(_OokResult)]
        L.orc_test_set_trigger_guard.argtypes = [sz]
        L.orc_test_set_trigger_guard.restype = None
        L.orc_b2d.restype = sz
        L.orc_b2d.argtypes = [vp, sz]
        L.orc_eat.argtypes = [vp, vp, sz, vp]
        L.orc_kissfft_batch.restype = sz
        L.orc_kissfft_batch.argtypes = [vp, vp, C.c_int, C.c_int, sz, vp, vp]
        L.orc_chain_psd.restype = sz
        L.orc_chain_psd.argtypes = [vp, sz, vp, sz, sz, C.c_int, vp, vp, vp, vp]
        L.orc_chain_psd_mode.restype = sz
        L.orc_chain_psd_mode.argtypes = [vp, sz, vp, sz, sz, C.c_int, vp, vp, vp, vp, C.c_int]
        _lib = L
    return _lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


# ---- (1) unpack -------------------------------------------------------------------------------
def i2f(b: int) -> np.float32:
    return np.float32(lib().orc_i2f(int(b) & 0xFF))


def data_to_samples(data: np.ndarray) -> np.ndarray:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.size & 1:
        raise IndexError("odd byte count: the reference indexes i[1] out of bounds and panics (rtlsdr.rs:161)")
    out = np.empty(data.size // 2, dtype=np.complex64)
    lib().orc_data_to_samples(_ptr(data), data.size, _ptr(out))
    return out


# ---- (2) FIR ----------------------------------------------------------------------------------
def convolve(u: np.ndarray, v: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    if v.size == 0 or u.size < v.size:
        return np.empty(0, dtype=np.float32)
    y = np.empty(u.size - v.size + 1, dtype=np.float32)
    lib().orc_convolve_f32(_ptr(u), u.size, _ptr(v), v.size, _ptr(y))
    return y


def fir_decimate(x: np.ndarray, taps: np.ndarray, d: int, full: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.complex64)
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    if x.size < taps.size or taps.size == 0:
        return np.empty(0, dtype=np.complex64)
    nz = (x.size - taps.size) // d + 1
    z = np.empty(nz, dtype=np.complex64)
    got = lib().orc_fir_decimate_cf32(_ptr(x), x.size, _ptr(taps), taps.size, d, _ptr(z), int(full))
    assert got == nz
    return z


def lpf(m: int, fc: float, faithful: bool = False) -> np.ndarray:
    out = np.empty(m, dtype=np.float32)
    lib().orc_lpf(m, C.c_float(fc), _ptr(out), int(faithful))
    return out


def hpf(m: int, fc: float, faithful: bool = False) -> np.ndarray:
    out = np.empty(m, dtype=np.float32)
    lib().orc_hpf(m, C.c_float(fc), _ptr(out), int(faithful))
    return out


def bsf(m: int, fc1: float, fc2: float, faithful: bool = False) -> np.ndarray:
    out = np.empty(m, dtype=np.float32)
    lib().orc_bsf(m, C.c_float(fc1), C.c_float(fc2), _ptr(out), int(faithful))
    return out


def bpf(m: int, fc1: float, fc2: float, faithful: bool = False) -> np.ndarray:
    out = np.empty(m, dtype=np.float32)
    lib().orc_bpf(m, C.c_float(fc1), C.c_float(fc2), _ptr(out), int(faithful))
    return out


def window(m: int, faithful: bool = False) -> np.ndarray:
    out = np.empty(m + 1, dtype=np.float32)
    lib().orc_window(m, _ptr(out), int(faithful))
    return out


# ---- (3) FFT ----------------------------------------------------------------------------------
def fft(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """Restated kissfft; x is (..., nfft) complex64."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    n = x.shape[-1]
    flat = x.reshape(-1, n)
    out = np.empty_like(flat)
    for k in range(flat.shape[0]):
        rc = lib().orc_fft(n, int(inverse), _ptr(flat[k]), _ptr(out[k]))
        assert rc == 0
    return out.reshape(x.shape)


def fastfir(h: np.ndarray, x: np.ndarray, nfft: int = 0, flush: bool = False) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.complex64)
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty(x.size + 1, dtype=np.complex64)
    n = lib().orc_fastfir(_ptr(h), h.size, nfft, _ptr(x), x.size, _ptr(out), int(flush))
    return out[:n].copy()


# ---- (5) OOK ----------------------------------------------------------------------------------
def norm(re, im) -> np.float32:
    return np.float32(lib().orc_norm(C.c_float(float(re)), C.c_float(float(im))))


def norm_table() -> np.ndarray:
    """|i2f(b0) + j i2f(b1)| for all 65536 byte pairs, table[b0, b1]."""
    t = np.array([lib().orc_i2f(b) for b in range(256)], dtype=np.float32).astype(np.float64)
    s = t[:, None] * t[:, None] + t[None, :] * t[None, :]
    return np.sqrt(s).astype(np.float32)


def discretize(burst: np.ndarray) -> np.ndarray:
    burst = np.ascontiguousarray(burst, dtype=np.float32)
    bits = np.empty(burst.size, dtype=np.uint8)
    lib().orc_discretize(_ptr(burst), burst.size, _ptr(bits))
    return bits


def b2d(bits) -> int:
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    return int(lib().orc_b2d(_ptr(bits), bits.size))


def eat(bits, widths) -> list:
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    w = np.ascontiguousarray(widths, dtype=np.uint64)
    out = np.empty(w.size, dtype=np.uint64)
    lib().orc_eat(_ptr(bits), _ptr(w), w.size, _ptr(out))
    return [int(v) for v in out]


def set_trigger_guard_blocks(blocks: int) -> None:
    """Test hook: shrink the trigger's OOM guard (bitfount.rs:52, 50 000 blocks) to `blocks` blocks in the C restatement and the
    pure-Python one; 0 restores the reference's constant.  The CUDA side has LRC_OOK_TEST_GUARD_BLOCKS."""
    from . import restated_py
    lib().orc_test_set_trigger_guard(int(blocks) * 512)
    restated_py.TEST_GUARD_SAMPLES = int(blocks) * 512


def ook_decode(iq: np.ndarray, s_rate: int = 256000) -> dict:
    """Whole OOK chain on one finite u8-IQ capture (length = n_blocks*1024 bytes)."""
    iq = np.ascontiguousarray(iq, dtype=np.uint8)
    assert iq.size % 1024 == 0
    res = _OokResult()
    rc = lib().orc_ook_decode(_ptr(iq), iq.size // 1024, s_rate, C.byref(res))
    assert rc == 0

    def grab(p, n, dt):
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.empty(0, dtype=dt)

    out = {
        "a_packets": grab(res.a_bits, res.a_count * 36, np.uint8).reshape(-1, 36),
        "b_packets": grab(res.b_bits, res.b_count * 24, np.uint8).reshape(-1, 24),
        "block_sums": grab(res.block_sums, res.n_blocks, np.float32),
        "bits": grab(res.bits, res.n_bits, np.uint8),
        "run_val": grab(res.run_val, res.n_runs, np.uint32),
        "run_len": grab(res.run_len, res.n_runs, np.uint32),
        "n_bursts": int(res.n_bursts),
    }
    lib().orc_ook_free(C.byref(res))
    return out


# ---- the vendored reference itself (oracle/_ref) ----------------------------------------------
_ref_cache: dict = {}


def have_ref() -> bool:
    return os.path.exists(os.path.join(_REFDIR, "libkissfft_ref.so"))


def _ref(name: str) -> C.CDLL:
    if name not in _ref_cache:
        L = C.CDLL(os.path.join(_REFDIR, name))
        L.kiss_fft_alloc.restype = C.c_void_p
        L.kiss_fft_alloc.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.kiss_fft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _ref_cache[name] = L
    return _ref_cache[name]


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def ref_kissfft(x: np.ndarray, inverse: bool = False, opt: bool = False) -> np.ndarray:
    """kiss_fft() of the vendored C, reference build flags (opt=False: libkissfft/Makefile:4) or -O3."""
    L = _ref("libkissfft_ref_O3.so" if opt else "libkissfft_ref.so")
    x = np.ascontiguousarray(x, dtype=np.complex64)
    n = x.shape[-1]
    flat = x.reshape(-1, n)
    out = np.empty_like(flat)
    cfg = L.kiss_fft_alloc(n, int(inverse), None, None)
    for k in range(flat.shape[0]):
        L.kiss_fft(cfg, _ptr(flat[k]), _ptr(out[k]))
    _libc.free(cfg)
    return out.reshape(x.shape)


def ref_fftr(x: np.ndarray) -> np.ndarray:
    """kiss_fftr() of the vendored tools/kiss_fftr.c: real input (nfft even) -> nfft/2+1 bins."""
    L = _ref("libkissfftr_ref.so")
    L.kiss_fftr_alloc.restype = C.c_void_p
    L.kiss_fftr_alloc.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kiss_fftr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 2 + 1, dtype=np.complex64)
    cfg = L.kiss_fftr_alloc(x.size, 0, None, None)
    L.kiss_fftr(cfg, _ptr(x), _ptr(out))
    _libc.free(cfg)
    return out


def ref_fftri(X: np.ndarray) -> np.ndarray:
    """kiss_fftri() of the vendored tools/kiss_fftr.c: nfft/2+1 bins -> nfft reals (unscaled)."""
    L = _ref("libkissfftr_ref.so")
    L.kiss_fftr_alloc.restype = C.c_void_p
    L.kiss_fftr_alloc.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kiss_fftri.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    X = np.ascontiguousarray(X, dtype=np.complex64)
    n = 2 * (X.size - 1)
    out = np.empty(n, dtype=np.float32)
    cfg = L.kiss_fftr_alloc(n, 1, None, None)
    L.kiss_fftri(cfg, _ptr(X), _ptr(out))
    _libc.free(cfg)
    return out


def ref_fastfir(h: np.ndarray, x: np.ndarray, nfft: int = 0, flush: bool = False) -> np.ndarray:
    """kiss_fastfir() of the vendored tools/kiss_fastfir.c, driven the way do_file_filter (:341-390) does
    for one buffer holding the whole input."""
    L = _ref("libkissfastfir_ref.so")
    L.kiss_fastfir_alloc.restype = C.c_void_p
    L.kiss_fastfir_alloc.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.kiss_fastfir.restype = C.c_size_t
    L.kiss_fastfir.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    h = np.ascontiguousarray(h, dtype=np.complex64)
    x = np.ascontiguousarray(x, dtype=np.complex64)
    nf = C.c_size_t(nfft)
    cfg = L.kiss_fastfir_alloc(_ptr(h), h.size, C.byref(nf), None, None)
    nfft = nf.value
    inbuf = np.zeros(x.size + nfft, dtype=np.complex64)
    inbuf[: x.size] = x
    outbuf = np.zeros(x.size + nfft, dtype=np.complex64)
    off = C.c_size_t(0)
    n1 = L.kiss_fastfir(cfg, _ptr(inbuf), _ptr(outbuf), x.size, C.byref(off))
    res = [outbuf[:n1].copy()]
    if flush:
        out2 = np.zeros(nfft * 2, dtype=np.complex64)
        n2 = L.kiss_fastfir(cfg, _ptr(inbuf), _ptr(out2), 0, C.byref(off))
        res.append(out2[:n2].copy())
    _libc.free(cfg)
    return np.concatenate(res)


def ref_kiss_fn_ptrs(opt: bool = False):
    """(alloc, fft) raw function pointers of the vendored build, for orc_chain_psd."""
    L = _ref("libkissfft_ref_O3.so" if opt else "libkissfft_ref.so")
    return C.cast(L.kiss_fft_alloc, C.c_void_p), C.cast(L.kiss_fft, C.c_void_p)


def chain_psd_cpu(x: np.ndarray, taps: np.ndarray, d: int, nfft: int, win: np.ndarray,
                  use_ref: bool = True, opt: bool = False, full: bool = False):
    """CPU baseline of the headline chain; returns (psd_sum f64[nfft], n_frames).  ``full``: the FIR computes every
    lag like dsputils::convolve (dsputils.rs:30-32) and then keeps every d-th output."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    win = np.ascontiguousarray(win, dtype=np.float32)
    psd = np.zeros(nfft, dtype=np.float64)
    if use_ref and have_ref():
        a, f = ref_kiss_fn_ptrs(opt)
    else:
        a, f = C.c_void_p(None), C.c_void_p(None)
    nfr = lib().orc_chain_psd_mode(_ptr(x), x.size, _ptr(taps), taps.size, d, nfft, _ptr(win), _ptr(psd), a, f, int(full))
    return psd, int(nfr)


def kissfft_batch_cpu(x: np.ndarray, inverse: bool = False, opt: bool = False) -> np.ndarray:
    """(batch, nfft) frames through the vendored kiss_fft (or the restatement) in one C loop -- for timing."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    if have_ref():
        a, f = ref_kiss_fn_ptrs(opt)
    else:
        a, f = C.c_void_p(None), C.c_void_p(None)
    lib().orc_kissfft_batch(_ptr(x), _ptr(out), x.shape[-1], int(inverse), x.size // x.shape[-1], a, f)
    return out


def psdpng_rows(pcm: np.ndarray, nfft: int = 1024, navg: int = 20, remove_dc: bool = False, stereo: bool = False,
                fftr=None) -> np.ndarray:
    """tools/psdpng.c transform_signal (:120-185) statement by statement in f32, with kiss_fftr supplied by the
    vendored build (``fftr=ref_fftr``, the default when oracle/_ref exists) or by f64 numpy."""
    if fftr is None:
        fftr = ref_fftr if have_ref() else (lambda t: np.fft.rfft(t.astype(np.float64)).astype(np.complex64))
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    nfreqs = nfft // 2 + 1
    per = nfft * (2 if stereo else 1)
    rows = []
    mag2 = np.zeros(nfreqs, dtype=np.float32)
    avgctr = 0
    for f in range(pcm.size // per):
        buf = pcm[f * per:(f + 1) * per]
        if stereo:
            t = (buf[0::2].astype(np.int32) + buf[1::2].astype(np.int32)).astype(np.float32)      # :146
        else:
            t = buf.astype(np.float32)                                                            # :152
        if remove_dc:                                                                             # :156-161
            avg = np.float32(0)
            for v in t:
                avg = np.float32(avg + v)
            avg = np.float32(avg / np.float32(nfft))
            t = (t - avg).astype(np.float32)
        F = fftr(t)
        mag2 = (mag2 + (F.real * F.real + F.imag * F.imag).astype(np.float32)).astype(np.float32)  # :167
        avgctr += 1
        if avgctr == navg:                                                                        # :169-177
            avgctr = 0
            rows.append((10 * np.log10((mag2 / np.float32(navg) + np.float32(1)).astype(np.float64))).astype(np.float32))
            mag2 = np.zeros(nfreqs, dtype=np.float32)
    return np.stack(rows) if rows else np.zeros((0, nfreqs), dtype=np.float32)
