"""GPU parity, part 1: unpack, FIR+decimate, FFT, PSD, fused chain -- all through the C ABI, checked
against oracle/ (the CPU restatement, itself pinned to the vendored kissfft and the golden fixtures).

Tolerances are BASELINE.json's: bit-exact for the u8 unpack; max |err| <= 1e-4 x RMS(reference output)
for the float FIR / FFT stages.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import defined_f64 as D
from libredio_b200 import capi, synth

pytestmark = pytest.mark.gpu

TOL = 1e-4   # x RMS of the reference output (north_star)


def rms(a):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a, dtype=np.complex128)) ** 2)))


def dev(a, ctx):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)


def assert_close_rms(got, ref, tol=TOL):
    ref = np.asarray(ref)
    got = np.asarray(got)
    assert got.shape == ref.shape
    err = float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128)))) if ref.size else 0.0
    assert err <= tol * rms(ref), f"max|err| {err:.3e} > {tol} x rms {rms(ref):.3e}"


# ---- (1) unpack: bit-exact --------------------------------------------------------------------------
def test_unpack_all_256_values_bit_exact(ctx):
    from libredio_b200 import blocks
    b = np.arange(256, dtype=np.uint8).repeat(2)          # every byte as I and as Q
    got = blocks.data_to_samples(ctx, dev(b, ctx)).cpu().numpy()
    ref = oracle.data_to_samples(b)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # hand-computed micro-cases
    t = got.view(np.float32)[0::2]
    assert t[0] == -1.0 and t[127] == 0.0 and t[254] == 1.0 and t[255] == np.float32(1.007874)


@pytest.mark.parametrize("nbytes", [2, 14, 16, 30, 4096, 1024 * 1024 + 6, 2_400_000 * 2])
def test_unpack_random_bit_exact(ctx, nbytes):
    from libredio_b200 import blocks
    rng = np.random.default_rng(nbytes)
    b = rng.integers(0, 256, nbytes, dtype=np.uint8)
    got = blocks.data_to_samples(ctx, dev(b, ctx)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), oracle.data_to_samples(b).view(np.uint32))


def test_unpack_unaligned_and_empty(ctx):
    from libredio_b200 import blocks
    rng = np.random.default_rng(5)
    b = rng.integers(0, 256, 1001 * 2 + 2, dtype=np.uint8)
    d = dev(b, ctx)[2:]                                     # 2-byte offset: not 16-byte aligned
    got = blocks.data_to_samples(ctx, d.contiguous() if not d.is_contiguous() else d).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), oracle.data_to_samples(b[2:]).view(np.uint32))
    assert blocks.data_to_samples(ctx, dev(np.empty(0, np.uint8), ctx)).numel() == 0


def test_unpack_odd_length_fails_like_reference(ctx):
    from libredio_b200 import blocks
    with pytest.raises(capi.LrcError) as e:
        blocks.data_to_samples(ctx, dev(np.zeros(7, np.uint8), ctx))
    assert e.value.status == capi.ERR_ODD_LENGTH
    with pytest.raises(IndexError):
        oracle.data_to_samples(np.zeros(7, np.uint8))


# ---- (2) FIR + decimate ----------------------------------------------------------------------------
@pytest.mark.parametrize("n", [64, 73, 74, 640, 8960 + 54, 8960 + 55, 8960 + 64, 100_003, 2_400_000])
def test_fir64_decim10_cf32_vs_oracle(ctx, n):
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    x = synth.cf32_noise_tones(n, seed=n)
    fir = blocks.Fir(ctx, taps, 10)
    got = fir.run(dev(x, ctx)).cpu().numpy()
    ref = oracle.fir_decimate(x, taps, 10)
    assert got.shape == ref.shape == ((n - 64) // 10 + 1,)
    assert_close_rms(got, ref)
    fir.close()


def test_fir_shorter_than_taps_is_empty(ctx):
    from libredio_b200 import blocks
    fir = blocks.Fir(ctx, synth.lpf_taps(64, 0.04), 10)
    assert fir.run(dev(synth.cf32_noise_tones(63), ctx)).numel() == 0
    assert oracle.fir_decimate(synth.cf32_noise_tones(63), synth.lpf_taps(64, 0.04), 10).size == 0
    fir.close()


def test_fir_multichannel_and_unaligned(ctx):
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    n_ch, n = 5, 20_001                                   # odd stride: rows 2..4 are not 16-byte aligned
    x = np.stack([synth.cf32_noise_tones(n, seed=100 + c) for c in range(n_ch)])
    fir = blocks.Fir(ctx, taps, 10)
    got = fir.run(dev(x, ctx)).cpu().numpy()
    for c in range(n_ch):
        assert_close_rms(got[c], oracle.fir_decimate(x[c], taps, 10))
    fir.close()


@pytest.mark.parametrize("ntaps,decim,n", [(1, 1, 100), (5, 3, 1000), (33, 1, 5000), (64, 5, 30_000), (129, 16, 40_000), (4, 10, 1003)])
def test_fir_generic_shapes_vs_oracle(ctx, ntaps, decim, n):
    from libredio_b200 import blocks
    rng = np.random.default_rng(ntaps * 1000 + decim)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    x = synth.cf32_noise_tones(n, seed=7)
    fir = blocks.Fir(ctx, taps, decim)
    got = fir.run(dev(x, ctx)).cpu().numpy()
    assert_close_rms(got, oracle.fir_decimate(x, taps, decim))
    # 'full' (convolve everything, then stride) is the same function
    assert np.array_equal(oracle.fir_decimate(x, taps, decim, full=True), oracle.fir_decimate(x, taps, decim))
    fir.close()


@pytest.mark.parametrize("ntaps,decim", [(64, 4), (33, 4), (128, 5), (17, 5), (64, 8), (127, 8), (31, 10), (128, 10), (64, 16), (100, 16), (2, 16)])
def test_fir_tile_shapes_multichannel_vs_oracle(ctx, ntaps, decim, monkeypatch):
    """cf32, ntaps <= 128, decim in {4,5,8,10,16}: the padded-chunk tile kernel (k_fir_gentile.cu).  Several tiles per channel, a
    ragged last tile, even stride (TMA path) and odd stride (cooperative-copy path), and the same bits as the one-thread-per-output
    kernel it replaces (same ascending-tap fma chain per lane)."""
    from libredio_b200 import blocks
    rng = np.random.default_rng(ntaps * 100 + decim)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    n_ch = 3
    for n in (decim * 2600 + ntaps + 3, decim * 1024 + ntaps - 1, ntaps, 20_001):
        x = np.stack([synth.cf32_noise_tones(n, seed=50 + c) for c in range(n_ch)])
        fir = blocks.Fir(ctx, taps, decim)
        got = fir.run(dev(x, ctx))
        monkeypatch.setenv("LRC_FIR_NO_GENTILE", "1")
        old = fir.run(dev(x, ctx))
        monkeypatch.delenv("LRC_FIR_NO_GENTILE")
        assert torch.equal(got.view(torch.float32), old.view(torch.float32))
        got = got.cpu().numpy()
        for c in range(n_ch):
            ref = oracle.fir_decimate(x[c], taps, decim)
            assert got[c].shape == ref.shape
            assert_close_rms(got[c], ref)
        fir.close()


@pytest.mark.parametrize("chunks", [[5000], [63, 1, 1, 700, 4235], [4096, 4096, 13], [16] * 40 + [4500]])
def test_fir_stream_tile_shape_is_seam_exact(ctx, chunks):
    from libredio_b200 import blocks
    taps = synth.lpf_taps(96, 0.05)
    n, n_ch = sum(chunks), 3
    x = np.stack([synth.cf32_noise_tones(n, seed=c) for c in range(n_ch)])
    fir = blocks.Fir(ctx, taps, 8)
    whole = fir.run(dev(x, ctx)).cpu().numpy()
    st = blocks.FirStream(fir, n_ch, max(chunks), u8=False)
    outs, pos = [], 0
    for c in chunks:
        outs.append(st.push(dev(x[:, pos:pos + c], ctx)).cpu().numpy())
        pos += c
    got = np.concatenate(outs, axis=1)
    assert got.shape == whole.shape
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    st.close(); fir.close()


def test_fir_matches_reference_convolve_on_real_planes(ctx):
    """dsputils::convolve is real-only: the cf32 kernel must equal convolve() on each plane."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    x = synth.cf32_noise_tones(5000, seed=9)
    fir = blocks.Fir(ctx, taps, 1)
    got = fir.run(dev(x, ctx)).cpu().numpy()
    assert_close_rms(got.real, oracle.convolve(x.real.copy(), taps))
    assert_close_rms(got.imag, oracle.convolve(x.imag.copy(), taps))
    fir.close()


@pytest.mark.parametrize("n", [64, 8960 + 54, 250_000])
def test_fir_fused_u8_vs_oracle(ctx, n):
    """config 1: u8 IQ -> i2f -> FIR64 /10, unpack fused into the FIR tile."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    iq = synth.iq_tone_noise_u8(n, seed=1)
    fir = blocks.Fir(ctx, taps, 10)
    got = fir.run_u8(dev(iq, ctx)).cpu().numpy()
    ref = oracle.fir_decimate(oracle.data_to_samples(iq), taps, 10)
    assert_close_rms(got, ref)
    fir.close()


def test_fir_nan_taps_rejected(ctx):
    """dsputils::lpf yields NaN taps in the reference (window bug); the GPU plan refuses them loudly."""
    from libredio_b200 import blocks
    bad = oracle.lpf(64, 0.04, faithful=True)
    assert not np.all(np.isfinite(bad))
    with pytest.raises(capi.LrcError):
        blocks.Fir(ctx, bad, 10)


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("chunks", [[5000], [63, 1, 1, 700, 4235], [9014, 9014, 13], [10] * 50 + [4500]])
def test_fir_stream_is_seam_exact(ctx, chunks, u8):
    """output must be independent of chunking: bit-identical to one call over the whole stream."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    n, n_ch = sum(chunks), 3
    if u8:
        x = np.stack([synth.iq_tone_noise_u8(n, seed=c) for c in range(n_ch)])
    else:
        x = np.stack([synth.cf32_noise_tones(n, seed=c) for c in range(n_ch)])
    fir = blocks.Fir(ctx, taps, 10)
    whole = (fir.run_u8(dev(x, ctx)) if u8 else fir.run(dev(x, ctx))).cpu().numpy()
    st = blocks.FirStream(fir, n_ch, max(chunks), u8=u8)
    outs, pos = [], 0
    for c in chunks:
        w = 2 * c if u8 else c
        p = 2 * pos if u8 else pos
        outs.append(st.push(dev(x[:, p:p + w], ctx)).cpu().numpy())
        pos += c
    got = np.concatenate(outs, axis=1)
    assert got.shape == whole.shape
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    st.close(); fir.close()


# ---- (3) FFT -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("inv", [0, 1])
def test_fft_vs_golden_and_oracle(ctx, golden, n, inv):
    from libredio_b200 import blocks
    import tests.golden.make_golden as mg
    x = mg.fft_input(n)
    f = blocks.Fft(ctx, n, inv)
    got = f.run(dev(x, ctx)).cpu().numpy()
    assert_close_rms(got, golden[f"{'inv' if inv else 'fwd'}_{n}"])          # the reference's own output
    assert_close_rms(got, oracle.fft(x, bool(inv)))                           # the restatement
    # and the self-test of the reference: SNR vs an f64 DFT (test_vs_dft.c), >= 100 dB (mk_test.py:30)
    exact = np.fft.ifft(x.astype(np.complex128)) * n if inv else np.fft.fft(x.astype(np.complex128))
    assert D.snr_db(exact, got) >= 100.0
    f.close()


@pytest.mark.parametrize("n,batch", [(1024, 1), (1024, 7), (1024, 4097), (64, 1000), (8, 3), (8192, 5), (256, 33)])
def test_fft_batched_inplace_and_ragged_batch(ctx, n, batch):
    from libredio_b200 import blocks
    rng = np.random.default_rng(n + batch)
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    f = blocks.Fft(ctx, n, 0)
    ref = np.fft.fft(x.astype(np.complex128), axis=-1)
    d = dev(x, ctx)
    got = f.run(d).cpu().numpy()
    assert_close_rms(got, ref)
    f.run(d, inplace=True)                                 # fin == fout is allowed (kiss_fft.c:373-379)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), got.view(np.uint32))
    f.close()


def test_fft_golden_vector_of_reference_tree(ctx, golden):
    """test/fft.py:95-98, tolerance 1e-5 (:104): 8-point real sequence and its spectrum."""
    from libredio_b200 import blocks
    f = blocks.Fft(ctx, 8, 0)
    got = f.run(dev(golden["fftpy_tvec"].astype(np.complex64), ctx)).cpu().numpy()
    assert np.max(np.abs(got - golden["fftpy_Ftvec"])) < 1e-5
    f.close()


def test_fft_roundtrip_is_scaled_by_n(ctx):
    """neither direction scales (libkissfft/README:105): ifft(fft(x)) == n*x"""
    from libredio_b200 import blocks
    n = 1024
    x = synth.cf32_noise_tones(n * 9, seed=3).reshape(9, n)
    fw, bw = blocks.Fft(ctx, n, 0), blocks.Fft(ctx, n, 1)
    back = bw.run(fw.run(dev(x, ctx))).cpu().numpy()
    assert_close_rms(back, n * x)
    fw.close(); bw.close()


def test_fft_block_size_mismatch_and_unsupported(ctx):
    from libredio_b200 import blocks
    f = blocks.Fft(ctx, 1024, 0)
    with pytest.raises(capi.LrcError) as e:                 # assert!(din.len() == block_size) kissfft.rs:24
        f.run(dev(np.zeros(1000, np.complex64), ctx))
    assert e.value.status == capi.ERR_LENGTH
    with pytest.raises(capi.LrcError) as e:
        f.run_host(np.zeros(1000, np.complex64))
    assert e.value.status == capi.ERR_LENGTH
    with pytest.raises(capi.LrcError) as e:                 # kissfft has no size limit; ours is said loudly
        blocks.Fft(ctx, 10_000, 0)
    assert e.value.status == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.LrcError) as e:
        blocks.Fft(ctx, 16384, 0)
    assert e.value.status == capi.ERR_UNSUPPORTED
    f.close()


# ---- sizes that are not powers of two (kissfft radix 3 / 5 / generic butterflies) --------------------
@pytest.mark.parametrize("n", [3, 5, 6, 7, 15, 30, 45, 97, 100, 240, 243, 625, 1000, 1001, 1536, 3125, 6000, 7919])
@pytest.mark.parametrize("inv", [0, 1])
def test_fft_mixed_radix_vs_golden_of_vendored_kissfft(ctx, golden, n, inv):
    from libredio_b200 import blocks
    import tests.golden.make_golden as mg
    x = mg.fft_input(n)
    f = blocks.Fft(ctx, n, inv)
    got = f.run(dev(x, ctx)).cpu().numpy()
    assert_close_rms(got, golden[f"{'inv' if inv else 'fwd'}_{n}"])          # the reference's own output
    exact = np.fft.ifft(x.astype(np.complex128)) * n if inv else np.fft.fft(x.astype(np.complex128))
    assert D.snr_db(exact, got) >= 100.0                                      # mk_test.py:30
    f.close()


@pytest.mark.parametrize("n,batch", [(1000, 1), (1000, 37), (15, 5000), (7919, 3), (6000, 20), (97, 300)])
def test_fft_mixed_radix_batched_and_inplace(ctx, n, batch):
    from libredio_b200 import blocks
    rng = np.random.default_rng(n + batch)
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    f = blocks.Fft(ctx, n, 0)
    d = dev(x, ctx)
    got = f.run(d).cpu().numpy()
    assert_close_rms(got, np.fft.fft(x.astype(np.complex128), axis=-1))
    f.run(d, inplace=True)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), got.view(np.uint32))
    b = blocks.Fft(ctx, n, 1)
    assert_close_rms(b.run(dev(got, ctx)).cpu().numpy(), n * x.astype(np.complex128))     # unscaled both ways
    f.close(); b.close()


# ---- real-input pair (tools/kiss_fftr.c) ---------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 8, 30, 256, 1000, 1024, 4096])
def test_rfft_pair_vs_golden_of_vendored_kiss_fftr(ctx, golden, n):
    from libredio_b200 import blocks
    import tests.golden.make_golden as mg
    t = mg.rfft_input(n)
    fw, bw = blocks.Rfft(ctx, n, 0), blocks.Rfft(ctx, n, 1)
    F = fw.run(dev(t, ctx)).cpu().numpy()
    assert F.shape == (n // 2 + 1,)
    assert_close_rms(F, golden[f"rfft_{n}"])
    assert_close_rms(F, np.fft.rfft(t.astype(np.float64)))
    assert F[0].imag == 0.0 and F[-1].imag == 0.0                              # kiss_fftr.c:101-106
    back = bw.run(dev(golden[f"rfft_{n}"], ctx)).cpu().numpy()
    assert_close_rms(back, golden[f"irfft_{n}"])
    assert_close_rms(back, n * t.astype(np.float64))                           # unscaled: round trip = nfft x
    fw.close(); bw.close()


def test_rfft_batched_golden_vector_and_errors(ctx, golden):
    from libredio_b200 import blocks
    # the reference tree's only hard-coded vector is a REAL sequence (test/fft.py:95-98)
    f8 = blocks.Rfft(ctx, 8, 0)
    got = f8.run(dev(golden["fftpy_tvec"], ctx)).cpu().numpy()
    assert np.max(np.abs(got - golden["fftpy_Ftvec"][:5])) < 1e-5
    f8.close()
    rng = np.random.default_rng(5)
    x = rng.standard_normal((300, 1024)).astype(np.float32)
    f = blocks.Rfft(ctx, 1024, 0)
    assert_close_rms(f.run(dev(x, ctx)).cpu().numpy(), np.fft.rfft(x.astype(np.float64), axis=-1))
    f.close()
    with pytest.raises(capi.LrcError) as e:                 # "Real FFT optimization must be even." kiss_fftr.c:34-37
        blocks.Rfft(ctx, 1023, 0)
    assert e.value.status == capi.ERR_INVALID


def test_fft_host_entry_point(ctx):
    from libredio_b200 import blocks
    x = synth.cf32_noise_tones(1024 * 3, seed=11)
    f = blocks.Fft(ctx, 1024, 0)
    got = f.run_host(x)
    assert_close_rms(got.reshape(3, 1024), oracle.fft(x.reshape(3, 1024)))
    f.close()


# ---- tools/psdpng.c spectrogram rows ------------------------------------------------------------------
@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("dc", [False, True])
def test_psdpng_rows_vs_golden_and_oracle(ctx, golden, stereo, dc):
    from libredio_b200 import blocks
    import tests.golden.make_golden as mg
    pcm = mg.psdpng_input(stereo)
    got = blocks.psdpng_rows(ctx, dev(pcm, ctx), 256, 3, dc, stereo).cpu().numpy()
    ref = golden[f"psdpng_{int(stereo)}{int(dc)}"]
    assert got.shape == ref.shape == (2, 129)                  # 7 whole frames -> 2 rows of navg = 3, rest dropped
    assert_close_rms(got, ref)
    assert_close_rms(got, oracle.psdpng_rows(pcm, 256, 3, dc, stereo))
    if dc:                                                     # -a removes the offset: the DC bin collapses
        assert got[0, 0] < golden[f"psdpng_{int(stereo)}0"][0, 0] - 20.0


def test_psdpng_default_shape_and_short_input(ctx):
    from libredio_b200 import blocks
    rng = np.random.default_rng(12)
    pcm = (3000 * rng.standard_normal(1024 * 45 + 17)).astype(np.int16)
    got = blocks.psdpng_rows(ctx, dev(pcm, ctx)).cpu().numpy()      # nfft 1024, navg 20 (psdpng.c:26,29)
    assert got.shape == (2, 513)
    assert_close_rms(got, oracle.psdpng_rows(pcm))
    assert blocks.psdpng_rows(ctx, dev(pcm[:1024 * 19], ctx)).shape[0] == 0   # fewer than navg frames: no row


# ---- window + |X|^2 averaging ------------------------------------------------------------------------
@pytest.mark.parametrize("nfft,k,frames", [(1024, 1, 3), (1024, 64, 256), (1024, 20, 47), (1024, 300, 650),
                                           (256, 8, 64), (64, 5, 100), (4096, 4, 16), (16, 3, 10)])
def test_psd_rows_vs_f64(ctx, nfft, k, frames):
    from libredio_b200 import blocks
    x = synth.cf32_noise_tones(nfft * frames, seed=nfft + k)
    p = blocks.Psd(ctx, nfft, capi.WINDOW_HANN)
    got = p.run(dev(x, ctx), k).cpu().numpy()
    ref = D.psd_rows(x, nfft, k, D.hann_periodic(nfft))
    assert got.shape == ref.shape == (frames // k, nfft)
    assert np.max(np.abs(got - ref)) <= TOL * np.sqrt(np.mean(ref ** 2))
    p.close()


def test_psd_no_window_matches_psdpng_accumulation(ctx):
    """tools/psdpng.c:165-166: mag2 += r^2 + i^2 per frame on kiss_fft output, no window."""
    from libredio_b200 import blocks
    nfft, k = 1024, 20                                       # navg = 20, psdpng.c:29
    x = synth.cf32_noise_tones(nfft * k * 2, seed=21)
    p = blocks.Psd(ctx, nfft, capi.WINDOW_NONE)
    got = p.run(dev(x, ctx), k).cpu().numpy()
    spec = oracle.fft(x.reshape(-1, nfft))
    mag2 = (spec.real.astype(np.float64) ** 2 + spec.imag.astype(np.float64) ** 2).reshape(2, k, nfft).sum(1) / k
    assert np.max(np.abs(got - mag2)) <= TOL * np.sqrt(np.mean(mag2 ** 2))
    p.close()


# ---- fused chain ---------------------------------------------------------------------------------------
def chain_ref(x, taps, decim, nfft, k):
    z = oracle.fir_decimate(x, taps, decim)                  # strict-f32 reference FIR
    return D.psd_rows(z, nfft, k, D.hann_periodic(nfft))


@pytest.mark.parametrize("frames,k", [(1, 1), (3, 1), (16, 16), (40, 8), (70, 35), (130, 64), (50, 50)])
def test_chain_fused_vs_oracle(ctx, frames, k):
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    n = frames * 10240 + 54 + 17                             # a few samples more than whole frames
    x = synth.cf32_noise_tones(n, seed=frames)
    ch = blocks.Chain(ctx, taps, 10, 1024)
    assert ch.frames(n) == frames
    got = ch.run(dev(x, ctx), k).cpu().numpy()
    ref = chain_ref(x, taps, 10, 1024, k)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL * np.sqrt(np.mean(ref ** 2))
    ch.close()


def test_chain_fused_equals_unfused_kernels(ctx):
    """linearity/consistency at a size the oracle would not finish quickly: the fused kernel must agree with
    FIR kernel -> PSD kernel to float rounding."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    frames, k = 2048, 64
    n = frames * 10240 + 54
    g = torch.Generator(device=ctx.tdev).manual_seed(1)
    x = torch.randn(n, 2, device=ctx.tdev, generator=g)
    x = torch.view_as_complex(x)
    ch, fir, psd = blocks.Chain(ctx, taps, 10, 1024), blocks.Fir(ctx, taps, 10), blocks.Psd(ctx, 1024)
    a = ch.run(x, k)
    b = psd.run(fir.run(x)[: frames * 1024], k)
    assert a.shape == b.shape == (frames // k, 1024)
    err = (a - b).abs().max().item()
    assert err <= 1e-5 * b.pow(2).mean().sqrt().item()
    ch.close(); fir.close(); psd.close()


def test_chain_generic_shape_falls_back_to_unfused(ctx):
    from libredio_b200 import blocks
    rng = np.random.default_rng(0)
    taps = (rng.standard_normal(33) / 6).astype(np.float32)
    frames, k, nfft, decim = 24, 4, 256, 4
    n = frames * nfft * decim + 33 - decim
    x = synth.cf32_noise_tones(n, seed=5)
    ch = blocks.Chain(ctx, taps, decim, nfft)
    got = ch.run(dev(x, ctx), k).cpu().numpy()
    ref = chain_ref(x, taps, decim, nfft, k)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL * np.sqrt(np.mean(ref ** 2))
    ch.close()


# the generic fused instances (chain_generic.cuh): ntaps <= 128 (run-time, zero-padded to 64 / 128), decim in {4,5,8,10,16},
# nfft in {512,1024,2048}.  Odd ntaps exercise the hand-copied last sample of a tile, decim 4/8/16 the padded chunk layout.
GENERIC_SHAPES = [(33, 4, 512), (64, 4, 1024), (128, 4, 2048), (17, 5, 512), (64, 5, 1024), (97, 5, 2048),
                  (64, 8, 512), (128, 8, 1024), (31, 8, 2048), (48, 10, 512), (63, 10, 1024), (128, 10, 1024),
                  (64, 10, 2048), (5, 16, 512), (64, 16, 1024), (127, 16, 1024), (64, 16, 2048), (128, 16, 2048), (77, 10, 2048)]


@pytest.mark.parametrize("ntaps,decim,nfft", GENERIC_SHAPES)
def test_chain_generic_shape_runs_a_fused_instance(ctx, ntaps, decim, nfft, monkeypatch):
    from libredio_b200 import blocks
    monkeypatch.setenv("LRC_CHAIN_GENERIC_ALL", "1")         # every instance, not only the ones the default policy prefers
    rng = np.random.default_rng(ntaps * 1000 + decim * 10 + nfft)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    frames, k = 21, 3
    n = (frames - 1) * nfft * decim + (nfft - 1) * decim + ntaps          # exactly what lrc_chain_frames asks for
    x = synth.cf32_noise_tones(n, seed=ntaps + decim)
    ch = blocks.Chain(ctx, taps, decim, nfft)
    assert ch.kind == 2, "no fused generic instance for this shape"
    assert ch.frames(n) == frames and ch.frames(n - 1) == frames - 1
    # the input ends exactly where the last frame ends: a kernel that reads one sample further faults or reads the guard
    xd = torch.full((n + 64,), float("nan"), dtype=torch.complex64, device=ctx.tdev)
    xd[:n] = dev(x, ctx)
    got = ch.run(xd[:n], k).cpu().numpy()
    ref = chain_ref(x, taps, decim, nfft, k)
    assert got.shape == ref.shape == (frames // k, nfft)
    assert np.isfinite(got).all()
    assert np.max(np.abs(got - ref)) <= TOL * np.sqrt(np.mean(ref ** 2))
    ch.close()


@pytest.mark.parametrize("ntaps,decim,nfft", [(64, 4, 1024), (100, 8, 2048), (64, 16, 1024), (33, 5, 512), (64, 16, 2048), (64, 10, 2048)])
def test_chain_generic_fused_equals_unfused_kernels_at_size(ctx, ntaps, decim, nfft, monkeypatch):
    """many CTAs, many items per CTA, rows of several work items: fused generic == FIR kernel -> PSD kernel to float rounding"""
    from libredio_b200 import blocks
    rng = np.random.default_rng(7)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    frames, k = 1500, 50
    n = (frames - 1) * nfft * decim + (nfft - 1) * decim + ntaps
    g = torch.Generator(device=ctx.tdev).manual_seed(3)
    x = torch.view_as_complex(torch.randn(n, 2, device=ctx.tdev, generator=g))
    monkeypatch.setenv("LRC_CHAIN_GENERIC_ALL", "1")
    fused = blocks.Chain(ctx, taps, decim, nfft)
    monkeypatch.delenv("LRC_CHAIN_GENERIC_ALL")
    monkeypatch.setenv("LRC_CHAIN_NO_GENERIC", "1")
    unfused = blocks.Chain(ctx, taps, decim, nfft)
    monkeypatch.delenv("LRC_CHAIN_NO_GENERIC")
    assert fused.kind == 2 and unfused.kind == 0
    a, b = fused.run(x, k), unfused.run(x, k)
    assert a.shape == b.shape == (frames // k, nfft)
    assert (a - b).abs().max().item() <= 1e-5 * b.pow(2).mean().sqrt().item()
    fused.close(); unfused.close()


def test_chain_default_policy_fuses_where_it_measured_faster(ctx):
    """prefer_fused (chain_generic.cuh): two or more CTAs per SM and not FP32-bound in the producers"""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    for (ntaps, decim, nfft), kind in {(64, 8, 1024): 2, (128, 10, 1024): 2, (64, 16, 512): 2, (64, 4, 512): 2, (64, 10, 512): 2,
                                       (64, 5, 1024): 2, (64, 4, 1024): 2, (128, 4, 1024): 0, (64, 5, 2048): 0, (64, 10, 2048): 0,
                                       (64, 16, 2048): 0, (128, 8, 2048): 0, (64, 16, 1024): 0}.items():
        ch = blocks.Chain(ctx, np.resize(taps, ntaps), decim, nfft)
        assert ch.kind == kind, (ntaps, decim, nfft, ch.kind)
        ch.close()


def test_chain_shapes_without_an_instance_stay_unfused(ctx):
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    for ntaps, decim, nfft in [(64, 3, 1024), (64, 10, 256), (129, 10, 1024), (64, 10, 4096)]:
        ch = blocks.Chain(ctx, np.resize(taps, ntaps), decim, nfft)
        assert ch.kind == 0
        ch.close()
    ch = blocks.Chain(ctx, taps, 10, 1024)
    assert ch.kind == 1
    ch.close()


@pytest.mark.parametrize("frames,k", [(100, 10), (900, 900), (1000, 64), (37, 1)])
def test_chain_host_ring_matches_device_path(ctx, frames, k):
    """the HOST-buffer entry point (pinned ring, chunked H2D overlapped with the kernel) must give the
    same rows as the device-resident call."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    n = frames * 10240 + 54
    xh = torch.from_numpy(synth.cf32_noise_tones(n, seed=k)).pin_memory()
    ch = blocks.Chain(ctx, taps, 10, 1024)
    a = ch.run(xh.to(ctx.tdev), k).cpu()
    b = ch.run_host(xh, k)
    assert a.shape == b.shape == (frames // k, 1024)
    # rows that span several ring segments are summed in a different order: float rounding only
    assert (a - b).abs().max().item() <= 1e-5 * a.pow(2).mean().sqrt().item()
    ch.close()


def test_chain_host_ring_u8_input_matches_unpack_then_chain(ctx):
    """lrc_chain_run_host_u8: rtlsdr bytes in host memory -> (device) data_to_samples -> chain; equals the cf32
    host path fed with the oracle's unpack of the same bytes, bit for bit (the unpack is bit-exact)."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    frames, k = 300, 20
    iq = synth.iq_tone_noise_u8(frames * 10240 + 54, seed=21)
    ch = blocks.Chain(ctx, taps, 10, 1024, capi.WINDOW_HANN)
    got = ch.run_host_u8(iq, k)
    ref = ch.run_host(oracle.data_to_samples(iq), k)
    assert got.shape == ref.shape == (frames // k, 1024)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    pref = D.psd_rows(oracle.fir_decimate(oracle.data_to_samples(iq), taps, 10), 1024, k, D.hann_periodic(1024))
    assert_close_rms(got, pref)
    ch.close()


_TIGHT_HOST = r"""
import ctypes, mmap, sys
import os

import numpy as np
sys.path.insert(0, {root!r})
import oracle
from oracle import defined_f64 as D
from libredio_b200 import blocks, synth
ntaps, decim, nfft, frames, k = {ntaps}, {decim}, {nfft}, {frames}, {k}
n = (frames - 1) * nfft * decim + (nfft - 1) * decim + ntaps          # the exact minimum for `frames` frames
page = mmap.PAGESIZE
nbytes = n * 8
npages = (nbytes + page - 1) // page
mm = mmap.mmap(-1, (npages + 1) * page)
base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
libc = ctypes.CDLL(None, use_errno=True)
assert libc.mprotect(ctypes.c_void_p(base + npages * page), ctypes.c_size_t(page), 0) == 0     # PROT_NONE guard page
off = npages * page - nbytes                                           # the buffer ENDS at the guard page
x = np.frombuffer(mm, dtype=np.complex64, count=n, offset=off)
x[:] = synth.cf32_noise_tones(n, seed=9)
rng = np.random.default_rng(1)
taps = (rng.standard_normal(ntaps) / 3).astype(np.float32)
ctx = blocks.Context(0)
ch = blocks.Chain(ctx, taps, decim, nfft)
assert ch.frames(n) == frames and ch.frames(n - 1) == frames - 1
got = ch.run_host(x, k)
ref = D.psd_rows(oracle.fir_decimate(np.array(x), taps, decim), nfft, k, D.hann_periodic(nfft))
assert got.shape == ref.shape, (got.shape, ref.shape)
err = np.max(np.abs(got - ref)) / np.sqrt(np.mean(ref ** 2))
assert err <= 1e-4, err
print("tight-host ok", err)
"""


@pytest.mark.parametrize("ntaps,decim,nfft,frames,k", [(5, 16, 256, 900, 4), (64, 10, 1024, 410, 5), (3, 7, 64, 40, 8)])
def test_chain_host_ring_never_reads_past_the_callers_buffer(ntaps, decim, nfft, frames, k):
    """lrc_chain_frames promises frames from exactly (frames-1)*nfft*decim + (nfft-1)*decim + ntaps samples; with
    decim > ntaps that is LESS than frames*nfft*decim, and the host ring must not copy more (it once did).
    The caller's buffer ends at a PROT_NONE page; run in a child process so an over-read is a failure, not a crash
    of the test session."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _TIGHT_HOST.format(root=root, ntaps=ntaps, decim=decim, nfft=nfft, frames=frames, k=k)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "tight-host ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-1500:])


@pytest.mark.parametrize("frames,k", [(1, 1), (40, 8), (130, 64)])
def test_chain_u8_instance_expands_the_tile_on_chip_bit_exact(ctx, frames, k):
    """lrc_chain_run_u8: rtlsdr bytes on the DEVICE; the fused kernel's producers run i2f on the TMA-loaded raw tile in shared
    memory (HBM carries 2 B per sample, no unpack launch) and then the same FIR code, so the rows equal the cf32 chain fed
    with the bit-exact unpack of the same bytes, bit for bit -- and the oracle within 1e-4 x RMS."""
    from libredio_b200 import blocks
    taps = synth.lpf_taps(64, 0.04)
    n = frames * 10240 + 54
    iq = synth.iq_tone_noise_u8(n, seed=100 + frames)
    ch = blocks.Chain(ctx, taps, 10, 1024, capi.WINDOW_HANN)
    got = ch.run_u8(dev(iq, ctx), k).cpu().numpy()
    x = oracle.data_to_samples(iq)
    ref = ch.run(dev(x, ctx), k).cpu().numpy()
    assert got.shape == ref.shape == (frames // k, 1024)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert_close_rms(got, chain_ref(x, taps, 10, 1024, k))
    ch.close()


def test_chain_u8_generic_shape_unpacks_then_runs_the_cf32_path(ctx):
    from libredio_b200 import blocks
    rng = np.random.default_rng(2)
    taps = (rng.standard_normal(33) / 6).astype(np.float32)
    frames, k, nfft, decim = 24, 4, 256, 4
    n = frames * nfft * decim + 33 - decim
    iq = synth.iq_tone_noise_u8(n, seed=8)
    ch = blocks.Chain(ctx, taps, decim, nfft)
    got = ch.run_u8(dev(iq, ctx), k).cpu().numpy()
    assert_close_rms(got, chain_ref(oracle.data_to_samples(iq), taps, decim, nfft, k))
    ch.close()
