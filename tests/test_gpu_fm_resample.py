"""GPU parity, part 2: quadrature FM discriminator and rational polyphase resampler (north_star subsystem 4).

Tolerance (BASELINE.json): output SNR >= 100 dB against the f64 oracle (oracle/defined_f64.py).
The resampler oracle is OUR definition -- libsamplerate parity is unpinned.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import defined_f64 as D
from libredio_b200 import capi, synth

pytestmark = pytest.mark.gpu
MIN_SNR_DB = 100.0


def dev(a, ctx):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.tdev)


def fm_signal(n, seed):
    iq = synth.fm_iq_u8(n * 10 + 64, seed=seed)
    return oracle.fir_decimate(oracle.data_to_samples(iq), synth.lpf_taps(64, 0.04), 10)[:n]


@pytest.mark.parametrize("n", [1, 2, 1000, 60_001])
def test_fm_discriminator_snr(ctx, n):
    from libredio_b200 import blocks
    x = fm_signal(n, seed=3)
    got = blocks.fm_demod(ctx, dev(x, ctx)).cpu().numpy()
    ref = D.fm_discriminator(x)
    assert got.shape == ref.shape
    assert got[0] == 0.0                                     # x[-1] = 0 -> atan2(0, 0) = 0
    if n > 1:
        assert D.snr_db(ref, got) >= MIN_SNR_DB


def test_fm_discriminator_state_is_carried_across_chunks(ctx):
    from libredio_b200 import blocks
    n_ch, n = 4, 30_000
    x = np.stack([fm_signal(n, seed=3 + c) for c in range(n_ch)])
    whole = blocks.fm_demod(ctx, dev(x, ctx)).cpu().numpy()
    state = torch.zeros(n_ch, dtype=torch.complex64, device=ctx.tdev)
    parts, pos = [], 0
    for c in (1, 999, 14_000, 15_000):
        parts.append(blocks.fm_demod(ctx, dev(x[:, pos:pos + c], ctx), state).cpu().numpy())
        pos += c
    got = np.concatenate(parts, axis=1)
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    for c in range(n_ch):
        assert D.snr_db(D.fm_discriminator(x[c]), got[c]) >= MIN_SNR_DB


@pytest.mark.parametrize("ratio,n", [(0.2, 48_000), (0.5, 10_000), (2.0, 1000), (48000 / 44100, 8820), (3 / 7, 7001), (1.0, 500)])
def test_resampler_snr_vs_f64_definition(ctx, ratio, n):
    from libredio_b200 import blocks
    rng = np.random.default_rng(int(ratio * 1000) + n)
    t = np.arange(n)
    x = (np.sin(2 * np.pi * 0.01 * t) + 0.5 * np.sin(2 * np.pi * 0.037 * t + 1) + 0.1 * rng.standard_normal(n)).astype(np.float32)
    rs = blocks.Resampler(ctx, ratio, 1, n)
    L, M = D.resampler_ratio(ratio)
    assert (rs.L, rs.M) == (L, M)
    # the device-side filter design equals the oracle's definition
    h_ref = D.resampler_taps(L, M)
    assert rs.taps().shape == h_ref.shape and np.max(np.abs(rs.taps() - h_ref)) < 1e-12 * np.max(np.abs(h_ref)) + 1e-15
    got = rs.process(dev(x, ctx)).cpu().numpy()
    ref = D.resample(x, ratio)
    assert got.shape == ref.shape == ((n * L - 1) // M + 1,)
    assert got.size <= int(ratio * n + 1) + 1                # fits the reference's output sizing (samplerate.rs:64)
    assert D.snr_db(ref, got) >= MIN_SNR_DB
    rs.close()


def test_resampler_smoke_case_of_reference(ctx):
    """the only 'test' in the reference: resample a 1000-sample sine by 2.0 and print the length
    (samplerate.rs:89-96)."""
    from libredio_b200 import blocks
    v = np.sin(np.arange(1000, dtype=np.float32) / 1000.0).astype(np.float32)
    rs = blocks.Resampler(ctx, 2.0, 1, 1000)
    out = rs.process(dev(v, ctx))
    assert out.numel() == 2000
    rs.close()


def test_resampler_streaming_is_chunk_independent(ctx):
    from libredio_b200 import blocks
    n_ch, n, ratio = 3, 24_000, 0.2
    rng = np.random.default_rng(8)
    x = rng.standard_normal((n_ch, n)).astype(np.float32)
    rs = blocks.Resampler(ctx, ratio, n_ch, n)
    whole = rs.process(dev(x, ctx)).cpu().numpy()
    rs.reset()
    parts, pos = [], 0
    for c in (1, 4, 5, 990, 11_000, 12_000):
        parts.append(rs.process(dev(x[:, pos:pos + c], ctx)).cpu().numpy())
        pos += c
    got = np.concatenate(parts, axis=1)
    assert got.shape == whole.shape
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    for c in range(n_ch):
        assert D.snr_db(D.resample(x[c], ratio), got[c]) >= MIN_SNR_DB
    rs.close()


@pytest.mark.parametrize("den,n_ch,n", [(2, 1, 3001), (2, 5, 9000), (3, 2, 7001), (4, 3, 5000), (5, 1, 4000), (5, 5, 23_994),
                                        (6, 2, 12_345)])
def test_resampler_integer_decimators_many_channels(ctx, den, n_ch, n):
    """L = 1 decimators run the packed two-tile kernel (M = 2, 3, 5, 6) or the scalar tile kernel (M = 4): odd channel
    counts (a tile pair whose second half is absent or belongs to the next channel), ragged last tiles, the first tile
    of every channel starting inside the carried history; every channel against the f64 definition."""
    from libredio_b200 import blocks
    rng = np.random.default_rng(100 * den + n_ch)
    t = np.arange(n)
    x = np.stack([(np.sin(2 * np.pi * (0.003 + 0.002 * c) * t) + 0.2 * rng.standard_normal(n)).astype(np.float32)
                  for c in range(n_ch)])
    rs = blocks.Resampler(ctx, 1.0 / den, n_ch, n)
    assert (rs.L, rs.M) == (1, den)
    first = rs.process(dev(x[:, : n // 3], ctx)).cpu().numpy()          # two pushes: the second starts in the carry
    second = rs.process(dev(x[:, n // 3:], ctx)).cpu().numpy()
    got = np.concatenate([first, second], axis=1)
    assert got.shape == (n_ch, (n - 1) // den + 1)
    for c in range(n_ch):
        assert D.snr_db(D.resample(x[c], 1.0 / den), got[c]) >= MIN_SNR_DB
    rs.close()


def test_resampler_packed_kernel_is_bit_identical_to_scalar_kernel(tmp_path):
    """The packed (FFMA2, two tiles per CTA) decimator must produce the same bits as the scalar one-tile kernel it
    replaced: per output it is the same ascending-tap fma chain.  The kernel choice is read once per process
    (LRC_RS_VARIANT), so each variant runs in its own interpreter."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from libredio_b200 import blocks\n"
        "ctx = blocks.Context(0)\n"
        "rng = np.random.default_rng(11)\n"
        "outs = []\n"
        "for den, n_ch, n in ((5, 3, 20_011), (2, 2, 5000), (3, 1, 9001), (6, 4, 7000)):\n"
        "    x = rng.standard_normal((n_ch, n)).astype(np.float32)\n"
        "    rs = blocks.Resampler(ctx, 1.0 / den, n_ch, n)\n"
        "    a = rs.process(torch.from_numpy(np.ascontiguousarray(x[:, :777])).to(ctx.tdev)).cpu().numpy()\n"
        "    b = rs.process(torch.from_numpy(np.ascontiguousarray(x[:, 777:])).to(ctx.tdev)).cpu().numpy()\n"
        "    outs.append(np.concatenate([a, b], axis=1).ravel())\n"
        "    rs.close()\n"
        "np.save(sys.argv[1], np.concatenate(outs))\n" % root)
    res = {}
    for variant in ("0", "1"):
        path = str(tmp_path / f"rs_variant_{variant}.npy")
        env = dict(os.environ, LRC_RS_VARIANT=variant)
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
        res[variant] = np.load(path)
    assert res["0"].shape == res["1"].shape and res["0"].size > 10_000
    assert np.array_equal(res["0"].view(np.uint32), res["1"].view(np.uint32))


def test_resampler_unsupported_ratio_is_loud(ctx):
    from libredio_b200 import blocks
    with pytest.raises(capi.LrcError) as e:
        blocks.Resampler(ctx, np.pi / 3, 1, 100)
    assert e.value.status == capi.ERR_UNSUPPORTED


def test_fm_broadcast_chain_config3(ctx):
    """config 3 end to end on a few channels: u8 IQ 2.4 Msps -> fused unpack+FIR64/10 -> discriminator ->
    resample 1/5 -> 48 kHz; every stage against its oracle on the previous stage's GPU output."""
    from libredio_b200 import blocks
    n_ch, n = 4, 240_000
    taps = synth.lpf_taps(64, 0.04)
    iq = np.stack([synth.fm_iq_u8(n, seed=3 + c) for c in range(n_ch)])
    fir = blocks.Fir(ctx, taps, 10)
    bb = fir.run_u8(dev(iq, ctx))
    d = blocks.fm_demod(ctx, bb)
    rs = blocks.Resampler(ctx, 0.2, n_ch, d.shape[1])
    audio = rs.process(d).cpu().numpy()
    bb_h, d_h = bb.cpu().numpy(), d.cpu().numpy()
    for c in range(n_ch):
        ref_bb = oracle.fir_decimate(oracle.data_to_samples(iq[c]), taps, 10)
        assert np.max(np.abs(bb_h[c] - ref_bb)) <= 1e-4 * np.sqrt(np.mean(np.abs(ref_bb) ** 2))
        assert D.snr_db(D.fm_discriminator(bb_h[c]), d_h[c]) >= MIN_SNR_DB
        assert D.snr_db(D.resample(d_h[c], 0.2), audio[c]) >= MIN_SNR_DB
    assert audio.shape[1] == (d.shape[1] - 1) // 5 + 1
    # the 1 kHz programme tone must dominate the demodulated audio spectrum
    spec = np.abs(np.fft.rfft(audio[0][2000:] * np.hanning(audio[0][2000:].size)))
    f = np.fft.rfftfreq(audio[0][2000:].size, 1 / 48000.0)
    assert abs(f[np.argmax(spec[5:]) + 5] - 1000.0) < 30.0
    fir.close(); rs.close()


def staged_pipeline(ctx, iq_dev, taps, n_ch):
    """the three stand-alone stages one after the other (each has its own oracle test above)"""
    from libredio_b200 import blocks
    fir = blocks.Fir(ctx, taps, 10)
    d = blocks.fm_demod(ctx, fir.run_u8(iq_dev))
    rs = blocks.Resampler(ctx, 0.2, n_ch, d.shape[1])
    ref = rs.process(d)
    fir.close(); rs.close()
    return ref


@pytest.mark.parametrize("n_ch,chunks", [(3, (2, 126, 20_000, 39_872, 60_000)),          # chunk lengths not multiples of 8: copy path
                                         (4, (8, 4000, 56_000, 59_992)),               # multiples of 8: TMA path
                                         (1, (120_000,)), (2, (30, 30, 30, 119_910)), (5, (65_536, 54_464))])
def test_fm_receiver_streaming_is_chunk_independent_and_matches_stage_oracles(ctx, n_ch, chunks):
    """blocks.FmReceiver (config 3 as one streaming object, ONE kernel per push): chunked == whole bit for bit, and the
    whole run equals the three stand-alone stages run one after the other bit for bit (the fused kernel performs the same
    operation sequence per sample), each of which is held to its oracle above."""
    from libredio_b200 import blocks
    n = sum(chunks)
    taps = synth.lpf_taps(64, 0.04)
    iq = np.stack([synth.fm_iq_u8(n, seed=30 + c) for c in range(n_ch)])
    rx = blocks.FmReceiver(ctx, taps, 10, 0.2, n_ch, n)
    assert rx.fused
    whole = rx.push(dev(iq, ctx)).cpu().numpy()
    rx.reset()
    parts, pos = [], 0
    for c in chunks:
        parts.append(rx.push(dev(iq[:, 2 * pos: 2 * (pos + c)], ctx)).cpu().numpy())
        pos += c
    rx.close()
    got = np.concatenate(parts, axis=1)
    assert got.shape == whole.shape == (n_ch, ((n - 64) // 10 + 1 - 1) // 5 + 1)
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    ref = staged_pipeline(ctx, dev(iq, ctx), taps, n_ch).cpu().numpy()
    assert np.array_equal(ref.view(np.uint32), whole.view(np.uint32))
    # and directly against the f64 definitions on the oracle's baseband (FIR 1e-4 x RMS feeds >= 100 dB stages)
    bb = oracle.fir_decimate(oracle.data_to_samples(iq[0]), taps, 10)
    want = D.resample(D.fm_discriminator(bb), 0.2)
    assert D.snr_db(want[400:], whole[0][400:]) >= 80.0          # end to end: FIR rounding propagates through atan2


def test_fm_receiver_other_shapes_run_the_three_stages_behind_the_same_interface(ctx):
    from libredio_b200 import blocks
    rng = np.random.default_rng(5)
    n_ch, n = 2, 50_000
    taps = (rng.standard_normal(33) / 6).astype(np.float32)
    iq = np.stack([synth.fm_iq_u8(n, seed=70 + c) for c in range(n_ch)])
    rx = blocks.FmReceiver(ctx, taps, 8, 0.25, n_ch, n)
    assert not rx.fused
    whole = rx.push(dev(iq, ctx)).cpu().numpy()
    rx.reset()
    parts, pos = [], 0
    for c in (1000, 24_000, 25_000):
        parts.append(rx.push(dev(iq[:, 2 * pos: 2 * (pos + c)], ctx)).cpu().numpy())
        pos += c
    rx.close()
    got = np.concatenate(parts, axis=1)
    assert got.shape == whole.shape and whole.shape[1] > 1000
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    fir = blocks.Fir(ctx, taps, 8)
    d = blocks.fm_demod(ctx, fir.run_u8(dev(iq, ctx)))
    rs = blocks.Resampler(ctx, 0.25, n_ch, d.shape[1])
    ref = rs.process(d).cpu().numpy()
    assert np.array_equal(ref.view(np.uint32), whole.view(np.uint32))
    fir.close(); rs.close()
