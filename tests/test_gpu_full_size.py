"""GPU parity, part 4: BASELINE.json's configurations at their FULL sizes.

The CPU oracle cannot process these whole in seconds, so each test checks (a) size-independent properties of the
full result (Parseval energy per row, shard-equals-whole, seam windows) and (b) the oracle itself on randomly
chosen pieces of the full run (rows, streams, output windows) copied back from the device.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import defined_f64 as D
from libredio_b200 import capi, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rms(a):
    return float(np.sqrt(np.mean(np.abs(np.asarray(a, dtype=np.complex128)) ** 2)))


def device_signal(ctx, n, seed):
    g = torch.Generator(device=ctx.tdev).manual_seed(seed)
    x = torch.view_as_complex(torch.randn(n, 2, device=ctx.tdev, generator=g))
    k = torch.arange(min(n, 1 << 22), device=ctx.tdev, dtype=torch.float32)
    tone = torch.polar(torch.full_like(k, 2.0), (2 * np.pi * 0.0112) * k)       # one strong line, first 4 M samples
    x[: tone.numel()] += tone
    return x


def test_config2_psd_2pow26_samples_parseval_and_spot_rows(ctx):
    """configs[1]: 2^26 cf32 samples = 65536 frames of 1024, Hann, rows of K = 64 plus the full average."""
    from libredio_b200 import blocks
    nfft, k, frames = 1024, 64, 65536
    x = device_signal(ctx, frames * nfft, seed=2)
    psd = blocks.Psd(ctx, nfft, capi.WINDOW_HANN)
    rows = psd.run(x, k)
    assert rows.shape == (frames // k, nfft)
    # Parseval per row: sum_b P[b] = (N / K) sum_{frames, n} |w[n] x[n]|^2   (independent device reduction in f64)
    w = torch.from_numpy(synth.hann_periodic(nfft)).to(ctx.tdev).double()
    e = (torch.view_as_real(x).double().pow(2).sum(-1).view(frames, nfft) * w * w).sum(-1).view(-1, k).sum(-1) * (nfft / k)
    got = rows.double().sum(-1)
    assert float(((got - e).abs() / e).max()) < 2e-5
    # the oracle on randomly chosen rows of the full run
    rng = np.random.default_rng(26)
    xv = x.view(frames // k, k * nfft)
    for r in [0, frames // k - 1] + [int(v) for v in rng.integers(0, frames // k, 4)]:
        ref = D.psd_rows(xv[r].cpu().numpy(), nfft, k, D.hann_periodic(nfft))[0]
        assert np.max(np.abs(rows[r].cpu().numpy() - ref)) <= TOL * rms(ref), r
    # the full average (K = 65536) equals the mean of the K = 64 rows
    full = psd.run(x, frames)
    assert full.shape == (1, nfft)
    m = rows.double().mean(0)
    assert float(((full[0].double() - m).abs()).max()) <= TOL * float(m.pow(2).mean().sqrt())
    psd.close()


def test_headline_chain_full_size_spot_rows_and_shards(ctx):
    """the bench.py workload: 671 088 694 cf32 samples -> FIR64/10 -> 65536 frames -> Hann FFT -> K = 64 rows."""
    from libredio_b200 import blocks, shard
    taps = synth.lpf_taps(64, 0.04)
    frames, k, nfft, d = 65536, 64, 1024, 10
    n_in = frames * nfft * d + 64 - d
    x = device_signal(ctx, n_in, seed=7)
    ch = blocks.Chain(ctx, taps, d, nfft, capi.WINDOW_HANN)
    rows = ch.run(x, k)
    assert rows.shape == (frames // k, nfft)
    rng = np.random.default_rng(1)
    for r in [0, frames // k - 1] + [int(v) for v in rng.integers(0, frames // k, 3)]:
        lo = r * k * nfft * d
        seg = x[lo: lo + k * nfft * d + 64 - d].cpu().numpy()
        ref = D.psd_rows(oracle.fir_decimate(seg, taps, d), nfft, k, D.hann_periodic(nfft))[0]
        assert np.max(np.abs(rows[r].cpu().numpy() - ref)) <= TOL * rms(ref), r
    # row-aligned shards, each reading its own halo from the source, reproduce the whole run bit for bit
    parts = []
    for rank in range(8):
        sh = shard.chain_shard(n_in, 64, d, nfft, k, rank, 8)
        parts.append(ch.run(x[sh.in_start: sh.in_start + sh.in_len], k))
        assert parts[-1].shape[0] == sh.row_hi - sh.row_lo
    assert torch.equal(torch.cat(parts), rows)
    ch.close()


def test_config4_ook_4096_streams_bit_exact(ctx):
    """configs[3]: 4096 streams in one batch; every stream equals the oracle's decode of its capture."""
    from libredio_b200 import blocks
    n_streams, n_blocks, distinct = 4096, 450, 64
    caps, sent = [], []
    for s in range(distinct):
        iq, snt = synth.ook_capture_u8(n_blocks, seed=400 + s, n_packets=2)
        caps.append(iq); sent.append(snt)
    perm = np.random.default_rng(4).permutation(n_streams) % distinct        # which capture each stream carries
    iq = torch.from_numpy(np.stack(caps)).to(ctx.tdev)[torch.from_numpy(perm).to(ctx.tdev)].contiguous()
    ook = blocks.Ook(ctx, n_streams, n_blocks, 256000, 4096, 64)
    ook.decode(iq)
    pk = ook.packets()
    dbg = ook.debug()
    refs = [oracle.ook_decode(c) for c in caps]
    by_stream = {}
    for p in pk:
        by_stream.setdefault((p[0], p[1]), []).append(p[3])
    n_sent = 0
    for s in range(n_streams):
        r = refs[perm[s]]
        assert np.array_equal(dbg["block_sums"][s].view(np.uint32), r["block_sums"].view(np.uint32)), s
        assert dbg["n_bits"][s] == r["bits"].size and int(dbg["n_runs"][s]) == r["run_val"].size, s
        a, b = by_stream.get((s, 0), []), by_stream.get((s, 1), [])
        assert len(a) == len(r["a_packets"]) and len(b) == len(r["b_packets"]), s
        for got, ref in zip(a + b, list(r["a_packets"]) + list(r["b_packets"])):
            assert np.array_equal(got, ref), s
        n_sent += len(a) + len(b)
    assert n_sent == len(pk) and n_sent >= n_streams            # packets known by construction are all there
    ook.close()


def test_config5_fastfir_2pow30_stream_seam_windows(ctx):
    """configs[4]: 4096 taps over a 2^30-sample stream (8 GiB in, 8 GiB out), chunk-sharded at 2/4/8 GPUs: random
    output windows and every shard seam against an f64 direct convolution; shard outputs equal the whole run.
    Shards are cut on the 16384-point blocks the kernel computes in (explicit nfft = 16384 plan), so a shard repeats the
    whole run's arithmetic exactly; the automatic plan (kiss_fastfir's output length) gives the same flushed stream."""
    from libredio_b200 import blocks, shard
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * (1 << 30):
        pytest.skip("needs 30 GiB of free device memory")
    nh, n, nfft = 4096, 1 << 30, 16384
    ngood = nfft - nh + 1
    rng = np.random.default_rng(6)
    h = ((rng.standard_normal(nh) + 1j * rng.standard_normal(nh)) / 64).astype(np.complex64)
    x = torch.empty(n, dtype=torch.complex64, device=ctx.tdev)
    g = torch.Generator(device=ctx.tdev).manual_seed(5)
    for lo in range(0, n, 1 << 27):                              # generated in pieces: randn's temporaries stay small
        x[lo: lo + (1 << 27)] = torch.view_as_complex(torch.randn(1 << 27, 2, device=ctx.tdev, generator=g))
    ff = blocks.FastFir(ctx, h, nfft)
    assert ff.nfft == nfft and ff.ngood == ngood
    y = ff.run(x, flush=True)
    assert y.numel() == n - nh + 1 == ff.out_len(n, True)
    hr = h[::-1].astype(np.complex128)
    seams = []
    for world in (2, 4, 8):
        for rank in range(1, world):
            seams.append(shard.fastfir_shard(n, nh, nfft, rank, world).out_start)
    starts = [0, ngood - 8, y.numel() - 64] + [s - 32 for s in seams] + [int(v) for v in rng.integers(0, y.numel() - 64, 16)]
    for s0 in starts:
        seg = x[s0: s0 + 64 + nh - 1].cpu().numpy().astype(np.complex128)
        ref = np.array([np.dot(seg[k:k + nh], hr) for k in range(64)])
        got = y[s0:s0 + 64].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= TOL * rms(ref), s0
    # block-aligned shards (each reading its own nh-1 halo from the source) reproduce the whole run bit for bit
    for rank in (0, 3, 7):
        sh = shard.fastfir_shard(n, nh, nfft, rank, 8)
        part = ff.run(x[sh.in_start: sh.in_start + sh.in_len])
        assert part.numel() == sh.out_len
        assert torch.equal(torch.view_as_real(part), torch.view_as_real(y[sh.out_start: sh.out_start + sh.out_len]))
    ff.close()
    # the drop-in plan: kiss_fastfir's own block size (8192) fixes the un-flushed length, the values are the same stream
    auto = blocks.FastFir(ctx, h, 0)
    assert auto.nfft == 8192 and auto.out_len(n) == ((n - 8192) // 4097 + 1) * 4097
    ya = auto.run(x)
    assert ya.numel() == auto.out_len(n)
    assert torch.equal(torch.view_as_real(ya), torch.view_as_real(y[: ya.numel()]))
    auto.close()


def test_config3_fm_receiver_1024_channels_full_size(ctx):
    """configs[2] at BASELINE size: 1024 channels x 2.4 Msps x 0.1 s of rtlsdr u8 IQ through the fused receiver (one
    kernel), fed in two pushes; every channel equals the three stand-alone stages bit for bit, and randomly chosen
    channels are held to the stage oracles (FIR 1e-4 x RMS, discriminator and resampler >= 100 dB on their inputs)."""
    from libredio_b200 import blocks
    n_ch, n, distinct = 1024, 240_000, 16
    taps = synth.lpf_taps(64, 0.04)
    base = np.stack([synth.fm_iq_u8(n, seed=3 + c) for c in range(distinct)])
    perm = np.random.default_rng(3).permutation(n_ch) % distinct
    iq = torch.from_numpy(base).to(ctx.tdev)[torch.from_numpy(perm).to(ctx.tdev)].contiguous()
    rx = blocks.FmReceiver(ctx, taps, 10, 0.2, n_ch, n)
    assert rx.fused
    cut = 96_000
    audio = torch.cat([rx.push(iq[:, : 2 * cut].contiguous()), rx.push(iq[:, 2 * cut:].contiguous())], dim=1)
    rx.close()
    n_bb = (n - 64) // 10 + 1
    assert audio.shape == (n_ch, (n_bb - 1) // 5 + 1)
    fir = blocks.Fir(ctx, taps, 10)
    bb = fir.run_u8(iq)
    d = blocks.fm_demod(ctx, bb)
    rs = blocks.Resampler(ctx, 0.2, n_ch, d.shape[1])
    ref = rs.process(d)
    assert torch.equal(audio, ref)
    for c in [0, n_ch - 1] + [int(v) for v in np.random.default_rng(33).integers(0, n_ch, 3)]:
        ref_bb = oracle.fir_decimate(oracle.data_to_samples(base[perm[c]]), taps, 10)
        bb_c, d_c = bb[c].cpu().numpy(), d[c].cpu().numpy()
        assert np.max(np.abs(bb_c - ref_bb)) <= TOL * rms(ref_bb), c
        assert D.snr_db(D.fm_discriminator(bb_c), d_c) >= 100.0, c
        assert D.snr_db(D.resample(d_c, 0.2), audio[c].cpu().numpy()) >= 100.0, c
    fir.close(); rs.close()
