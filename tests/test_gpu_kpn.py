"""GPU: the C++ kpn GPU blocks in a threaded graph, and the libkissfft / libsamplerate ABI shims driven by a
plain C program written against the reference's FFI declarations."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exe,ok", [("test_gpu_blocks", "kpn gpu OK"), ("test_shims", "shims OK")])
def test_cpp_programs(exe, ok):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kpn"), exe])
    out = subprocess.run([os.path.join(ROOT, "kpn", exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and ok in out.stdout, out.stdout + out.stderr


def _apps_exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kpn"), "test_gpu_apps"])
    return os.path.join(ROOT, "kpn", "test_gpu_apps")


def test_ook_application_graph_from_an_iq_capture_file(tmp_path):
    """the reference's one complete application (ratpak.rs:60-119) as a KPN graph on the GPU block, fed from the raw
    rtl_sdr capture format: iq_file_source_u8 -> batch -> kpn_gpu::ook_decode -> split_protocols -> fork -> binconv x 2 (36-bit
    packets) / applicator (24-bit packets).  What the sinks receive must equal kpn::eat over (a) the bits the capture was built
    to carry and (b) the CPU oracle's packets."""
    import numpy as np
    import oracle
    from libredio_b200 import synth
    n_blocks = 600
    iq, sent = synth.ook_capture_u8(n_blocks, seed=77, n_packets=3)          # protocols B, B, A
    ref = oracle.ook_decode(iq)
    layouts = {"A": (4, 8, 4, 12, 8), "C": (4, 8, 2, 10, 12)}     # ratpak.rs:115,119: BOTH applied to the 36-bit packets (fork :98)

    def eat(bits, widths):                                        # kpn.rs:111-124
        out, i = [], 0
        for w in widths:
            out.append(int("".join(str(int(b)) for b in bits[i:i + w]), 2))
            i += w
        return out
    a_ref, b_ref = [list(p) for p in ref["a_packets"]], [list(p) for p in ref["b_packets"]]
    a_sent, b_sent = [list(b) for pr, b in sent if pr == 0], [list(b) for pr, b in sent if pr == 1]
    assert [[int(v) for v in p] for p in a_ref] == [[int(v) for v in p] for p in a_sent] and len(a_ref) >= 1
    assert [[int(v) for v in p] for p in b_ref] == [[int(v) for v in p] for p in b_sent] and len(b_ref) >= 1
    lines = []
    for tag, widths in layouts.items():
        lines += [tag + " " + " ".join(str(v) for v in eat(p, widths)) for p in a_ref]
    lines += ["B " + " ".join(str(int(v)) for v in p) + " 0" for p in b_ref]      # applicator(|x| {x.push(0); x}) ratpak.rs:120-123
    assert lines
    cap = tmp_path / "capture.iq"
    iq.tofile(cap)
    exp = tmp_path / "expected.txt"
    exp.write_text("\n".join(lines) + "\n")
    out = subprocess.run([_apps_exe(), "ook", str(cap), str(n_blocks), str(exp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "kpn app ook OK" in out.stdout, out.stdout + out.stderr


def test_psd_application_graph_from_a_wav_file(tmp_path):
    """wavio's complex source (2-channel WAV, wavio.rs:30-46, as Vec chunks) -> kpn_gpu::chain_psd: every chunk's rows against
    the oracle FIR + f64 Hann PSD of the same chunk."""
    import struct
    import numpy as np
    import oracle
    from oracle import defined_f64 as D
    from libredio_b200 import synth
    rate, k, rows_per_chunk = 2_400_000, 4, 3
    chunk = rows_per_chunk * k * 10240 + 54
    n = 2 * chunk + chunk // 2                                   # two full chunks and a short one (1 row)
    x = (synth.cf32_noise_tones(n, seed=12) * 0.1).astype(np.complex64)
    wav = tmp_path / "capture.wav"
    data = x.view(np.float32).tobytes()
    with open(wav, "wb") as f:                                   # IEEE-float stereo WAV: I = left, Q = right
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, 2, rate, rate * 8, 8, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
    taps = synth.lpf_taps(64, 0.04)
    rows = []
    for lo in range(0, n, chunk):
        seg = x[lo: lo + chunk]
        z = oracle.fir_decimate(seg, taps, 10)
        r = D.psd_rows(z, 1024, k, D.hann_periodic(1024))
        if r.size:
            rows.append(r.astype(np.float32))
    exp = tmp_path / "expected.f32"
    np.concatenate(rows).astype(np.float32).tofile(exp)
    taps.astype(np.float32).tofile(str(exp) + ".taps")
    out = subprocess.run([_apps_exe(), "psd", str(wav), str(rate), str(chunk), str(k), str(exp)], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "kpn app psd OK" in out.stdout, out.stdout + out.stderr
