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
