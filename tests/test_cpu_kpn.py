"""CPU suite, part 3: the C++ kpn block/port contract (kpn/kpn.hpp) against the semantics of
src/kpn/src/kpn.rs -- built with g++ and run as a subprocess."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kpn_contract_and_cpu_blocks():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kpn"), "test_kpn_cpu"])
    out = subprocess.run([os.path.join(ROOT, "kpn", "test_kpn_cpu")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "kpn cpu OK" in out.stdout, out.stdout + out.stderr


def test_gpu_blocks_and_shims_compile_and_link_against_the_c_abi_only():
    """no compute here: just that the C++ GPU blocks and the two ABI shims build and link (they must not
    reference the oracle)"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kpn"), "test_gpu_blocks", "test_shims"])
    for exe in ("test_gpu_blocks", "test_shims"):
        ldd = subprocess.run(["ldd", os.path.join(ROOT, "kpn", exe)], capture_output=True, text=True).stdout
        assert "liboracle" not in ldd
    ldd = subprocess.run(["ldd", os.path.join(ROOT, "kpn", "test_shims")], capture_output=True, text=True).stdout
    assert "libkissfft.so" in ldd and "libsamplerate.so" in ldd
