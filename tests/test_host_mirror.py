"""The C++ mirror of the reference's API (rust-modem_b200/host/modem.hpp): the reference's own
unit tests restated on it (CPU), and streaming DigitalModulator -> Demodulator round trips plus
the `src/bin`-style loopback caller through the CUDA library (GPU)."""
import os
import subprocess

import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "rust-modem_b200", "host")


@pytest.fixture(scope="module")
def host_bins(pkg):
    subprocess.check_call(["make", "-C", HOST, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return os.path.join(HOST, "bin")


def test_reference_unit_tests_on_the_mirror(host_bins):
    r = subprocess.run([os.path.join(host_bins, "host_tests"), "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok (0 failures)" in r.stdout


@pytest.mark.gpu
def test_streaming_api_roundtrip_on_gpu(host_bins):
    r = subprocess.run([os.path.join(host_bins, "host_tests"), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_loopback_binary(host_bins):
    """BASELINE config 0: random 1 Mbit payload, reference default rates, bit-exact round trip."""
    r = subprocess.run([os.path.join(host_bins, "loopback"), "-n", str(1 << 20)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 bit errors" in r.stdout.splitlines()[0] and "0 bit errors" in r.stdout.splitlines()[1]


def test_loopback_binary_panics_without_gpu(host_bins, pkg):
    import ctypes as C
    n = C.c_int(0)
    if pkg.lib().modem_gpu_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([os.path.join(host_bins, "loopback"), "-n", "4096"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 101 and "panicked" in r.stderr  # no CPU fallback
