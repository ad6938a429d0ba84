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


def _rust_display(x):
    """Rust's `{}` for an f32: shortest round-trip digits, positional."""
    import numpy as np
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "-inf" if x < 0 else "inf"
    if x == 0:
        return "0"
    return np.format_float_positional(x, unique=True, trim="-")


@pytest.mark.gpu
@pytest.mark.parametrize("dmod,extra", [("qpsk", []), ("qpsk", ["-p", "20"]), ("qam16", ["-b", "1250"]), ("oqpsk", ["-b", "1000"]),
                                        ("msk", ["-b", "1000", "-p", "3"]), ("bfsk", []), ("dqpsk", ["-c", "500"]),
                                        ("16cpfsk", []), ("mfsk", ["-r", "8000", "-b", "200"])])
def test_modulate_binary_matches_oracle(host_bins, orc, dmod, extra):
    """src/bin/modulate.rs drop-in: ASCII bits on stdin -> f32-LE waveform on stdout, byte for byte the oracle's stream."""
    import numpy as np
    opt = dict(zip(extra[::2], extra[1::2]))
    sr, br, cf, pc = int(opt.get("-r", 10000)), int(opt.get("-b", 220)), int(opt.get("-c", 1000)), int(opt.get("-p", 0))
    o = orc.OraclePath(dmod, br, sr, cf)
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2, 97 * o.bps + 1, dtype=np.uint8)  # a partial trailing symbol is dropped (data.rs:164-174)
    text = " ".join("".join(map(str, bits[i:i + 7])) for i in range(0, len(bits), 7)) + "\n"
    r = subprocess.run([os.path.join(host_bins, "modulate"), "-m", dmod] + extra, input=text.encode(), capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(r.stdout, dtype="<f4")
    P = sr // cf * pc - 1 if pc else 0
    ref = o.modulate_real(bits[None, : len(bits) // o.bps * o.bps], preamble=P)[0]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_modulate_binary_iq(host_bins, orc):
    import numpy as np
    o = orc.OraclePath("16apsk", 220, 10000, 1000)
    bits = np.random.default_rng(6).integers(0, 2, 4 * 50, dtype=np.uint8)
    r = subprocess.run([os.path.join(host_bins, "modulate"), "-m", "16apsk", "--iq"], input="".join(map(str, bits)).encode(),
                       capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(r.stdout, dtype="<f4").reshape(-1, 2)
    _, iq = o.modulate(bits[None, :], want_iq=True)
    assert np.array_equal(got.view(np.uint32), iq[0].view(np.uint32))


@pytest.mark.gpu
def test_demodulate_binary_matches_oracle(host_bins, orc):
    """src/bin/demodulate.rs drop-in: native-endian i16 on stdin -> `i:{}\tq:{}` lines, Rust float formatting; carrier
    900 Hz / 10 kHz as in the reference."""
    import numpy as np
    bits = np.random.default_rng(7).integers(0, 2, (1, 2 * 60), dtype=np.uint8)
    tx = orc.OraclePath("qpsk", 220, 10000, 900).modulate_real(bits, preamble=99)
    wire = np.round(tx * 12000.0).astype(np.int16)
    r = subprocess.run([os.path.join(host_bins, "demodulate")], input=wire.tobytes() + b"\x01", capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    lines = r.stdout.decode().splitlines()
    po, filt, _, _ = orc.OraclePath("qpsk", 220, 10000, 900).demodulate_real(wire, lock=64)
    assert len(lines) == filt.shape[1] == wire.shape[1] - 64
    want = [f"i:{_rust_display(i)}\tq:{_rust_display(q)}" for i, q in filt[0]]
    assert lines == want


def test_binaries_panic_like_the_reference(host_bins, pkg):
    """Argument handling happens before any device work: same panics / exit code 101 as the Rust binaries."""
    mod = os.path.join(host_bins, "modulate")
    for args, msg in [([], "digital modulation is required"), (["-m", "qpsk", "-c", "6000"], "cf < sr / 2"),
                      (["-m", "qpsk", "-c", "900", "-p", "3"], "sr % cf == 0"), (["-m", "qpsk", "-b", "x"], "invalid baud rate")]:
        r = subprocess.run([mod] + args, input=b"01", capture_output=True, timeout=60)
        assert r.returncode == 101 and msg in r.stderr.decode(), (args, r.stderr)
    r = subprocess.run([mod, "-h"], capture_output=True, timeout=60)
    assert r.returncode == 0 and b"Modulate the bits on stdin to a waveform on stdout" in r.stdout
    import ctypes as C
    n = C.c_int(0)
    if not (pkg.lib().modem_gpu_device_count(C.byref(n)) == 0 and n.value > 0):
        r = subprocess.run([mod, "-m", "nope"], input=b"01", capture_output=True, timeout=60)
        assert r.returncode == 101 and "invalid digital modulation" in r.stderr.decode()
        r = subprocess.run([mod, "-m", "qpsk"], input=b"0110", capture_output=True, timeout=60)
        assert r.returncode == 101 and "no CPU fallback" in r.stderr.decode()  # no silent host arithmetic
        r = subprocess.run([os.path.join(host_bins, "demodulate")], input=b"\0" * 200, capture_output=True, timeout=60)
        assert r.returncode == 101


def test_loopback_binary_panics_without_gpu(host_bins, pkg):
    import ctypes as C
    n = C.c_int(0)
    if pkg.lib().modem_gpu_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([os.path.join(host_bins, "loopback"), "-n", "4096"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 101 and "panicked" in r.stderr  # no CPU fallback
