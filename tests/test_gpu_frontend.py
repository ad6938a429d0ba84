"""GPU parity for the rows either side of the memoryless hot path, through the C ABI
against the CPU oracle: stateful mappers (bfsk, mfsk, 16cpfsk, msk, dqpsk, dbpsk), the sync tone and real-valued
wire output of src/bin/modulate.rs, and the Hilbert + PLL lock + Demodulator path of src/bin/demodulate.rs.
Every buffer is expected BIT-identical (same bar as tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from test_gpu_parity import assert_buffers, rand_bits

pytestmark = pytest.mark.gpu

STATEFUL = ["bfsk", "mfsk", "16cpfsk", "msk", "dqpsk", "dbpsk"]


def rates(sps):
    return {8: (1250, 10000), 45: (220, 10000), 4: (2500, 10000), 10: (1000, 10000)}[sps]


# ----------------------------------------------------------------------------- stateful TX
@pytest.mark.parametrize("scheme", STATEFUL)
@pytest.mark.parametrize("sps", [8, 10])
def test_stateful_tx(pkg, orc, scheme, sps):
    br, sr = rates(sps)
    kw = dict(scheme=scheme, baud_rate=br, sample_rate=sr, carrier_hz=1000)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(21, 5, 301 * o.bps + (o.bps - 1))  # ragged tail
    tx, iq = m.modulate(bits, want_iq=True)
    tx_ref, iq_ref = o.modulate(bits, want_iq=True)
    assert_buffers(iq, iq_ref, f"{scheme} sps={sps} baseband iq")
    assert_buffers(tx, tx_ref, f"{scheme} sps={sps} tx")


@pytest.mark.parametrize("scheme", ["mfsk", "dqpsk", "bfsk"])
def test_stateful_tx_long_frames(pkg, orc, scheme):
    """Long recurrences (the f32 phase accumulates rounding symbol by symbol, dmpsk.rs:30-32) and a sample
    counter that starts late (Carrier.sample carried over from a preamble)."""
    br, sr = rates(8)
    kw = dict(scheme=scheme, baud_rate=br, sample_rate=sr, carrier_hz=2500, sample0=999)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(22, 130, 4100 * o.bps)  # > 128 frames: more than one scan CTA; nsym not a multiple of 4 rows
    assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} long")
    bits = rand_bits(23, 3, 4099 * o.bps)
    assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} odd symbol count")


def test_stateful_reference_default_rates(pkg, orc):
    """The binaries' own rates: sr 10000, baud 220 -> sps 45 (odd: msk must refuse, msk.rs:13)."""
    for scheme in ["bfsk", "mfsk", "16cpfsk", "dqpsk"]:
        kw = dict(scheme=scheme, baud_rate=220, sample_rate=10000, carrier_hz=1000)
        m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
        bits = rand_bits(24, 2, 77 * o.bps)
        assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} sps 45")
    with pytest.raises(pkg.ModemError):
        pkg.Modem(scheme="msk", baud_rate=220, sample_rate=10000, carrier_hz=1000)


# ----------------------------------------------------------------------------- modulate binary output
@pytest.mark.parametrize("scheme", ["qpsk", "qam16", "oqpsk", "dcqpsk", "bfsk", "msk", "dqpsk", "16cpfsk"])
@pytest.mark.parametrize("preamble", [0, 29])
def test_modulate_real(pkg, orc, scheme, preamble):
    """modulate.rs:118-133: sync tone then data, real part only, one shared Carrier."""
    sps = 10 if scheme in ("msk", "oqpsk") else 45
    br, sr = rates(sps)
    kw = dict(scheme=scheme, baud_rate=br, sample_rate=sr, carrier_hz=1000)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(31, 3, 53 * o.bps)
    got = m.modulate_real(bits, preamble=preamble)
    ref = o.modulate_real(bits, preamble=preamble)
    assert_buffers(got, ref, f"{scheme} real output, preamble {preamble}")


@pytest.mark.parametrize("scheme", ["qpsk", "qam16", "bpsk", "qam256"])
@pytest.mark.parametrize("preamble", [0, 29, 64])
def test_modulate_real_fast_kernel(pkg, orc, scheme, preamble):
    """sps 8 takes the tuned TX kernel's real-output mode (two 32-bit stores per thread, NCO table built for
    Carrier.sample = preamble); odd preambles make every row start on an odd float."""
    kw = dict(scheme=scheme, baud_rate=1250, sample_rate=10000, carrier_hz=2500)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(33, 37, 1300 * o.bps)
    assert_buffers(m.modulate_real(bits, preamble=preamble), o.modulate_real(bits, preamble=preamble),
                   f"{scheme} real output (fast kernel), preamble {preamble}")
    # and the complex entry afterwards still uses its own table (sample0 = 0)
    assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} complex output after a real call")


@pytest.mark.parametrize("scheme,sps", [("qpsk", 45), ("qam16", 45), ("bpsk", 45)])
def test_tx_fast_kernel_odd_sps(pkg, orc, scheme, sps):
    """The reference's default rates (sps 45): the two samples of a 128-bit store straddle symbol edges; both the
    complex and the real-output form of the tuned kernel (even frame length)."""
    kw = dict(scheme=scheme, baud_rate=220, sample_rate=10000, carrier_hz=1000)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(34, 9, 146 * o.bps)
    assert o.frame_samples(bits.shape[1]) % 2 == 0
    assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} sps {sps} tx")
    assert_buffers(m.modulate_real(bits, preamble=19), o.modulate_real(bits, preamble=19), f"{scheme} sps {sps} real output")


def test_modulate_real_shaped(pkg, orc):
    rrc = orc.rrc_taps(16, 8, 0.35)
    kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, tx_taps=rrc, rx_taps=rrc,
              decision_delay=128)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    bits = rand_bits(32, 2, 2 * 300)
    tx = o.modulate(bits)
    assert_buffers(m.modulate_real(bits), np.ascontiguousarray(tx[:, :, 0]), "shaped real output")


def test_preamble_complex(pkg, orc):
    kw = dict(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=1000, sample0=7)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    assert_buffers(m.preamble(3, 199, 0.75), o.preamble(3, 199, 0.75), "sync tone")


# ----------------------------------------------------------------------------- demodulate binary path
def _wire(orc, bits, scheme="qpsk", br=220, sr=10000, cf=1000, preamble=199, scale=8000.0, theta=0.0):
    tx = orc.OraclePath(scheme, br, sr, cf).modulate_real(bits, preamble=preamble)
    return np.round(tx * scale).astype(np.int16)


@pytest.mark.parametrize("fmt", ["i16", "f32"])
def test_lock_phase(pkg, orc, fmt):
    """Demodulator::lock_phase over the Hilbert analytic signal: 64 sequential sincos + atan2 steps per frame."""
    rng = np.random.default_rng(41)
    n = np.arange(64 + 100)
    rows = []
    for f in range(37):  # tones with random phase / amplitude plus noise: exercises every atan2 quadrant
        th, a = rng.uniform(-np.pi, np.pi), rng.uniform(100, 20000)
        rows.append(a * np.cos(2 * np.pi * 0.09 * n + th) + rng.normal(0, 0.05 * a, n.size))
    x = np.array(rows)
    x = np.round(x).astype(np.int16) if fmt == "i16" else x.astype(np.float32)
    kw = dict(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=900)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    po = m.lock_phase(x, lock=64)
    po_ref, _, _, _ = o.demodulate_real(x, lock=64, want_filt=False)
    assert_buffers(po, po_ref, f"PLL phase offsets ({fmt})")
    assert len(np.unique(po)) > 30


def test_lock_phase_zero_input_quirk(pkg, orc):
    """All-zero input: 0 * cos is -0.0 when the cosine is negative and atan2(+-0, -0) = +-pi, so the reference's
    PLL random-walks; the GPU must reproduce the signed zeros."""
    kw = dict(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=1000)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    x = np.zeros((2, 80), np.int16)
    po_ref, _, _, _ = o.demodulate_real(x, lock=64, want_filt=False)
    assert_buffers(m.lock_phase(x, lock=64), po_ref, "PLL on zeros")
    assert po_ref[0] != 0.0


def test_lock_phase_analytic_input(pkg, orc):
    """Caller-supplied analytic signal (the generic Demodulator<S> of demodulator.rs:7-18)."""
    rng = np.random.default_rng(42)
    z = rng.normal(0, 1, (9, 70, 2)).astype(np.float32)
    kw = dict(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=900)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    po_ref, _, _, _ = o.demodulate_real(np.ascontiguousarray(z[:, :, 0]), lock=64, analytic_im=np.ascontiguousarray(z[:, :, 1]),
                                        want_filt=False)
    assert_buffers(m.lock_phase(z, lock=64), po_ref, "PLL on analytic input")


@pytest.mark.parametrize("fmt", ["i16", "f32"])
def test_demodulate_bin_path(pkg, orc, fmt):
    """src/bin/demodulate.rs end to end: i16 / f32 wire -> Hilbert -> lock -> Demodulator: the full-rate (I,Q)
    stream the binary prints, the locked offsets, and the decimated decisions."""
    sr, br, cf = 10000, 220, 1000
    sps = sr // br
    bits = rand_bits(43, 6, 2 * 150)
    P = sr // cf * 20 - 1
    wire = _wire(orc, bits, preamble=P)
    x = wire if fmt == "i16" else wire.astype(np.float32) / np.float32(3.0)
    lp = orc.lowpass_taps()
    g = 8000.0 if fmt == "i16" else 8000.0 / 3.0
    kw = dict(scheme="qpsk", baud_rate=br, sample_rate=sr, carrier_hz=cf, rx_taps=lp,
              decision_delay=(P - 64) + 31 + sps // 2, slicer_gain=float(np.float32(lp.sum()) * np.float32(g)))
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    got = m.demodulate_real(x, lock=64, want_soft=True)
    po, filt, sym, out = o.demodulate_real(x, lock=64)
    assert_buffers(got["phase_offset"], po, "locked phase offsets")
    assert_buffers(got["filt"], filt, "full-rate (I,Q) stream")
    assert np.array_equal(got["sym"], sym) and np.array_equal(got["bits"], out)
    K = sym.shape[1]
    assert np.array_equal(out, bits[:, : 2 * K])  # and the loop closes: 0 bit errors
    # soft values = the full-rate stream at the decision instants
    d = kw["decision_delay"]
    assert_buffers(got["soft"], np.ascontiguousarray(filt[:, d::sps][:, :K]), "decision-instant soft values")


def test_demodulate_real_no_lock_uses_cfg_offset(pkg, orc):
    bits = rand_bits(44, 4, 2 * 64)
    x = _wire(orc, bits, br=1250, preamble=0).astype(np.float32)
    lp = orc.lowpass_taps()
    kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=1000, rx_taps=lp, decision_delay=35,
              slicer_gain=float(np.float32(lp.sum()) * np.float32(8000.0)), phase_offset=0.25)
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    got = m.demodulate_real(x, lock=0)
    po, filt, sym, out = o.demodulate_real(x, lock=0)
    assert_buffers(got["filt"], filt, "no-lock (I,Q) stream")
    assert np.array_equal(got["sym"], sym)
    # identical to the complex-input entry fed (x, 0)
    z = np.stack([x, np.zeros_like(x)], axis=-1)
    assert_buffers(m.demodulate(z, want_filt=True)["filt"], filt, "complex entry on the same samples")


@pytest.mark.parametrize("fmt,lock,delay,pre,br", [("i16", 64, 35, 64, 1250), ("f32", 64, 35, 79, 1250), ("i16", 0, 35, 0, 1250), ("f32", 64, 36, 64, 1250),
                                                   ("i16", 64, 36, 79, 1250), ("i16", 64, 35, 39 + 64, 1250),
                                                   ("i16", 64, 53, 199, 220), ("f32", 64, 53, 64, 220), ("i16", 64, 36, 79, 1000), ("f32", 64, 33, 65, 2000)])
def test_bin_path_fast_kernel_sps8(pkg, orc, fmt, lock, delay, pre, br, monkeypatch):
    """The demodulate binary's path through the tuned kernels (sps 8: rx_fast_raw.cu; the reference's own rates, sps 45, and
    other counts: rx_dec_kernel<..., RAW>): real f32 / i16
    rows, lock samples skipped, one PLL offset per frame, glibc cos / sin evaluated per frame): 23 frames x 3 tiles,
    odd and even decision delays, against the oracle (offsets, decisions, soft values bit for bit) and against the
    generic kernel (MODEM_GPU_FORCE_GENERIC)."""
    sr, cf = 10000, 900
    sps = sr // br  # 8: rx_fast_raw.cu; 45 (the reference's own rates), 10, 5: rx_dec_kernel<..., RAW>
    bits = rand_bits(47, 23, 2 * (700 if sps <= 10 else 300))
    # the carrier tone (the lock consumes its first 64 samples), then the data; odd preambles (the reference's sr/cf*pc - 1)
    # give rows with an odd stride and an odd number of samples behind the lock
    wire = _wire(orc, bits, br=br, cf=cf, preamble=pre)
    if lock:  # every frame its own dither on the tone: every frame locks to its own offset
        wire[:, :lock] += np.random.default_rng(48).integers(-900, 900, (wire.shape[0], lock)).astype(np.int16)
    x = wire if fmt == "i16" else wire.astype(np.float32) / np.float32(3.0)
    lp = orc.lowpass_taps()
    g = 8000.0 if fmt == "i16" else 8000.0 / 3.0
    delay += pre - lock  # the rest of the tone delays the first symbol
    kw = dict(scheme="qpsk", baud_rate=br, sample_rate=sr, carrier_hz=cf, rx_taps=lp, decision_delay=delay,
              slicer_gain=float(np.float32(lp.sum()) * np.float32(g)))
    m, o = pkg.Modem(**kw), orc.OraclePath(**kw)
    got = m.demodulate_real(x, lock=lock, want_soft=True, want_filt=False)
    po, filt, sym, out = o.demodulate_real(x, lock=lock)
    if lock:
        assert_buffers(got["phase_offset"], po, "locked phase offsets")
        assert len(np.unique(po)) > 1
    assert np.array_equal(got["sym"], sym) and np.array_equal(got["bits"], out)
    K = sym.shape[1]
    assert_buffers(got["soft"], np.ascontiguousarray(filt[:, delay::sps][:, :K]), "decision-instant soft values")
    monkeypatch.setenv("MODEM_GPU_FORCE_GENERIC", "1")
    m2 = pkg.Modem(**kw)
    ref = m2.demodulate_real(x, lock=lock, want_soft=True, want_filt=False)
    assert np.array_equal(ref["sym"], got["sym"]) and np.array_equal(ref["soft"].view(np.uint32), got["soft"].view(np.uint32))


def test_demodulate_real_short_input(pkg):
    m = pkg.Modem(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=1000)
    with pytest.raises(pkg.ModemError):  # demodulator.rs:34 unwrap() on None
        m.demodulate_real(np.zeros((1, 63), np.int16), lock=64)
    out = m.demodulate_real(np.zeros((1, 64), np.int16), lock=64)
    assert out["filt"].shape == (1, 0, 2)


def test_bin_path_many_frames_multichannel(pkg, orc):
    """A bank: every frame its own carrier AND its own locked offset."""
    sr, br = 10000, 1250
    bits = rand_bits(45, 12, 2 * 200)
    lp = orc.lowpass_taps()
    hz = [500 + 40 * c for c in range(4)]
    kw = dict(scheme="qpsk", baud_rate=br, sample_rate=sr, rx_taps=lp, decision_delay=35,
              slicer_gain=float(np.float32(lp.sum()) * np.float32(8000.0)))
    m = pkg.Modem(carrier_hz=hz[0], **kw)
    m.set_channels([pkg.sample_freq(h, sr) for h in hz], 3)
    rows, refs = [], []
    for c, h in enumerate(hz):
        b = bits[3 * c: 3 * c + 3]
        w = _wire(orc, b, br=br, cf=h, preamble=64)
        rows.append(w)
        refs.append(orc.OraclePath(carrier_hz=h, **kw).demodulate_real(w, lock=64))
    x = np.concatenate(rows)
    got = m.demodulate_real(x, lock=64)
    assert_buffers(got["phase_offset"], np.concatenate([r[0] for r in refs]), "bank offsets")
    assert_buffers(got["filt"], np.concatenate([r[1] for r in refs]), "bank (I,Q)")
    assert np.array_equal(got["sym"], np.concatenate([r[2] for r in refs]))


def test_device_libm_matches_glibc_exhaustively():
    """tools/check_sincos_dev.cu: the device's cosf / sinf routines for every binary32 with |y| < 120 (both signs, 4.5e9
    evaluations) and atanf for all 2^32 inputs against this box's glibc, through per-run checksums."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "bin", "check_sincos_dev")
    if not os.path.exists(exe):
        pytest.skip("tools/bin/check_sincos_dev is not built (tools/build_tools.sh)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIT-EXACT" in r.stdout and "0 runs differ" in r.stdout.splitlines()[-1]
