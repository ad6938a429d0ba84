"""CPU checks of the oracle's restatement of the two binaries' sample paths:
the sync tone + real output of src/bin/modulate.rs, the Hilbert + PLL lock + Demodulator of
src/bin/demodulate.rs, and the stateful mappers.  The reference has no test for any of these except
test_dmpsk (tests/test_oracle_kat.py): parity is pinned by restating the cited lines; the checks here are
independent numpy models of the same formulas plus end-to-end behaviour (the PLL locks, the loop closes)."""
import ctypes as C

import numpy as np
import pytest


def f32(x):
    return np.float32(x)


def mod_trig(x):
    two_pi = f32(np.float32(np.pi) * f32(2.0))
    return f32(x - f32(two_pi * np.floor(f32(x / two_pi))))


def test_preamble_tone_is_real_part_of_carrier(orc):
    """modulate.rs:118-126: Raw(A) -> i = A, q = 0.0; re = A*cos - 0.0*sin (modulator.rs:37-39)."""
    o = orc.OraclePath("qpsk", 220, 10000, 1000)
    tone = o.preamble(2, 99, 1.0)
    w = orc.lib().orc_sample_freq(1000, 10000)
    n = np.arange(99, dtype=np.float32)
    ph = np.array([mod_trig(f32(w) * k) for k in n], np.float32)
    assert np.allclose(tone[0, :, 0], np.cos(ph.astype(np.float64)), atol=2e-7)
    assert np.allclose(tone[0, :, 1], np.sin(ph.astype(np.float64)), atol=2e-7)
    assert np.array_equal(tone[0], tone[1])


def test_modulate_real_shares_one_carrier(orc):
    """modulate.rs:120,128: the data continues the tone's sample counter; .re of the complex path at sample0 = P."""
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, (2, 40), dtype=np.uint8)
    P = 10000 // 1000 * 3 - 1  # sr / cf * pc - 1 (modulate.rs:125) with -p 3
    o = orc.OraclePath("qpsk", 220, 10000, 1000)
    out = o.modulate_real(bits, preamble=P)
    L = o.frame_samples(40)
    assert out.shape == (2, P + L)
    assert np.array_equal(out[:, :P], o.preamble(2, P)[:, :, 0])
    shifted = orc.OraclePath("qpsk", 220, 10000, 1000, sample0=P)
    assert np.array_equal(out[:, P:], shifted.modulate(bits)[:, :, 0])


@pytest.mark.parametrize("scheme", ["bfsk", "mfsk", "16cpfsk", "msk", "dqpsk", "dbpsk"])
def test_stateful_phasors_unit_envelope(orc, scheme):
    """Every FSK/MSK/DPSK mapper of the reference is constant-envelope: |i + jq| = AMPLITUDE."""
    o = orc.OraclePath(scheme, 1250, 10000, 2500)
    bits = np.random.default_rng(3).integers(0, 2, (2, 64 * o.bps), dtype=np.uint8)
    tx, iq = o.modulate(bits, want_iq=True)
    if scheme == "msk":  # i = +-cos, q = -+sin of the same angle (msk.rs:29-35)
        assert np.allclose(np.hypot(iq[..., 0], iq[..., 1]), 1.0, atol=1e-6)
    else:
        assert np.allclose(np.hypot(iq[..., 0], iq[..., 1]), 1.0, atol=1e-6)
    assert np.allclose(np.hypot(tx[..., 0], tx[..., 1]), 1.0, atol=1e-6)


def test_dmpsk_phase_recurrence_numpy_model(orc):
    """dmpsk.rs:29-33 against an independent numpy float32 model of the recurrence."""
    o = orc.OraclePath("dqpsk", 1250, 10000, 2500)
    bits = np.random.default_rng(4).integers(0, 2, (1, 2 * 200), dtype=np.uint8)
    _, iq = o.modulate(bits, want_iq=True)
    phase = f32(np.float32(np.pi) / f32(4.0))
    shift = f32(np.float32(np.pi) / f32(2.0))
    for k in range(200):
        sym = int(bits[0, 2 * k]) * 2 + int(bits[0, 2 * k + 1])
        phase = mod_trig(f32(phase + f32(f32(sym) * shift)))
        got = iq[0, k * 8: (k + 1) * 8]
        assert np.all(got == got[0])  # held for the whole symbol
        assert abs(got[0, 0] - np.cos(np.float64(phase))) < 2e-7 and abs(got[0, 1] - np.sin(np.float64(phase))) < 2e-7


def test_bfsk_phase_continuity(orc):
    """bfsk.rs:43-55: update() re-bases the phase so the waveform is continuous across bit changes."""
    o = orc.OraclePath("bfsk", 1250, 10000, 2500)
    bits = np.random.default_rng(5).integers(0, 2, (1, 300), dtype=np.uint8)
    _, iq = o.modulate(bits, want_iq=True)
    ang = np.unwrap(np.arctan2(iq[0, :, 1].astype(np.float64), iq[0, :, 0].astype(np.float64)))
    step = np.diff(ang)
    dev = float(orc.lib().orc_sample_freq(200, 10000))
    assert step.min() > -1e-3 and step.max() < dev + 1e-3  # frequency is 0 or +deviation, never a jump


def test_pll_locks_to_carrier_phase(orc):
    """demodulate.rs:31-39 on a pure tone with a phase offset: after 64 samples PLL.phase_offset ~ the offset
    (mod 2pi) and the demodulated (I,Q) settles at (A, 0) rotated by the residual.  The carrier is the binary's
    own 900 Hz (demodulate.rs:36): the 23-tap Hilbert FIR delays the imaginary part by 11 samples = 0.99 carrier
    cycles there, so the analytic signal is nearly consistent (at 1000 Hz the same lock is biased by ~0.2-0.5 rad)."""
    sr, cf, A = 10000, 900, 8000.0
    n = np.arange(64 + 400)
    for theta in (0.0, 0.7, -1.3, 2.5):
        x = np.round(A * np.cos(2 * np.pi * cf / sr * n + theta)).astype(np.int16)
        o = orc.OraclePath("qpsk", 220, sr, cf)
        po, filt, _, _ = o.demodulate_real(x[None, :], lock=64)
        err = (po[0] - theta + np.pi) % (2 * np.pi) - np.pi
        assert abs(err) < 0.06, (theta, po[0])
        I, Q = filt[0, 200:, 0], filt[0, 200:, 1]
        assert np.allclose(np.hypot(I, Q), A * 0.99864417, rtol=2e-3)
        assert np.all(np.abs(np.arctan2(Q, I) + err) < 0.01)  # rotated by exactly the residual lock error


def test_bin_loop_closes(orc):
    """modulate (sync tone + QPSK, reference default rates) -> i16 wire -> demodulate (Hilbert + PLL lock +
    low-pass) -> decimate/slice: 0 bit errors.  The lock consumes 64 of the tone's samples."""
    rng = np.random.default_rng(9)
    sr, br, cf = 10000, 220, 1000
    sps = sr // br
    bits = rng.integers(0, 2, (3, 2 * 120), dtype=np.uint8)
    P = sr // cf * 20 - 1  # -p 20
    tx = orc.OraclePath("qpsk", br, sr, cf).modulate_real(bits, preamble=P)
    wire = np.round(tx * 8000.0).astype(np.int16)
    lp = orc.lowpass_taps()
    delay = (P - 64) + 31 + sps // 2
    rx = orc.OraclePath("qpsk", br, sr, cf, rx_taps=lp, decision_delay=delay, slicer_gain=float(f32(lp.sum()) * 8000.0))
    po, filt, sym, out = rx.demodulate_real(wire, lock=64)
    K = sym.shape[1]
    assert K >= 118
    assert np.array_equal(out, bits[:, : K * 2])
    assert np.all(np.abs(po) < 0.6)  # biased by the Hilbert FIR's 11-sample delay at 1000 Hz, inside QPSK's pi/4 margin


def test_demodulate_real_short_input_panics(orc):
    o = orc.OraclePath("qpsk", 220, 10000, 1000)
    with pytest.raises(ValueError):
        o.demodulate_real(np.zeros((1, 63), np.int16), lock=64)
    po, filt, _, _ = o.demodulate_real(np.zeros((1, 64), np.int16), lock=64)
    assert filt.shape == (1, 0, 2)
    # all-zero input is NOT a fixed point of the reference's PLL: 0 * cos(inner) is -0.0 whenever the cosine is
    # negative and atan2(+-0, -0) = +-pi (pll.rs:19), so the offset random-walks in steps of 0.447214 * pi.
    steps = po[0] / (np.float32(0.447214) * np.float32(np.pi))
    assert abs(steps - round(steps)) < 1e-4 and po[0] != 0.0
