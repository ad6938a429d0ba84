"""The reference's own unit tests, restated against the CPU oracle.

Each test cites the reference test it reproduces (paths relative to /root/reference/).
These are the only known-answer vectors the reference holds for this path; they pin
the oracle's symbol clock, bit source, mapper helpers and constellation formulas.
"""
import ctypes as C
import math

import numpy as np

PI32 = float(np.float32(math.pi))


def _bits(*b):
    return (C.c_uint8 * len(b))(*b)


def test_symbol_clock(orc):
    # src/modem/data.rs:195-209
    L = orc.lib()
    c = orc.SymbolClock()
    L.orc_symbol_clock_new(C.byref(c), 5)
    got = [bool(L.orc_symbol_clock_next(C.byref(c))) for _ in range(11)]
    assert got == [True, False, False, False, False, True, False, False, False, False, True]


def _drain(orc, src, n):
    L = orc.lib()
    out = []
    for _ in range(n):
        u = L.orc_source_next(C.byref(src))
        out.append((u.kind, None if u.kind == orc.FINISHED else [u.bits[i] for i in range(u.len)]))
    return out


def test_bits(orc):
    # src/modem/data.rs:212-224
    L = orc.lib()
    raw = _bits(1, 0, 1, 1)
    s = orc.Source()
    L.orc_bits_new(C.byref(s), raw, 4, 3, 2)
    Ch, Un, Fi = orc.CHANGED, orc.UNCHANGED, orc.FINISHED
    assert _drain(orc, s, 7) == [(Ch, [1, 0]), (Un, [1, 0]), (Un, [1, 0]), (Ch, [1, 1]), (Un, [1, 1]),
                                 (Un, [1, 1]), (Fi, None)]


def test_bits_drops_partial_symbol(orc):
    # data.rs:58-62: a trailing partial symbol ends the stream
    L = orc.lib()
    raw = _bits(1, 0, 1)
    s = orc.Source()
    L.orc_bits_new(C.byref(s), raw, 3, 2, 2)
    Ch, Un, Fi = orc.CHANGED, orc.UNCHANGED, orc.FINISHED
    assert _drain(orc, s, 3) == [(Ch, [1, 0]), (Un, [1, 0]), (Fi, None)]


def test_evenodd(orc):
    # src/modem/data.rs:227-246
    L = orc.lib()
    raw = _bits(1, 1, 1, 0, 0, 1)
    s = orc.Source()
    L.orc_evenodd_new(C.byref(s), raw, 6, 4, 2)
    Ch, Un, Fi = orc.CHANGED, orc.UNCHANGED, orc.FINISHED
    assert _drain(orc, s, 13) == [
        (Ch, [1, 0]), (Un, [1, 0]), (Ch, [1, 1]), (Un, [1, 1]),
        (Ch, [1, 1]), (Un, [1, 1]), (Ch, [1, 0]), (Un, [1, 0]),
        (Ch, [0, 0]), (Un, [0, 0]), (Ch, [0, 1]), (Un, [0, 1]), (Fi, None)]


def test_b2b(orc):
    # src/modem/digital/util.rs:22-25
    L = orc.lib()
    assert L.orc_bytes_to_bits(_bits(0, 0, 0, 1), 4) == 0b0001
    assert L.orc_bytes_to_bits(_bits(0, 1, 0, 1), 4) == 0b0101


def test_max_symbol(orc):
    # src/modem/digital/util.rs:28-33
    L = orc.lib()
    assert [L.orc_max_symbol(b) for b in (1, 2, 4, 8)] == [0b1, 0b11, 0b1111, 0b11111111]


def test_bit_to_sign(orc):
    # src/modem/digital/util.rs:1-3
    L = orc.lib()
    assert L.orc_bit_to_sign(0) == -1.0 and L.orc_bit_to_sign(1) == 1.0


def test_mpsk(orc):
    # src/modem/digital/mpsk.rs:50-63 (exact equality where the reference asserts it)
    L = orc.lib()
    p = orc.Phasor()
    L.orc_mpsk_new(C.byref(p), 2, 0.0, 1.0)
    i = lambda *b: L.orc_phasor_i(C.byref(p), 0, _bits(*b))
    q = lambda *b: L.orc_phasor_q(C.byref(p), 0, _bits(*b))
    assert i(0, 0) == 1.0 and q(0, 0) == 0.0
    assert abs(i(0, 1)) < 0.001 and q(0, 1) == 1.0
    assert i(1, 0) == -1.0 and abs(q(1, 0)) < 0.001
    assert abs(i(1, 1)) < 0.001 and q(1, 1) == -1.0


def test_qam(orc):
    # src/modem/digital/qam.rs:69-84
    L = orc.lib()
    p = orc.Phasor()
    L.orc_qam_new(C.byref(p), 4, 0.0, 6.0)
    i = lambda *b: L.orc_phasor_i(C.byref(p), 0, _bits(*b))
    q = lambda *b: L.orc_phasor_q(C.byref(p), 0, _bits(*b))
    assert (i(0, 0, 0, 0), q(0, 0, 0, 0)) == (-3.0, -3.0)
    assert (i(0, 0, 0, 1), q(0, 0, 0, 1)) == (-3.0, -1.0)
    assert (i(1, 0, 1, 1), q(1, 0, 1, 1)) == (1.0, 3.0)
    assert (i(1, 1, 1, 1), q(1, 1, 1, 1)) == (3.0, 3.0)


def test_dmpsk(orc):
    # src/modem/digital/dmpsk.rs:51-84
    L = orc.lib()
    d = orc.Phasor()
    L.orc_dmpsk_new(C.byref(d), 2, 1.0, 0.0, PI32 / 2.0)
    e = _bits(0)
    iq = lambda: (L.orc_phasor_i(C.byref(d), 0, e), L.orc_phasor_q(C.byref(d), 0, e))

    def near(got, want):
        assert abs(got[0] - want[0]) < 0.000001 and abs(got[1] - want[1]) < 0.000001, (got, want)

    near(iq(), (1.0, 0.0))
    for bits, want in [((0, 0), (1.0, 0.0)), ((0, 1), (0.0, 1.0)), ((1, 0), (0.0, -1.0)), ((1, 1), (-1.0, 0.0)),
                       ((0, 0), (-1.0, 0.0)), ((0, 0), (-1.0, 0.0)), ((1, 1), (0.0, 1.0))]:
        L.orc_phasor_update(C.byref(d), 123, _bits(*bits))
        near(iq(), want)


def test_apsk_verify(orc):
    # src/modem/digital/apsk.rs:85-97 and the 16apsk rings of src/bin/modulate.rs:88-91
    L = orc.lib()
    p = orc.Phasor()
    good = (orc.Ring * 2)(orc.Ring(0, 4, 0.5, PI32 / 4), orc.Ring(4, 16, 1.0, PI32 / 12))
    assert L.orc_apsk_new(C.byref(p), 1.0, 4, good, 2) == 1
    gap = (orc.Ring * 2)(orc.Ring(0, 4, 0.5, 0.0), orc.Ring(5, 16, 1.0, 0.0))
    assert L.orc_apsk_new(C.byref(p), 1.0, 4, gap, 2) == 0
    short = (orc.Ring * 1)(orc.Ring(0, 4, 0.5, 0.0))
    assert L.orc_apsk_new(C.byref(p), 1.0, 4, short, 1) == 0


def test_modulate_names(orc):
    # every -m name of src/bin/modulate.rs:74-95 resolves; anything else is the panic path (:94)
    L = orc.lib()
    p = orc.Phasor()
    bps = {"bask": 1, "bpsk": 1, "bfsk": 1, "qpsk": 2, "qam16": 4, "qam256": 8, "msk": 2, "mfsk": 4, "16psk": 4,
           "oqpsk": 2, "dcqpsk": 2, "16cpfsk": 4, "16apsk": 4, "dqpsk": 2, "dbpsk": 1}
    for name, b in bps.items():
        assert L.orc_phasor_by_name(C.byref(p), name.encode(), 220, 10000) == 1, name
        assert p.bits_per_symbol == b, name
    assert L.orc_phasor_by_name(C.byref(p), b"nope", 220, 10000) == 0


def test_rates_and_freq(orc):
    # src/modem/rates.rs:16 integer division; src/modem/freq.rs:19-26; defaults of modulate.rs:44-58
    L = orc.lib()
    assert L.orc_samples_per_symbol(220, 10000) == 45
    two_pi = np.float32(2.0) * np.float32(math.pi)
    w = np.float32(two_pi * np.float32(1000)) / np.float32(10000)
    assert L.orc_sample_freq(1000, 10000) == float(w)


def test_tap_tables(orc):
    # src/bin/demodulate.rs:82-147 (64 symmetric taps, DC gain 0.99864417) and :48-72 (23 antisymmetric)
    lp = orc.lowpass_taps()
    assert lp.shape == (64,) and np.array_equal(lp, lp[::-1])
    assert abs(float(lp.astype(np.float64).sum()) - 0.99864417) < 1e-7
    assert lp[0] == np.float32(8.6464950643449706e-05) and lp[31] == np.float32(0.24347923270043995)
    hb = orc.hilbert_taps()
    assert hb.shape == (23,) and hb[11] == 0.0 and hb[10] == np.float32(-0.62794) and hb[12] == np.float32(0.62794)


def test_philox_kat(orc):
    # Random123 known-answer vectors for philox4x32-10 (Salmon et al., kat_vectors)
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert orc.philox4x32_10([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
