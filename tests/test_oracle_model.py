"""Independent numpy model of the stages the reference has no test for (carrier, mixer, FIR,
demodulator), checked against the C oracle.  Integer / IEEE-exact steps must agree bit for
bit; steps that go through a different libm (numpy's sin/cos) agree to a few ULP."""
import math

import numpy as np

F32 = np.float32
TWO_PI = F32(math.pi) * F32(2.0)


def np_phase(w, n):
    """carrier.rs:17-19 + util.rs:3-6 in numpy binary32 (mul, div, floor, mul, sub: all IEEE exact)."""
    x = F32(w) * n.astype(F32)
    return x - TWO_PI * np.floor(x / TWO_PI)


def np_fir(x, h):
    """fir.rs:18-34: y[n] = sum_k h[k] x[n-k], x[<0] = 0, folded k = 0..N-1 from 0.0f, mul then add."""
    acc = np.zeros_like(x, dtype=F32)
    for k in range(len(h)):
        shifted = np.zeros_like(x, dtype=F32)
        shifted[k:] = x[: len(x) - k]
        acc = acc + shifted * F32(h[k])
    return acc


def test_nco_phase_bit_exact(orc):
    import ctypes as C
    L = orc.lib()
    c = orc.Carrier()
    L.orc_carrier_new(C.byref(c), 1000, 10000)
    got = np.array([L.orc_carrier_next(C.byref(c)) for _ in range(5000)], F32)
    want = np_phase(L.orc_sample_freq(1000, 10000), np.arange(5000, dtype=np.uint64))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # binary32 mod_trig can land a few 1e-5 ABOVE 2*pi (x/2pi rounds down across an integer): a reference quirk
    assert got.min() >= 0 and got.max() < float(TWO_PI) + 1e-3


def test_fir_matches_convolution(orc):
    import ctypes as C
    L = orc.lib()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(700).astype(F32)
    for h in (orc.lowpass_taps(), orc.rrc_taps(16, 8, 0.35), orc.hilbert_taps(), np.array([0.5], F32)):
        f = orc.Fir()
        hh = np.ascontiguousarray(h, F32)
        assert L.orc_fir_new(C.byref(f), hh.ctypes.data_as(C.POINTER(C.c_float)), len(hh))
        got = np.array([L.orc_fir_add(C.byref(f), float(v)) for v in x], F32)
        L.orc_fir_free(C.byref(f))
        want = np_fir(x, hh)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # same order, same roundings
        ref64 = np.convolve(x.astype(np.float64), hh.astype(np.float64))[: len(x)]  # SURVEY: == np.convolve(x,h)[:len(x)]
        assert np.abs(got - ref64).max() < 2e-5 * max(1.0, np.abs(ref64).max())


def test_modulator_and_demodulator_against_numpy(orc):
    o = orc.OraclePath("qpsk", 1250, 10000, 2500, rx_taps=orc.lowpass_taps(), decision_delay=35,
                       slicer_gain=float(F32(orc.lowpass_taps().sum())))
    bits = np.random.default_rng(9).integers(0, 2, (1, 600), dtype=np.uint8)
    tx, iq = o.modulate(bits, want_iq=True)
    n = np.arange(tx.shape[1], dtype=np.uint64)
    ph = np_phase(orc.lib().orc_sample_freq(2500, 10000), n)
    # rectangular hold of the constellation (data.rs:66-79): sample n carries symbol n // sps
    const = o.constellation()[0]
    idx = bits[0, 0::2] * 2 + bits[0, 1::2]
    want_iq = const[np.repeat(idx, o.sps)]
    assert np.array_equal(iq[0].view(np.uint32), want_iq.view(np.uint32))
    # mixer (modulator.rs:37-48) with numpy trig: a few ULP of libm difference at most
    c, s = np.cos(ph, dtype=F32), np.sin(ph, dtype=F32)
    re = want_iq[:, 0] * c - want_iq[:, 1] * s
    im = want_iq[:, 0] * s + want_iq[:, 1] * c
    assert np.abs(tx[0, :, 0] - re).max() < 5e-7 and np.abs(tx[0, :, 1] - im).max() < 5e-7
    # demodulator (demodulator.rs:44-55) on the oracle's own tx
    filt, sym, dec = o.demodulate(tx)
    x = tx[0, :, 0]
    want_i = F32(2.0) * np_fir(x * c, orc.lowpass_taps())
    want_q = F32(2.0) * np_fir(x * -s, orc.lowpass_taps())
    assert np.abs(filt[0, :, 0] - want_i).max() < 2e-6 and np.abs(filt[0, :, 1] - want_q).max() < 2e-6
    K = sym.shape[1]
    assert np.array_equal(dec[0], bits[0, : 2 * K])  # noise-free round trip
    assert np.array_equal(sym[0], idx[:K])


def test_awgn_statistics(orc):
    o = orc.OraclePath("qpsk", 1250, 10000, 2500)
    z = o.awgn(np.zeros((4, 50000, 2), F32), 1.0, seed=123)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1])) < 0.01
    # counter-based: frame f of a call with frame0 = k equals frame 0 of a call with frame0 = k + f
    a = o.awgn(np.zeros((3, 64, 2), F32), 0.5, seed=9, frame0=10)
    b = o.awgn(np.zeros((1, 64, 2), F32), 0.5, seed=9, frame0=12)
    assert np.array_equal(a[2], b[0])


def test_ber_matches_theory(orc):
    """RRC/RRC QPSK through the AWGN extension: BER within 4 sigma of Q(sqrt(2 Eb/N0))."""
    rrc = orc.rrc_taps(16, 8, 0.35)
    o = orc.OraclePath("qpsk", 1250, 10000, 2500, tx_taps=rrc, rx_taps=rrc, decision_delay=128, slicer_gain=1.0)
    bits = np.random.default_rng(4).integers(0, 2, (48, 8192), dtype=np.uint8)
    for db in (2.0, 5.0):
        _, _, (err, n) = o.loopback(bits, sigma=o.sigma_for_ebn0(db), seed=77, threads=8, want_out=False)
        p = 0.5 * math.erfc(math.sqrt(10 ** (db / 10)))
        assert abs(err / n - p) < 4 * math.sqrt(p / n), (db, err / n, p)


def test_shaped_tx_sign_product_form(orc):
    """The GPU's shaped TX replaces multiply-round-add-round by ONE fma(+-1, round(h[k]*v), acc) when every
    constellation point is (+-vi, +-vq), and skips the zero-stuffed terms (csrc/modem_api.cu `tx_sign_form`).
    Model that form in numpy binary32 (a product by +-1 is exact, so `acc + s*p` rounds once, like the fma) and
    compare it with the oracle's literal fold over all taps (fir.rs:21-24): bit-identical, signed zeros included."""
    from conftest import path_kwargs

    for scheme in ("qpsk", "bpsk"):
        kw = path_kwargs(scheme, sps=8, shaped=True)
        o = orc.OraclePath(**kw)
        con = o.constellation()[0]
        v = np.abs(con[0])
        assert (np.abs(con).view(np.uint32) == v.view(np.uint32)).all(), "one magnitude per rail"
        sign = np.where(np.signbit(con), F32(-1), F32(1))
        h = np.asarray(kw["tx_taps"], F32)
        rail_taps = (h[:, None] * v[None, :]).astype(F32)  # round(h[k] * v), one rounding
        bits = np.random.default_rng(5).integers(0, 2, (1, o.bps * 90), dtype=np.uint8)
        _, iq_ref = o.modulate(bits, want_iq=True)
        nsym = bits.shape[1] // o.bps
        sym = np.zeros(nsym, np.int64)
        for b in range(o.bps):
            sym = (sym << 1) | bits[0, b::o.bps][:nsym]
        s = sign[sym]  # [nsym][2]
        sps, L = 8, nsym * 8
        acc = np.zeros((L, 2), F32)
        n = np.arange(L)
        for k in range(len(h)):  # ascending taps; only n-k on a symbol instant contributes
            m = n - k
            hit = (m >= 0) & (m % sps == 0)
            acc[hit] = acc[hit] + s[m[hit] // sps] * rail_taps[k]
        assert (acc.view(np.uint32) == iq_ref[0].view(np.uint32)).all()
