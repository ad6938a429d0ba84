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


# ----------------------------------------------------------------------------- the AWGN extension, against its written spec
def _rn32(x):
    """Nearest binary32 (ties to even) of an exact rational, as a Fraction."""
    from fractions import Fraction

    if x == 0:
        return Fraction(0)
    s, a = (-1 if x < 0 else 1), abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    e = max(e, -126)
    quantum = Fraction(2) ** (e - 23)
    q = a / quantum
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return s * n * quantum


def _box_muller_spec(r0, r1):
    """oracle/modem_oracle.h "AWGN", evaluated in exact rational arithmetic with one binary32 rounding per written
    operation (fmaf = one rounding of a*b + c) -- an independent restatement of the spec, not of the C code."""
    from decimal import Decimal, getcontext
    from fractions import Fraction as Fr

    hexf = lambda s: Fr(float.fromhex(s))
    Lc = [hexf(s) for s in ("-0x1.00001cp-1", "0x1.555802p-2", "-0x1.ffa938p-3", "0x1.97ecccp-3", "-0x1.5e404cp-3", "0x1.495358p-3", "-0x1.ab64d2p-4")]
    Sc = [hexf(s) for s in ("-0x1.555552p-3", "0x1.110c2ap-7", "-0x1.9aca02p-13")]
    Cc = [hexf(s) for s in ("-0x1p-1", "0x1.55554cp-5", "-0x1.6c0e0cp-10", "0x1.9a6fd8p-16")]
    LN2, ANG = hexf("0x1.62e43p-1"), hexf("0x1.921fb6p-22")
    fma = lambda a, b, c: _rn32(a * b + c)
    mul = lambda a, b: _rn32(a * b)
    u = mul(_rn32(Fr(r0 >> 9) + Fr(1, 2)), Fr(1, 2 ** 23))
    ub = int(np.float32(float(u)).view(np.uint32))
    ix = (ub - 0x3F3504F3) & 0xFFFFFFFF
    e = (ix >> 23) - (512 if ix & 0x80000000 else 0)
    m = Fr(float(np.uint32((ix & 0x7FFFFF) + 0x3F3504F3).view(np.float32)))
    f = _rn32(m - 1)
    q = Lc[6]
    for k in range(5, -1, -1):
        q = fma(q, f, Lc[k])
    lnm = mul(f, fma(f, q, Fr(1)))
    lnu = fma(Fr(e), LN2, lnm)
    t = mul(Fr(-2), lnu)
    getcontext().prec = 60
    rad = _rn32(Fr((Decimal(t.numerator) / Decimal(t.denominator)).sqrt()))
    j = r1 >> 8
    octant, k = j >> 21, j & 0x1FFFFF
    if octant & 1:
        k = 0x200000 - k
    x = mul(Fr(k), ANG)
    z = mul(x, x)
    sn = fma(mul(x, z), fma(fma(Sc[2], z, Sc[1]), z, Sc[0]), x)
    cs = fma(z, fma(fma(fma(Cc[3], z, Cc[2]), z, Cc[1]), z, Cc[0]), Fr(1))
    if (octant + 1) & 2:
        sn, cs = cs, sn
    if (octant + 2) & 4:
        cs = -cs
    if octant & 4:
        sn = -sn
    return float(mul(rad, cs)), float(mul(rad, sn))


def test_box_muller_follows_its_written_spec(orc):
    """The C oracle's normal pair against an exact-rational evaluation of the specification in modem_oracle.h: 400
    word pairs incl. the extremes of both uniforms and every octant boundary -- bit for bit."""
    import math

    rng = np.random.default_rng(2024)
    words = [(int(a), int(b)) for a, b in rng.integers(0, 2 ** 32, (340, 2), dtype=np.uint64)]
    words += [(0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (0x1FF, 0xFF), (0x200, 0x100), (0xFFFFFE00, 0x80000000), (0x5A82799A, 0x7FFFFFFF)]
    words += [(int(rng.integers(0, 2 ** 32)), (o << 29) + d & 0xFFFFFFFF) for o in range(8) for d in (0, 0x100, 0x1FFFFF00 & 0xFFFFFFFF)]
    for r0, r1 in words:
        got = orc.box_muller(r0, r1)
        want = _box_muller_spec(r0, r1)
        for g, w in zip(got, want):  # exact rationals carry no sign of zero: -0.0 (sine negated in octants 4..7) == 0
            assert (g == 0 and w == 0) or np.float32(g).view(np.uint32) == np.float32(w).view(np.uint32), (hex(r0), hex(r1), got, want)
        # ... and the spec is a Box-Muller transform: the same values as the textbook formula to ~1e-6
        u, th = ((r0 >> 9) + 0.5) * 2.0 ** -23, (r1 >> 8) * 2.0 ** -24 * 2 * math.pi
        rad = math.sqrt(-2 * math.log(u))
        assert abs(got[0] - rad * math.cos(th)) < 3e-6 and abs(got[1] - rad * math.sin(th)) < 3e-6


def test_awgn_word_assignment_and_tails(orc):
    """One Philox block per aligned quad and rail; normality far into the tails (the BER points of config 4 at 10 dB sit
    at 4.5 sigma): the empirical tail mass of 4e6 samples follows erfc out to 4 sigma."""
    import math

    o = orc.OraclePath("qpsk", 1250, 10000, 2500)
    z = o.awgn(np.zeros((1, 10, 2), F32), 1.0, seed=0x1234, frame0=5)[0]
    for n in range(10):
        for rail in range(2):
            r = orc.philox4x32_10([n >> 2, rail << 31, 5, 0], [0x1234, 0])
            a = orc.box_muller(r[2], r[3]) if n & 2 else orc.box_muller(r[0], r[1])
            assert z[n, rail] == np.float32(a[n & 1])
    big = o.awgn(np.zeros((8, 250000, 2), F32), 1.0, seed=99).ravel()
    for t in (1.0, 2.0, 3.0, 4.0):
        p = math.erfc(t / math.sqrt(2))  # two-sided
        k = int((np.abs(big) > t).sum())
        assert abs(k - p * big.size) < 5 * math.sqrt(p * big.size) + 1, (t, k, p * big.size)
    assert abs(float(big.mean())) < 2e-3 and abs(float(big.std()) - 1) < 2e-3


def test_random_bits_definition(orc):
    """Philox payload bits: bit j of frame g is bit j % 32 of word (j % 128) / 32 of block (j / 128, g); balanced."""
    b = orc.random_bits(3, 300, seed=0xA5A5 + (7 << 32), frame0=11)
    for f in range(3):
        for j in (0, 1, 31, 32, 127, 128, 129, 255, 256, 299):
            r = orc.philox4x32_10([j // 128, 0, 11 + f, 0], [0xA5A5, 7 ^ 0x62697473])
            assert b[f, j] == (r[(j % 128) // 32] >> (j % 32)) & 1
    big = orc.random_bits(64, 16384, seed=1)
    assert abs(float(big.mean()) - 0.5) < 0.002 and set(np.unique(big)) == {0, 1}


def test_packed_payload_definition(orc):
    """Packed payload rows (extension): first bit = most significant bit of its byte (np.packbits order, the order of
    bytes_to_bits, digital/util.rs:5-11), pad bits zero, unpack ignores them; only bit 0 of a bit byte counts."""
    rng = np.random.default_rng(3)
    for nbits in (1, 7, 8, 9, 64, 301):
        bits = rng.integers(0, 2, (5, nbits), dtype=np.uint8)
        packed = orc.pack_bits(bits)
        assert packed.shape == (5, (nbits + 7) // 8)
        assert np.array_equal(packed, np.packbits(bits, axis=1))
        assert np.array_equal(orc.unpack_bits(packed, nbits), bits)
        dirty = packed.copy()
        if nbits % 8:
            dirty[:, -1] |= (1 << (8 - nbits % 8)) - 1  # pad bits set: ignored
        assert np.array_equal(orc.unpack_bits(dirty, nbits), bits)
        assert np.array_equal(orc.pack_bits(bits | 0xFE), packed)
    # bytes_to_bits of one whole byte is the byte: 0b10110001 -> bits 1,0,1,1,0,0,0,1
    assert orc.unpack_bits(np.array([[0xB1]], np.uint8), 8).tolist() == [[1, 0, 1, 1, 0, 0, 0, 1]]
