"""ctypes binding of the CPU oracle (oracle/modem_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle/_build/liboracle.so with gcc (oracle/Makefile)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("modem_oracle.c", "modem_oracle.h"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return _SO


class Ring(C.Structure):
    _fields_ = [("start", C.c_uint8), ("end", C.c_uint8), ("radius", C.c_float), ("phase", C.c_float)]


class Phasor(C.Structure):
    _fields_ = [
        ("scheme", C.c_int), ("bits_per_symbol", C.c_size_t), ("amplitude", C.c_float), ("phase", C.c_float),
        ("phase_cos", C.c_float), ("phase_sin", C.c_float), ("max_symbol", C.c_float),
        ("bits_per_carrier", C.c_size_t), ("num_symbols", C.c_float), ("even", C.c_int),
        ("rings", Ring * 8), ("n_rings", C.c_size_t), ("deviation", C.c_float), ("prev", C.c_uint8),
        ("cur_coef", C.c_float), ("max_symbol_i", C.c_int), ("shift", C.c_float), ("samples_per_bit", C.c_size_t),
    ]


class SymbolClock(C.Structure):
    _fields_ = [("samples_per_symbol", C.c_size_t), ("counter", C.c_size_t)]


class Update(C.Structure):
    _fields_ = [("kind", C.c_int), ("bits", C.POINTER(C.c_uint8)), ("len", C.c_size_t)]


class Source(C.Structure):
    _fields_ = [
        ("bits", C.POINTER(C.c_uint8)), ("nbits", C.c_size_t), ("clock", SymbolClock),
        ("bits_per_symbol", C.c_size_t), ("idx", C.c_size_t), ("evenodd", C.c_int),
        ("half_clock", SymbolClock), ("cur", C.c_uint8 * 2),
    ]


class Carrier(C.Structure):
    _fields_ = [("sample_freq", C.c_float), ("sample", C.c_size_t)]


class Fir(C.Structure):
    _fields_ = [("coefs", C.POINTER(C.c_float)), ("n", C.c_size_t), ("history", C.POINTER(C.c_float)),
                ("idx", C.c_size_t)]


class Pll(C.Structure):
    _fields_ = [("phase_offset", C.c_float)]


class Path(C.Structure):
    _fields_ = [
        ("scheme", C.c_char * 16), ("baud_rate", C.c_size_t), ("sample_rate", C.c_size_t),
        ("carrier_hz", C.c_size_t), ("sample0", C.c_size_t),
        ("tx_taps", C.POINTER(C.c_float)), ("n_tx_taps", C.c_size_t),
        ("rx_taps", C.POINTER(C.c_float)), ("n_rx_taps", C.c_size_t),
        ("phase_offset", C.c_float), ("decision_delay", C.c_size_t), ("slicer_gain", C.c_float),
    ]


CHANGED, UNCHANGED, FINISHED = 0, 1, 2
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        f32, sz, u8p, f32p = C.c_float, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L = _lib
        L.orc_mod_trig.restype = f32; L.orc_mod_trig.argtypes = [f32]
        L.orc_ang_freq.restype = f32; L.orc_ang_freq.argtypes = [sz]
        L.orc_sample_freq.restype = f32; L.orc_sample_freq.argtypes = [sz, sz]
        L.orc_samples_per_symbol.restype = sz; L.orc_samples_per_symbol.argtypes = [sz, sz]
        L.orc_carrier_new.argtypes = [C.POINTER(Carrier), sz, sz]
        L.orc_carrier_next.restype = f32; L.orc_carrier_next.argtypes = [C.POINTER(Carrier)]
        L.orc_symbol_clock_new.argtypes = [C.POINTER(SymbolClock), sz]
        L.orc_symbol_clock_next.restype = C.c_int; L.orc_symbol_clock_next.argtypes = [C.POINTER(SymbolClock)]
        L.orc_bits_new.argtypes = [C.POINTER(Source), u8p, sz, sz, sz]
        L.orc_evenodd_new.argtypes = [C.POINTER(Source), u8p, sz, sz, sz]
        L.orc_source_next.restype = Update; L.orc_source_next.argtypes = [C.POINTER(Source)]
        L.orc_bit_to_sign.restype = f32; L.orc_bit_to_sign.argtypes = [C.c_uint8]
        L.orc_bytes_to_bits.restype = C.c_uint8; L.orc_bytes_to_bits.argtypes = [u8p, sz]
        L.orc_max_symbol.restype = sz; L.orc_max_symbol.argtypes = [sz]
        PP = C.POINTER(Phasor)
        L.orc_bask_new.argtypes = [PP, f32]
        L.orc_bpsk_new.argtypes = [PP, f32, f32]
        L.orc_qpsk_new.argtypes = [PP, f32, f32]
        L.orc_qam_new.argtypes = [PP, sz, f32, f32]
        L.orc_mpsk_new.argtypes = [PP, sz, f32, f32]
        L.orc_oqpsk_new.argtypes = [PP, f32]
        L.orc_dcqpsk_new.argtypes = [PP, f32]
        L.orc_apsk_new.restype = C.c_int; L.orc_apsk_new.argtypes = [PP, f32, sz, C.POINTER(Ring), sz]
        L.orc_bfsk_new.argtypes = [PP, sz, sz, f32]
        L.orc_mfsk_new.argtypes = [PP, sz, sz, sz, f32, C.c_int]
        L.orc_cpfsk_new.argtypes = [PP, sz, sz, sz, f32, sz]
        L.orc_msk_new.argtypes = [PP, f32, sz]
        L.orc_dmpsk_new.argtypes = [PP, sz, f32, f32, f32]
        L.orc_phasor_by_name.restype = C.c_int; L.orc_phasor_by_name.argtypes = [PP, C.c_char_p, sz, sz]
        L.orc_phasor_update.argtypes = [PP, sz, u8p]
        L.orc_phasor_i.restype = f32; L.orc_phasor_i.argtypes = [PP, sz, u8p]
        L.orc_phasor_q.restype = f32; L.orc_phasor_q.argtypes = [PP, sz, u8p]
        L.orc_fir_new.restype = C.c_int; L.orc_fir_new.argtypes = [C.POINTER(Fir), f32p, sz]
        L.orc_fir_add.restype = f32; L.orc_fir_add.argtypes = [C.POINTER(Fir), f32]
        L.orc_fir_free.argtypes = [C.POINTER(Fir)]
        L.orc_pll_handle.argtypes = [C.POINTER(Pll), f32, f32, f32]
        L.orc_lowpass_taps.restype = f32p; L.orc_lowpass_taps.argtypes = [C.POINTER(sz)]
        L.orc_hilbert_taps.restype = f32p; L.orc_hilbert_taps.argtypes = [C.POINTER(sz)]
        L.orc_rrc_taps.argtypes = [f32p, sz, sz, C.c_double]
        u32p = C.POINTER(C.c_uint32)
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_box_muller.argtypes = [C.c_uint32, C.c_uint32, f32p, f32p]
        L.orc_awgn.argtypes = [f32p, sz, sz, f32, C.c_uint64, C.c_uint64]
        L.orc_random_bits.argtypes = [u8p, sz, sz, C.c_uint64, C.c_uint64]
        L.orc_unpack_bits.argtypes = [u8p, u8p, sz, sz]
        L.orc_pack_bits.argtypes = [u8p, u8p, sz, sz]
        L.orc_sigma_for_ebn0.restype = f32
        L.orc_sigma_for_ebn0.argtypes = [f32p, sz, sz, f32, f32, f32p, sz, C.c_double]
        PA = C.POINTER(Path)
        L.orc_frame_samples.restype = sz; L.orc_frame_samples.argtypes = [PA, sz]
        L.orc_decided_symbols.restype = sz; L.orc_decided_symbols.argtypes = [PA, sz]
        L.orc_bits_per_symbol.restype = sz; L.orc_bits_per_symbol.argtypes = [PA]
        L.orc_constellation.restype = sz; L.orc_constellation.argtypes = [PA, f32p, sz]
        L.orc_modulate.restype = C.c_int; L.orc_modulate.argtypes = [PA, u8p, sz, sz, f32p, f32p]
        L.orc_demodulate.restype = C.c_int; L.orc_demodulate.argtypes = [PA, f32p, sz, sz, f32p, u8p, u8p]
        L.orc_preamble.restype = C.c_int; L.orc_preamble.argtypes = [PA, sz, sz, f32, f32p]
        L.orc_modulate_real.restype = C.c_int; L.orc_modulate_real.argtypes = [PA, u8p, sz, sz, sz, f32, f32p]
        L.orc_demodulate_real.restype = C.c_int
        L.orc_demodulate_real.argtypes = [PA, f32p, f32p, sz, sz, sz, f32p, sz, f32p, f32p, u8p, u8p]
        L.orc_loopback.restype = C.c_int
        L.orc_loopback.argtypes = [PA, u8p, sz, sz, f32, C.c_uint64, C.c_uint64, C.c_int, u8p, u8p,
                                   C.POINTER(C.c_uint64)]
    return _lib


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def lowpass_taps():
    n = C.c_size_t()
    p = lib().orc_lowpass_taps(C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def hilbert_taps():
    n = C.c_size_t()
    p = lib().orc_hilbert_taps(C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def rrc_taps(span, sps, beta):
    out = np.empty(span * sps + 1, np.float32)
    lib().orc_rrc_taps(_f32p(out), span, sps, beta)
    return out


def random_bits(F, nbits, seed, frame0=0):
    """Philox payload bits (extension, modem_oracle.h): [F][nbits] bytes 0/1."""
    out = np.zeros((F, nbits), np.uint8)
    lib().orc_random_bits(_u8p(out), F, nbits, seed, frame0)
    return out


def unpack_bits(packed, nbits):
    """Packed payload rows (extension, modem_oracle.h) -> [F][nbits] bytes 0/1."""
    packed = np.ascontiguousarray(packed, np.uint8)
    assert packed.shape[1] == (nbits + 7) // 8
    out = np.zeros((packed.shape[0], nbits), np.uint8)
    lib().orc_unpack_bits(_u8p(packed), _u8p(out), packed.shape[0], nbits)
    return out


def pack_bits(bits):
    """[F][nbits] bytes 0/1 -> packed payload rows [F][ceil(nbits/8)], pad bits zero."""
    bits = np.ascontiguousarray(bits, np.uint8)
    F, nbits = bits.shape
    out = np.zeros((F, (nbits + 7) // 8), np.uint8)
    lib().orc_pack_bits(_u8p(bits), _u8p(out), F, nbits)
    return out


def box_muller(r0, r1):
    a, b = C.c_float(), C.c_float()
    lib().orc_box_muller(r0, r1, C.byref(a), C.byref(b))
    return a.value, b.value


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


class OraclePath:
    """One configured modulate->demodulate path (mirrors what a src/bin caller composes)."""

    def __init__(self, scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, sample0=0,
                 tx_taps=None, rx_taps=None, phase_offset=0.0, decision_delay=0, slicer_gain=1.0):
        self.p = Path()
        self.p.scheme = scheme.encode()
        self.p.baud_rate, self.p.sample_rate, self.p.carrier_hz, self.p.sample0 = baud_rate, sample_rate, carrier_hz, sample0
        self._tx = None if tx_taps is None or len(tx_taps) == 0 else np.ascontiguousarray(tx_taps, np.float32)
        self._rx = np.ascontiguousarray(rx_taps if rx_taps is not None else lowpass_taps(), np.float32)
        self.p.tx_taps = _f32p(self._tx); self.p.n_tx_taps = 0 if self._tx is None else len(self._tx)
        self.p.rx_taps = _f32p(self._rx); self.p.n_rx_taps = len(self._rx)
        self.p.phase_offset = phase_offset
        self.p.decision_delay = decision_delay
        self.p.slicer_gain = slicer_gain
        self.bps = lib().orc_bits_per_symbol(C.byref(self.p))
        if self.bps == 0:
            raise ValueError("invalid digital modulation")  # modulate.rs:94
        self.sps = lib().orc_samples_per_symbol(baud_rate, sample_rate)

    def frame_samples(self, nbits):
        return lib().orc_frame_samples(C.byref(self.p), nbits)

    def decided_symbols(self, L):
        return lib().orc_decided_symbols(C.byref(self.p), L)

    def constellation(self):
        out = np.zeros((512, 2), np.float32)
        nt = lib().orc_constellation(C.byref(self.p), _f32p(out), 512)
        return out[: nt * (1 << self.bps)].reshape(nt, 1 << self.bps, 2).copy()

    def modulate(self, bits, want_iq=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        L = self.frame_samples(nbits)
        tx = np.zeros((F, L, 2), np.float32)
        iq = np.zeros((F, L, 2), np.float32) if want_iq else None
        rc = lib().orc_modulate(C.byref(self.p), _u8p(bits), F, nbits, _f32p(tx), _f32p(iq))
        assert rc == 0, rc
        return (tx, iq) if want_iq else tx

    def demodulate(self, rx, want_filt=True):
        rx = np.ascontiguousarray(rx, np.float32)
        F, L, _ = rx.shape
        K = self.decided_symbols(L)
        filt = np.zeros((F, L, 2), np.float32) if want_filt else None
        sym = np.zeros((F, K), np.uint8)
        bits = np.zeros((F, K * self.bps), np.uint8)
        rc = lib().orc_demodulate(C.byref(self.p), _f32p(rx), F, L, _f32p(filt), _u8p(sym), _u8p(bits))
        assert rc == 0, rc
        return filt, sym, bits

    def awgn(self, buf, sigma, seed, frame0=0):
        buf = np.ascontiguousarray(buf, np.float32).copy()
        F, L, _ = buf.shape
        lib().orc_awgn(_f32p(buf), F, L, sigma, seed, frame0)
        return buf

    def sigma_for_ebn0(self, ebn0_db, rx_gain=2.0):
        c = self.constellation()[0]
        return lib().orc_sigma_for_ebn0(_f32p(c), len(c), self.bps, self.p.slicer_gain, rx_gain,
                                        _f32p(self._rx), len(self._rx), ebn0_db)

    def loopback(self, bits, sigma=0.0, seed=0, frame0=0, threads=1, want_out=True):
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        L = self.frame_samples(nbits)
        K = self.decided_symbols(L)
        sym = np.zeros((F, K), np.uint8) if want_out else None
        bo = np.zeros((F, K * self.bps), np.uint8) if want_out else None
        cnt = (C.c_uint64 * 2)(0, 0)
        rc = lib().orc_loopback(C.byref(self.p), _u8p(bits), F, nbits, sigma, seed, frame0, threads,
                                _u8p(sym), _u8p(bo), cnt)
        assert rc == 0, rc
        return sym, bo, (cnt[0], cnt[1])

    # ---- src/bin sample paths -------------------------------------------------------------
    def preamble(self, F, n, amplitude=1.0):
        tx = np.zeros((F, n, 2), np.float32)
        assert lib().orc_preamble(C.byref(self.p), F, n, amplitude, _f32p(tx)) == 0
        return tx

    def modulate_real(self, bits, preamble=0, preamble_amplitude=1.0):
        """What src/bin/modulate.rs writes without --iq: [F][preamble + L] f32."""
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        out = np.zeros((F, preamble + self.frame_samples(nbits)), np.float32)
        rc = lib().orc_modulate_real(C.byref(self.p), _u8p(bits), F, nbits, preamble, preamble_amplitude, _f32p(out))
        assert rc == 0, rc
        return out

    def demodulate_real(self, x, lock=64, analytic_im=None, hilbert=None, want_filt=True):
        """src/bin/demodulate.rs per frame: x [F][L] real (any numeric dtype, `x as f32`).  Returns
        (phase_offset [F], filt [F][L-lock][2], sym, bits)."""
        x = np.ascontiguousarray(x, np.float32)
        F, L = x.shape
        Lr = L - lock
        K = self.decided_symbols(Lr)
        po = np.zeros(F, np.float32)
        filt = np.zeros((F, Lr, 2), np.float32) if want_filt else None
        sym = np.zeros((F, K), np.uint8)
        bits = np.zeros((F, K * self.bps), np.uint8)
        im = None if analytic_im is None else np.ascontiguousarray(analytic_im, np.float32)
        h = None if hilbert is None else np.ascontiguousarray(hilbert, np.float32)
        rc = lib().orc_demodulate_real(C.byref(self.p), _f32p(x), _f32p(im), F, L, lock, _f32p(h), 0 if h is None else len(h),
                                       _f32p(po), _f32p(filt), _u8p(sym), _u8p(bits))
        if rc == -3:
            raise ValueError("called `Option::unwrap()` on a `None` value")  # demodulator.rs:34
        assert rc == 0, rc
        return po, filt, sym, bits
