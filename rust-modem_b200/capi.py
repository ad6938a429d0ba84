"""ctypes binding of include/modem_gpu.h (libmodem_gpu.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("MODEM_GPU_LIB") or os.path.join(_HERE, "lib", "libmodem_gpu.so")  # MODEM_GPU_LIB: a tuning build (tools/)

FLAG_FUSED_MAC = 0x1
FLAG_NO_TMEM = 0x2
COMM_ID_BYTES = 128


class ModemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"modem_gpu error {code}: {msg}")
        self.code = code


class ModemCfg(C.Structure):
    """modem_cfg_t (include/modem_gpu.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("bits_per_symbol", C.c_uint32), ("samples_per_symbol", C.c_uint32),
        ("n_tables", C.c_uint32), ("const_iq", C.POINTER(C.c_float)), ("q_offset", C.c_uint32),
        ("sample_freq", C.c_float), ("phase_offset", C.c_float), ("sample0", C.c_uint64),
        ("n_tx_taps", C.c_uint32), ("tx_taps", C.POINTER(C.c_float)),
        ("n_rx_taps", C.c_uint32), ("rx_taps", C.POINTER(C.c_float)),
        ("decision_delay", C.c_uint32), ("rx_gain", C.c_float), ("slicer_gain", C.c_float), ("flags", C.c_uint32),
    ]


class Phasor(C.Structure):
    """modem_phasor_t (include/modem_gpu.h): the stateful / time-varying mappers."""
    _fields_ = [("struct_size", C.c_uint32), ("kind", C.c_uint32), ("bits_per_symbol", C.c_uint32),
                ("amplitude", C.c_float), ("deviation", C.c_float), ("phase", C.c_float), ("shift", C.c_float),
                ("mfsk_increase_map", C.c_uint32)]


PHASOR_TABLE, PHASOR_BFSK, PHASOR_MFSK, PHASOR_CPFSK, PHASOR_MSK, PHASOR_DMPSK = range(6)
SAMPLES_C32, SAMPLES_F32, SAMPLES_I16 = 0, 1, 2
STATEFUL_NAMES = ("bfsk", "mfsk", "16cpfsk", "msk", "dqpsk", "dbpsk")


class Ring(C.Structure):
    _fields_ = [("start", C.c_uint8), ("end", C.c_uint8), ("radius", C.c_float), ("phase", C.c_float)]


def library_path():
    return _SO


def build_library(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a (csrc/Makefile); cross-compiles without a GPU."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(_HERE, "..", "include", "modem_gpu.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["make", "-C", csrc, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    """Load libmodem_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ModemError(-6, f"{_SO} is not built: run __graft_entry__.build() (nvcc, sm_100a). "
                             "There is no CPU fallback.")
    L = C.CDLL(_SO)
    vp, sz, f32, u8p, f32p = C.c_void_p, C.c_size_t, C.c_float, C.c_void_p, C.POINTER(C.c_float)
    u64, u32p = C.c_uint64, C.POINTER(C.c_uint32)
    L.modem_sample_freq.restype = f32; L.modem_sample_freq.argtypes = [sz, sz]
    L.modem_samples_per_symbol.restype = sz; L.modem_samples_per_symbol.argtypes = [sz, sz]
    L.modem_const_bask.argtypes = [f32, f32p]
    L.modem_const_bpsk.argtypes = [f32, f32, f32p]
    L.modem_const_qpsk.argtypes = [f32, f32, f32p]
    L.modem_const_qam.argtypes = [C.c_uint32, f32, f32, f32p]
    L.modem_const_mpsk.argtypes = [C.c_uint32, f32, f32, f32p]
    L.modem_const_oqpsk.argtypes = [f32, f32p]
    L.modem_const_dcqpsk.argtypes = [f32, f32p]
    L.modem_const_apsk.argtypes = [f32, C.c_uint32, C.POINTER(Ring), sz, f32p]
    L.modem_const_by_name.argtypes = [C.c_char_p, f32p, u32p, u32p]
    L.modem_lowpass_taps.restype = f32p; L.modem_lowpass_taps.argtypes = [C.POINTER(sz)]
    L.modem_rrc_taps.argtypes = [f32p, sz, sz, C.c_double]
    L.modem_hilbert_taps.restype = f32p; L.modem_hilbert_taps.argtypes = [C.POINTER(sz)]
    L.modem_phasor_by_name.argtypes = [C.c_char_p, sz, sz, C.POINTER(Phasor), u32p]
    L.modem_gpu_set_phasor.argtypes = [vp, C.POINTER(Phasor)]
    L.modem_gpu_preamble.argtypes = [vp, sz, sz, f32, vp]
    L.modem_gpu_modulate_real.argtypes = [vp, u8p, sz, sz, sz, f32, vp]
    L.modem_gpu_lock_phase.argtypes = [vp, vp, C.c_uint32, sz, sz, f32p, sz, sz, vp]
    L.modem_gpu_demodulate_real.argtypes = [vp, vp, C.c_uint32, sz, sz, sz, f32p, sz, vp, vp, vp, vp, vp]
    L.modem_sigma_for_ebn0.restype = f32; L.modem_sigma_for_ebn0.argtypes = [C.POINTER(ModemCfg), C.c_double]
    L.modem_gpu_device_count.argtypes = [C.POINTER(C.c_int)]
    L.modem_gpu_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(ModemCfg)]
    L.modem_gpu_destroy.restype = None; L.modem_gpu_destroy.argtypes = [vp]
    L.modem_gpu_set_stream.argtypes = [vp, vp]
    L.modem_gpu_set_channels.argtypes = [vp, sz, f32p, f32p, sz]
    L.modem_gpu_synchronize.argtypes = [vp]
    L.modem_gpu_frame_samples.restype = sz; L.modem_gpu_frame_samples.argtypes = [vp, sz]
    L.modem_gpu_decided_symbols.restype = sz; L.modem_gpu_decided_symbols.argtypes = [vp, sz]
    L.modem_gpu_modulate.argtypes = [vp, u8p, sz, sz, vp, vp]
    L.modem_gpu_awgn.argtypes = [vp, vp, sz, sz, f32, u64, u64]
    L.modem_gpu_random_bits.argtypes = [vp, vp, sz, sz, u64, u64]
    L.modem_gpu_demodulate.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp, f32, u64, u64]
    L.modem_gpu_demodulate_count.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, vp, f32, u64, u64]
    L.modem_gpu_ber_sweep.argtypes = [vp, vp, sz, sz, sz, f32p, u64, u64, vp, vp]
    L.modem_gpu_loopback_device.argtypes = [vp, vp, sz, sz, f32, u64, u64, vp, vp, vp, vp]
    L.modem_gpu_loopback.argtypes = [vp, u8p, sz, sz, f32, u64, u64, vp, vp, vp, C.POINTER(u64)]
    L.modem_gpu_loopback_packed.argtypes = [vp, u8p, sz, sz, f32, u64, u64, vp, C.POINTER(u64)]
    L.modem_gpu_malloc.argtypes = [vp, C.POINTER(vp), sz]
    L.modem_gpu_free.argtypes = [vp, vp]
    L.modem_gpu_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.modem_gpu_host_free.argtypes = [vp]
    L.modem_gpu_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.modem_gpu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.modem_gpu_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.modem_gpu_comm_create.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.modem_gpu_allreduce_counters.argtypes = [vp, vp, sz]
    L.modem_gpu_comm_destroy.restype = None; L.modem_gpu_comm_destroy.argtypes = [vp]
    L.modem_gpu_strerror.restype = C.c_char_p; L.modem_gpu_strerror.argtypes = [C.c_int]
    L.modem_gpu_last_error.restype = C.c_char_p; L.modem_gpu_last_error.argtypes = [vp]
    L.modem_gpu_launch_count.restype = u64; L.modem_gpu_launch_count.argtypes = [vp]
    L.modem_gpu_abi_version.restype = C.c_int
    _lib = L
    return L


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def sample_freq(hz, sr):
    return lib().modem_sample_freq(hz, sr)


def samples_per_symbol(br, sr):
    return lib().modem_samples_per_symbol(br, sr)


def lowpass_taps():
    n = C.c_size_t()
    p = lib().modem_lowpass_taps(C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def hilbert_taps():
    n = C.c_size_t()
    p = lib().modem_hilbert_taps(C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def host_phasor(name, baud_rate, sample_rate):
    """(Phasor, bps, evenodd) for a stateful src/bin/modulate.rs -m name (bfsk, mfsk, 16cpfsk, msk, dqpsk, dbpsk)."""
    ph, eo = Phasor(), C.c_uint32()
    bps = lib().modem_phasor_by_name(name.encode(), baud_rate, sample_rate, C.byref(ph), C.byref(eo))
    if bps < 0:
        raise ModemError(bps, f"invalid digital modulation {name!r}")
    return ph, bps, bool(eo.value)


def rrc_taps(span, sps, beta):
    out = np.empty(span * sps + 1, np.float32)
    rc = lib().modem_rrc_taps(_f32p(out), span, sps, beta)
    if rc:
        raise ModemError(rc, "modem_rrc_taps")
    return out


def host_constellation(name):
    """(table [n_tables][2^bps][2], bps, evenodd) for a src/bin/modulate.rs -m name."""
    out = np.zeros((512, 2), np.float32)
    nt, eo = C.c_uint32(), C.c_uint32()
    bps = lib().modem_const_by_name(name.encode(), _f32p(out), C.byref(nt), C.byref(eo))
    if bps < 0:
        # modulate.rs:94 panics with "invalid digital modulation"
        raise ModemError(bps, f"invalid digital modulation {name!r}")
    np_ = 1 << bps
    return out[: nt.value * np_].reshape(nt.value, np_, 2).copy(), bps, bool(eo.value)


def _ptr(x):
    """Raw address of a numpy array (host) or torch tensor (device or host); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


class Modem:
    """One configured path on one GPU (wraps modem_ctx_t).

    Parameters mirror what a src/bin caller passes to Rates::new / Carrier::new(Freq::new) /
    the digital::* constructor / Demodulator::new.
    """

    def __init__(self, scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, sample0=0,
                 tx_taps=None, rx_taps=None, phase_offset=0.0, decision_delay=0, slicer_gain=1.0, rx_gain=2.0,
                 flags=0, device=0, const_iq=None, bps=None, q_offset=None):
        L = lib()
        phasor = None
        if const_iq is None and scheme in STATEFUL_NAMES:
            # TX-only schemes: the context still needs a (dummy) table of the right size for its slicer
            phasor, bps, evenodd = host_phasor(scheme, baud_rate, sample_rate)
            const_iq = np.zeros((1, 1 << bps, 2), np.float32)
        elif const_iq is None:
            const_iq, bps, evenodd = host_constellation(scheme)
        else:
            const_iq = np.ascontiguousarray(const_iq, np.float32)
            evenodd = False
        self.bps = bps
        self.sps = samples_per_symbol(baud_rate, sample_rate)
        self._const = np.ascontiguousarray(const_iq, np.float32)
        self._tx = None if tx_taps is None or len(tx_taps) == 0 else np.ascontiguousarray(tx_taps, np.float32)
        self._rx = np.ascontiguousarray(rx_taps if rx_taps is not None else lowpass_taps(), np.float32)
        cfg = ModemCfg()
        cfg.struct_size = C.sizeof(ModemCfg)
        cfg.bits_per_symbol = bps
        cfg.samples_per_symbol = self.sps
        cfg.n_tables = self._const.shape[0]
        cfg.const_iq = _f32p(self._const)
        cfg.q_offset = (self.sps // 2 if evenodd else 0) if q_offset is None else q_offset
        cfg.sample_freq = sample_freq(carrier_hz, sample_rate)
        cfg.phase_offset = phase_offset
        cfg.sample0 = sample0
        cfg.n_tx_taps = 0 if self._tx is None else len(self._tx)
        cfg.tx_taps = _f32p(self._tx)
        cfg.n_rx_taps = len(self._rx)
        cfg.rx_taps = _f32p(self._rx)
        cfg.decision_delay = decision_delay
        cfg.rx_gain = rx_gain
        cfg.slicer_gain = slicer_gain
        cfg.flags = flags
        self.cfg = cfg
        self._ctx = C.c_void_p()
        rc = L.modem_gpu_create(C.byref(self._ctx), device, C.byref(cfg))
        if rc:
            raise ModemError(rc, (L.modem_gpu_last_error(None) or b"").decode() or L.modem_gpu_strerror(rc).decode())
        self.phasor = phasor
        if phasor is not None:
            self._ck(L.modem_gpu_set_phasor(self._ctx, C.byref(phasor)))

    # -- plumbing
    def _ck(self, rc):
        if rc:
            L = lib()
            raise ModemError(rc, (L.modem_gpu_last_error(self._ctx) or b"").decode() or L.modem_gpu_strerror(rc).decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib().modem_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        self._ck(lib().modem_gpu_set_stream(self._ctx, cuda_stream_handle))

    def set_channels(self, sample_freqs, frames_per_channel, phase_offsets=None):
        w = np.ascontiguousarray(sample_freqs, np.float32)
        po = None if phase_offsets is None else np.ascontiguousarray(phase_offsets, np.float32)
        self._ck(lib().modem_gpu_set_channels(self._ctx, len(w), _f32p(w), _f32p(po), frames_per_channel))

    def synchronize(self):
        self._ck(lib().modem_gpu_synchronize(self._ctx))

    def frame_samples(self, nbits):
        return lib().modem_gpu_frame_samples(self._ctx, nbits)

    def decided_symbols(self, L):
        return lib().modem_gpu_decided_symbols(self._ctx, L)

    def sigma_for_ebn0(self, ebn0_db):
        return lib().modem_sigma_for_ebn0(C.byref(self.cfg), ebn0_db)

    @property
    def launch_count(self):
        return lib().modem_gpu_launch_count(self._ctx)

    # -- raw-pointer entry points (numpy host arrays or torch tensors)
    def modulate_into(self, bits, F, nbits, tx=None, iq=None):
        self._ck(lib().modem_gpu_modulate(self._ctx, _ptr(bits), F, nbits, _ptr(tx), _ptr(iq)))

    def random_bits_into(self, bits, F, nbits, seed, frame0=0):
        """Philox payload bits generated on the device (host or device destination)."""
        self._ck(lib().modem_gpu_random_bits(self._ctx, _ptr(bits), F, nbits, seed, frame0))

    def random_bits(self, F, nbits, seed, frame0=0):
        out = np.zeros((F, nbits), np.uint8)
        self.random_bits_into(out, F, nbits, seed, frame0)
        return out

    def awgn_inplace(self, buf, F, L, sigma, seed, frame0=0):
        self._ck(lib().modem_gpu_awgn(self._ctx, _ptr(buf), F, L, sigma, seed, frame0))

    def demodulate_into(self, rx, F, L, sym=None, bits=None, soft=None, filt=None, sigma=0.0, seed=0, frame0=0):
        self._ck(lib().modem_gpu_demodulate(self._ctx, _ptr(rx), F, L, _ptr(sym), _ptr(bits), _ptr(soft), _ptr(filt),
                                            sigma, seed, frame0))

    def demodulate_count_into(self, rx, F, L, ref_bits, ref_stride, counters, sym=None, bits=None, sigma=0.0, seed=0,
                              frame0=0):
        """Stream-ordered: device pointers only, accumulates into device counters u64[2]."""
        self._ck(lib().modem_gpu_demodulate_count(self._ctx, _ptr(rx), F, L, _ptr(sym), _ptr(bits), _ptr(ref_bits),
                                                  ref_stride, _ptr(counters), sigma, seed, frame0))

    def ber_sweep_into(self, bits, F, nbits, sigmas, counters, seed=0, frame0=0, tx=None):
        """Modulate once, then one noisy RX pass per sigma; device counters [len(sigmas)][2] are accumulated."""
        sg = np.ascontiguousarray(sigmas, np.float32)
        self._ck(lib().modem_gpu_ber_sweep(self._ctx, _ptr(bits), F, nbits, len(sg), _f32p(sg), seed, frame0, _ptr(tx),
                                           _ptr(counters)))

    def loopback_device_into(self, bits, F, nbits, counters, tx=None, sym=None, bits_out=None, sigma=0.0, seed=0, frame0=0):
        """Stream-ordered chunked TX||RX pipeline on device buffers; accumulates into device counters u64[2]."""
        self._ck(lib().modem_gpu_loopback_device(self._ctx, _ptr(bits), F, nbits, sigma, seed, frame0, _ptr(tx), _ptr(sym),
                                                 _ptr(bits_out), _ptr(counters)))

    def loopback_into(self, bits, F, nbits, sigma=0.0, seed=0, frame0=0, tx=None, sym=None, bits_out=None):
        cnt = (C.c_uint64 * 2)(0, 0)
        self._ck(lib().modem_gpu_loopback(self._ctx, _ptr(bits), F, nbits, sigma, seed, frame0, _ptr(tx), _ptr(sym),
                                          _ptr(bits_out), cnt))
        return cnt[0], cnt[1]

    def loopback_packed_into(self, packed, F, nbits, packed_out=None, sigma=0.0, seed=0, frame0=0):
        """Extension: the loopback on packed payloads (8 bits per byte, first bit = MSB; np.packbits order)."""
        cnt = (C.c_uint64 * 2)(0, 0)
        self._ck(lib().modem_gpu_loopback_packed(self._ctx, _ptr(packed), F, nbits, sigma, seed, frame0, _ptr(packed_out), cnt))
        return cnt[0], cnt[1]

    # -- numpy conveniences (host buffers through the C ABI)
    def modulate(self, bits, want_iq=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        L = self.frame_samples(nbits)
        tx = np.zeros((F, L, 2), np.float32)
        iq = np.zeros((F, L, 2), np.float32) if want_iq else None
        self.modulate_into(bits, F, nbits, tx, iq)
        return (tx, iq) if want_iq else tx

    def demodulate(self, rx, want_filt=False, want_soft=False, sigma=0.0, seed=0, frame0=0):
        rx = np.ascontiguousarray(rx, np.float32)
        F, L, _ = rx.shape
        K = self.decided_symbols(L)
        sym = np.zeros((F, K), np.uint8)
        bits = np.zeros((F, K * self.bps), np.uint8)
        soft = np.zeros((F, K, 2), np.float32) if want_soft else None
        filt = np.zeros((F, L, 2), np.float32) if want_filt else None
        self.demodulate_into(rx, F, L, sym, bits, soft, filt, sigma, seed, frame0)
        return {"sym": sym, "bits": bits, "soft": soft, "filt": filt}

    def awgn(self, buf, sigma, seed, frame0=0):
        buf = np.ascontiguousarray(buf, np.float32).copy()
        F, L, _ = buf.shape
        self.awgn_inplace(buf, F, L, sigma, seed, frame0)
        return buf

    # -- src/bin sample paths
    def preamble(self, F, n, amplitude=1.0):
        tx = np.zeros((F, n, 2), np.float32)
        self._ck(lib().modem_gpu_preamble(self._ctx, F, n, amplitude, _ptr(tx)))
        return tx

    def modulate_real(self, bits, preamble=0, preamble_amplitude=1.0):
        """The f32 stream src/bin/modulate.rs writes without --iq: [F][preamble + L]."""
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        out = np.zeros((F, preamble + self.frame_samples(nbits)), np.float32)
        self._ck(lib().modem_gpu_modulate_real(self._ctx, _ptr(bits), F, nbits, preamble, preamble_amplitude, _ptr(out)))
        return out

    @staticmethod
    def _fmt(x):
        if x.dtype == np.int16:
            return SAMPLES_I16
        if x.dtype == np.float32 and x.ndim == 3:
            return SAMPLES_C32
        if x.dtype == np.float32:
            return SAMPLES_F32
        raise TypeError(x.dtype)

    def lock_phase(self, x, lock=64, hilbert=None):
        x = np.ascontiguousarray(x)
        F, L = x.shape[:2]
        po = np.zeros(F, np.float32)
        h = None if hilbert is None else np.ascontiguousarray(hilbert, np.float32)
        self._ck(lib().modem_gpu_lock_phase(self._ctx, _ptr(x), self._fmt(x), F, L, _f32p(h), 0 if h is None else len(h),
                                            lock, _ptr(po)))
        return po

    def demodulate_real(self, x, lock=64, hilbert=None, want_filt=True, want_soft=False):
        """src/bin/demodulate.rs per frame: x [F][L] int16 / float32 real samples (or [F][L][2] analytic)."""
        x = np.ascontiguousarray(x)
        F, L = x.shape[:2]
        Lr = L - lock
        K = self.decided_symbols(Lr) if Lr > 0 else 0
        po = np.zeros(F, np.float32)
        sym = np.zeros((F, K), np.uint8)
        bits = np.zeros((F, K * self.bps), np.uint8)
        soft = np.zeros((F, K, 2), np.float32) if want_soft else None
        filt = np.zeros((F, max(Lr, 0), 2), np.float32) if want_filt else None
        h = None if hilbert is None else np.ascontiguousarray(hilbert, np.float32)
        self._ck(lib().modem_gpu_demodulate_real(self._ctx, _ptr(x), self._fmt(x), F, L, lock, _f32p(h),
                                                 0 if h is None else len(h), _ptr(po), _ptr(sym), _ptr(bits),
                                                 _ptr(soft), _ptr(filt)))
        return {"phase_offset": po, "sym": sym, "bits": bits, "soft": soft, "filt": filt}

    def loopback(self, bits, sigma=0.0, seed=0, frame0=0, want_tx=False):
        bits = np.ascontiguousarray(bits, np.uint8)
        F, nbits = bits.shape
        L = self.frame_samples(nbits)
        K = self.decided_symbols(L)
        sym = np.zeros((F, K), np.uint8)
        bo = np.zeros((F, K * self.bps), np.uint8)
        tx = np.zeros((F, L, 2), np.float32) if want_tx else None
        cnt = self.loopback_into(bits, F, nbits, sigma, seed, frame0, tx, sym, bo)
        return {"sym": sym, "bits": bo, "tx": tx, "errors": cnt[0], "compared": cnt[1]}


    def loopback_packed(self, packed, nbits, sigma=0.0, seed=0, frame0=0):
        """packed [F][ceil(nbits/8)] uint8 -> {'packed': [F][ceil(K*bps/8)], 'errors', 'compared'}."""
        packed = np.ascontiguousarray(packed, np.uint8)
        F = packed.shape[0]
        assert packed.shape[1] == (nbits + 7) // 8
        K = self.decided_symbols(self.frame_samples(nbits))
        out = np.zeros((F, (K * self.bps + 7) // 8), np.uint8)
        cnt = self.loopback_packed_into(packed, F, nbits, out, sigma, seed, frame0)
        return {"packed": out, "errors": cnt[0], "compared": cnt[1]}


class Comm:
    """The single NCCL all-reduce of error counters (modem_comm_t)."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        rc = lib().modem_gpu_comm_unique_id(buf)
        if rc:
            raise ModemError(rc, (lib().modem_gpu_last_error(None) or b"").decode())
        return bytes(buf)

    def __init__(self, modem, n_ranks, rank, uid):
        self._modem = modem
        self._c = C.c_void_p()
        idb = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(uid)
        modem._ck(lib().modem_gpu_comm_create(C.byref(self._c), modem._ctx, n_ranks, rank, idb))

    def allreduce(self, counters):
        arr = np.ascontiguousarray(counters, np.uint64).copy()
        self._modem._ck(lib().modem_gpu_allreduce_counters(self._c, arr.ctypes.data, arr.size))
        return arr

    def close(self):
        if self._c.value:
            lib().modem_gpu_comm_destroy(self._c)
            self._c = C.c_void_p()
