"""Out-of-bounds and coverage guards for every kernel family (compute-sanitizer is not available on the GPU pool this
project is tested on, so the checks are the repository's own).

Every DEVICE buffer a kernel writes is an interior window of a larger allocation whose margins hold a canary pattern;
after the call the margins must be untouched (no write before or behind a buffer) and the window must be fully
written (it is pre-filled with a second pattern no result can contain: NaN payloads for samples, 0xEE for bytes).
Every buffer a kernel READS is likewise a window between canaries of NaN / 0xEE, so a read past either end would
poison the (oracle-checked) results.  Shapes are ragged on purpose: odd frame counts against the frames-per-CTA
grouping, frame lengths that end inside a tile, tiles that start before sample 0."""
import ctypes as C

import numpy as np
import pytest

from conftest import path_kwargs

pytestmark = pytest.mark.gpu

G = 1 << 14  # guard elements on each side


class Guarded:
    """A device tensor window between two canary margins."""

    def __init__(self, torch, shape, dtype, fill):
        self.torch = torch
        n = int(np.prod(shape))
        self.raw = torch.empty(n + 2 * G, dtype=dtype, device="cuda")
        self.canary = float("nan") if dtype.is_floating_point else 0xA5
        self.raw.fill_(self.canary)
        self.win = self.raw[G:G + n].view(*shape)
        self.fill = fill
        if fill is not None:
            self.win.fill_(fill)

    def put(self, arr):
        self.win.copy_(self.torch.from_numpy(np.ascontiguousarray(arr)).view(self.win.shape))
        return self

    def ptr(self):
        return self.win.data_ptr()

    def check(self, what, written=True):
        t = self.torch
        lo, hi = self.raw[:G], self.raw[G + self.win.numel():]
        if self.raw.dtype.is_floating_point:
            assert bool(t.isnan(lo).all()) and bool(t.isnan(hi).all()), f"{what}: write outside the buffer"
            if written:
                assert not bool(t.isnan(self.win).any()), f"{what}: buffer not completely written (or poisoned by an out-of-range read)"
        else:
            assert bool((lo == 0xA5).all()) and bool((hi == 0xA5).all()), f"{what}: write outside the buffer"
            if written and self.fill is not None:
                assert not bool((self.win == self.fill).any()), f"{what}: buffer not completely written"
        return self.win.cpu().numpy()


@pytest.fixture()
def torch_mod():
    import torch

    return torch


CASES = [
    # name, scheme, sps, shaped, F, symbols per frame, sigma dB, extra
    ("fused loopback, ragged tiles", "qpsk", 8, False, 19, 512 * 2 + 300, None, {}),
    ("fused loopback, one short frame", "qpsk", 8, False, 1, 40, None, {}),
    ("two kernels 64 taps + noise", "qpsk", 8, False, 21, 1111 + 1, 5.0, {}),
    ("129-tap RRC both sides", "qpsk", 8, True, 13, 700, None, {}),
    ("129-tap RRC + noise", "qpsk", 8, True, 7, 613, 4.0, {}),
    ("qam16 table, sps 8", "qam16", 8, False, 9, 333, None, {}),
    ("generic kernels, sps 45 (reference default rates)", "qpsk", 45, False, 5, 97, None, {}),
    ("generic kernels, odd frame length", "qpsk", 3, False, 6, 333, None, {}),
    ("oqpsk half-symbol offset", "oqpsk", 8, False, 4, 160, None, {}),
    ("no tensor memory", "qpsk", 8, False, 10, 700, None, {"flags": "NO_TMEM"}),
    ("fused loopback at the reference's default rates (sps 45)", "qpsk", 45, False, 5, 300, None, {}),
    ("fused loopback, sps 5, three tiles, no tensor memory", "qpsk", 5, False, 7, 300, None, {"flags": "NO_TMEM"}),
    ("fused loopback, sps 3, frame shorter than a tile", "qpsk", 3, False, 3, 90, None, {}),
]


@pytest.mark.parametrize("name,scheme,sps,shaped,F,nsym,db,extra", CASES, ids=[c[0] for c in CASES])
def test_device_entries_stay_inside_their_buffers(pkg, orc, torch_mod, name, scheme, sps, shaped, F, nsym, db, extra):
    torch = torch_mod
    kw = path_kwargs(scheme, sps=sps, shaped=shaped)
    o = orc.OraclePath(**kw)
    flags = pkg.FLAG_NO_TMEM if extra.get("flags") == "NO_TMEM" else 0
    m = pkg.Modem(flags=flags, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    bits = np.random.default_rng(len(name)).integers(0, 2, (F, nsym * o.bps), dtype=np.uint8)
    nbits = bits.shape[1]
    L = m.frame_samples(nbits)
    K = m.decided_symbols(L)
    sigma = 0.0 if db is None else o.sigma_for_ebn0(db)
    tx_ref = o.modulate(bits)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, sigma=sigma, seed=31, frame0=5, threads=4)

    g_bits = Guarded(torch, (F, nbits), torch.uint8, None).put(bits)
    g_bits.raw[:G] = 0xEE  # a read before / behind the bits would change symbols
    g_bits.raw[G + F * nbits:] = 0xEE
    g_tx = Guarded(torch, (F, L, 2), torch.float32, float("nan"))
    g_sym = Guarded(torch, (F, K), torch.uint8, 0xEE)
    g_out = Guarded(torch, (F, K * o.bps), torch.uint8, 0xEE)
    g_cnt = Guarded(torch, (2,), torch.int64, 0)

    # modulate, then demodulate + count, each into guarded windows
    m.modulate_into(g_bits.win, F, nbits, tx=g_tx.win)
    torch.cuda.synchronize()
    tx = g_tx.check(name + ": modulate tx")
    assert np.array_equal(tx.view(np.uint32), tx_ref.view(np.uint32))
    m.demodulate_count_into(g_tx.win, F, L, g_bits.win, nbits, g_cnt.win, sym=g_sym.win, bits=g_out.win, sigma=sigma, seed=31, frame0=5)
    torch.cuda.synchronize()
    assert np.array_equal(g_sym.check(name + ": sym"), sym_ref) and np.array_equal(g_out.check(name + ": bits"), bits_ref)
    assert tuple(g_cnt.check(name + ": counters", written=False).tolist()) == cnt_ref
    g_bits.raw[:G].fill_(0xA5); g_bits.raw[G + F * nbits:].fill_(0xA5)
    g_bits.check(name + ": bits are read-only", written=False)

    # the loopback entry (fused kernel where the shape allows) into fresh windows
    g_tx2 = Guarded(torch, (F, L, 2), torch.float32, float("nan"))
    g_sym2 = Guarded(torch, (F, K), torch.uint8, 0xEE)
    g_out2 = Guarded(torch, (F, K * o.bps), torch.uint8, 0xEE)
    g_cnt2 = Guarded(torch, (2,), torch.int64, 0)
    m.loopback_device_into(g_bits.win, F, nbits, g_cnt2.win, tx=g_tx2.win, sym=g_sym2.win, bits_out=g_out2.win, sigma=sigma, seed=31, frame0=5)
    torch.cuda.synchronize()
    assert np.array_equal(g_tx2.check(name + ": loopback tx").view(np.uint32), tx_ref.view(np.uint32))
    assert np.array_equal(g_sym2.check(name + ": loopback sym"), sym_ref) and np.array_equal(g_out2.check(name + ": loopback bits"), bits_ref)
    assert tuple(g_cnt2.check(name + ": loopback counters", written=False).tolist()) == cnt_ref

    # full-rate (I, Q) stream and soft values
    g_filt = Guarded(torch, (F, L, 2), torch.float32, float("nan"))
    g_soft = Guarded(torch, (F, K, 2), torch.float32, float("nan"))
    m.demodulate_into(g_tx.win, F, L, sym=g_sym.win, bits=g_out.win, soft=g_soft.win, filt=g_filt.win)
    torch.cuda.synchronize()
    filt_ref, sym0, _ = o.demodulate(tx_ref)
    assert np.array_equal(g_filt.check(name + ": filt").view(np.uint32), filt_ref.view(np.uint32))
    g_soft.check(name + ": soft")
    assert np.array_equal(g_sym.check(name + ": sym (2)"), sym0)
    m.close()


def test_awgn_random_bits_and_frontend_stay_inside_their_buffers(pkg, orc, torch_mod):
    torch = torch_mod
    L = pkg.lib()
    kw = path_kwargs("qpsk", sps=8)
    o = orc.OraclePath(**kw)
    m = pkg.Modem(**kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    # in-place AWGN on an odd-length buffer (a ragged last quad)
    F, Ls = 5, 1003
    g = Guarded(torch, (F, Ls, 2), torch.float32, 0.0)
    m.awgn_inplace(g.win, F, Ls, 0.5, seed=3, frame0=9)
    torch.cuda.synchronize()
    assert np.array_equal(g.check("awgn").view(np.uint32), o.awgn(np.zeros((F, Ls, 2), np.float32), 0.5, seed=3, frame0=9).view(np.uint32))
    # Philox bits: vector path (rows of 16 bytes) and scalar path
    for nbits in (4096, 1001):
        gb = Guarded(torch, (3, nbits), torch.uint8, 0xEE)
        m.random_bits_into(gb.win, 3, nbits, seed=77, frame0=2)
        torch.cuda.synchronize()
        assert np.array_equal(gb.check(f"random bits {nbits}"), orc.random_bits(3, nbits, 77, 2))
    # sync tone and the real-valued output of the modulate binary
    n_pre, nb = 29, 2 * 77
    bits = np.random.default_rng(8).integers(0, 2, (4, nb), dtype=np.uint8)
    Ld = m.frame_samples(nb)
    g_bits = Guarded(torch, (4, nb), torch.uint8, None).put(bits)
    g_re = Guarded(torch, (4, n_pre + Ld), torch.float32, float("nan"))
    assert L.modem_gpu_modulate_real(m._ctx, g_bits.ptr(), 4, nb, n_pre, C.c_float(1.0), g_re.ptr()) == 0
    torch.cuda.synchronize()
    assert np.array_equal(g_re.check("modulate_real").view(np.uint32), o.modulate_real(bits, n_pre, 1.0).view(np.uint32))
    g_tone = Guarded(torch, (3, 50, 2), torch.float32, float("nan"))
    assert L.modem_gpu_preamble(m._ctx, 3, 50, C.c_float(0.7), g_tone.ptr()) == 0
    torch.cuda.synchronize()
    assert np.array_equal(g_tone.check("preamble").view(np.uint32), o.preamble(3, 50, 0.7).view(np.uint32))
    # carrier recovery from the i16 wire: lock + demodulate, per-frame offsets
    wire = np.clip(np.round(o.modulate_real(bits, 0, 1.0) * 12000), -32768, 32767).astype(np.int16)
    Fw, Lw = wire.shape
    Lr = Lw - 64
    Kw = m.decided_symbols(Lr)
    g_wire = Guarded(torch, (Fw, Lw), torch.int16, None).put(wire)
    g_po = Guarded(torch, (Fw,), torch.float32, float("nan"))
    g_sym = Guarded(torch, (Fw, Kw), torch.uint8, 0xEE)
    g_filt = Guarded(torch, (Fw, Lr, 2), torch.float32, float("nan"))
    assert L.modem_gpu_demodulate_real(m._ctx, g_wire.ptr(), pkg.capi.SAMPLES_I16, Fw, Lw, 64, None, 0, g_po.ptr(), g_sym.ptr(), None, None, g_filt.ptr()) == 0
    torch.cuda.synchronize()
    po_ref, filt_ref, sym_ref, _ = o.demodulate_real(wire, lock=64)
    assert np.array_equal(g_po.check("lock phase").view(np.uint32), po_ref.view(np.uint32))
    assert np.array_equal(g_filt.check("demodulate_real filt").view(np.uint32), filt_ref.view(np.uint32))
    assert np.array_equal(g_sym.check("demodulate_real sym"), sym_ref)
    m.close()


@pytest.mark.parametrize("scheme", ["bfsk", "mfsk", "dqpsk", "msk", "16cpfsk"])
def test_stateful_tx_stays_inside_its_buffers(pkg, orc, torch_mod, scheme):
    torch = torch_mod
    kw = dict(scheme=scheme, baud_rate=1250, sample_rate=10000, carrier_hz=1000)
    o = orc.OraclePath(**kw)
    m = pkg.Modem(rx_taps=pkg.lowpass_taps(), **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    F, nsym = 37, 203  # more frames than one scan CTA (32), a ragged scan chunk
    bits = np.random.default_rng(3).integers(0, 2, (F, nsym * o.bps), dtype=np.uint8)
    Ld = m.frame_samples(bits.shape[1])
    g_bits = Guarded(torch, bits.shape, torch.uint8, None).put(bits)
    g_tx = Guarded(torch, (F, Ld, 2), torch.float32, float("nan"))
    m.modulate_into(g_bits.win, F, bits.shape[1], tx=g_tx.win)
    torch.cuda.synchronize()
    assert np.array_equal(g_tx.check(scheme + " tx").view(np.uint32), o.modulate(bits).view(np.uint32))
    m.close()
