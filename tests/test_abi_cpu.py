"""CPU-side checks of the product boundary: the C-ABI library loads, exports every symbol
include/modem_gpu.h declares, its host-side tables equal the oracle's, and compute entries
fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import MEMORYLESS, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "modem_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(modem_(?:gpu_)?[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = _declared_symbols()
    assert len(names) >= 42
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.modem_gpu_abi_version() == 4


def test_cfg_struct_matches_header(pkg):
    # field order/size of the ctypes mirror vs sizeof(modem_cfg_t) as the library checks it
    assert C.sizeof(pkg.ModemCfg) == 96


def test_host_freq_and_rates(pkg, orc):
    L = orc.lib()
    for hz, sr in [(1000, 10000), (900, 10000), (2500, 10000), (1, 44100), (21000, 48000)]:
        assert pkg.sample_freq(hz, sr) == L.orc_sample_freq(hz, sr)
    assert pkg.samples_per_symbol(220, 10000) == 45  # rates.rs:16
    assert pkg.samples_per_symbol(1250, 10000) == 8


@pytest.mark.parametrize("name", MEMORYLESS)
def test_host_constellation_equals_oracle(pkg, orc, name):
    table, bps, evenodd = pkg.host_constellation(name)
    o = orc.OraclePath(name, 1250, 10000, 2500)
    ref = o.constellation()
    assert bps == o.bps
    assert evenodd == (name == "oqpsk")
    assert table.shape == ref.shape
    assert np.array_equal(table.view(np.uint32), ref.view(np.uint32)), name


def test_host_constellation_known_answers(pkg):
    # the reference's own vectors: qam.rs:69-84 (QAM::new(4, 0.0, 6.0)) and mpsk.rs:50-63
    L = pkg.lib()
    out = np.zeros((16, 2), np.float32)
    assert L.modem_const_qam(4, 0.0, 6.0, out.ctypes.data_as(C.POINTER(C.c_float))) == 4
    assert tuple(out[0b0000]) == (-3.0, -3.0) and tuple(out[0b0001]) == (-3.0, -1.0)
    assert tuple(out[0b1011]) == (1.0, 3.0) and tuple(out[0b1111]) == (3.0, 3.0)
    out4 = np.zeros((4, 2), np.float32)
    assert L.modem_const_mpsk(2, 0.0, 1.0, out4.ctypes.data_as(C.POINTER(C.c_float))) == 2
    assert out4[0, 0] == 1.0 and out4[0, 1] == 0.0 and out4[1, 1] == 1.0 and out4[2, 0] == -1.0 and out4[3, 1] == -1.0
    assert abs(out4[1, 0]) < 1e-3 and abs(out4[2, 1]) < 1e-3 and abs(out4[3, 0]) < 1e-3


def test_invalid_modulation_name(pkg):
    # src/bin/modulate.rs:94 panics on unknown names; stateful schemes are not on this path
    for bad in ["nope", "bfsk", "msk", "dqpsk"]:
        with pytest.raises(pkg.ModemError):
            pkg.host_constellation(bad)


def test_host_phasor_by_name(pkg, orc):
    """The stateful -m rows of modulate.rs:74-95: same constructor constants as the oracle's phasors."""
    L = orc.lib()
    for name, kind in [("bfsk", pkg.capi.PHASOR_BFSK), ("mfsk", pkg.capi.PHASOR_MFSK), ("16cpfsk", pkg.capi.PHASOR_CPFSK),
                       ("msk", pkg.capi.PHASOR_MSK), ("dqpsk", pkg.capi.PHASOR_DMPSK), ("dbpsk", pkg.capi.PHASOR_DMPSK)]:
        ph, bps, evenodd = pkg.host_phasor(name, 1250, 10000)
        o = orc.Phasor()
        assert L.orc_phasor_by_name(C.byref(o), name.encode(), 1250, 10000) == 1
        assert ph.kind == kind and bps == o.bits_per_symbol == ph.bits_per_symbol
        assert evenodd == (name == "msk")
        assert np.float32(ph.amplitude) == np.float32(o.amplitude)
        if name in ("bfsk", "mfsk", "16cpfsk"):
            assert np.float32(ph.deviation).view(np.uint32) == np.float32(o.deviation).view(np.uint32)
        if name in ("dqpsk", "dbpsk"):
            assert np.float32(ph.phase) == np.float32(o.phase) and np.float32(ph.shift) == np.float32(o.shift)
    assert C.sizeof(pkg.Phasor) == 32
    with pytest.raises(pkg.ModemError):
        pkg.host_phasor("qpsk", 1250, 10000)
    with pytest.raises(pkg.ModemError):
        pkg.host_phasor("msk", 3333, 10000)  # sps 3: msk.rs:13 assert


def test_hilbert_taps_equal_oracle(pkg, orc):
    h = pkg.hilbert_taps()
    assert len(h) == 23 and np.array_equal(h, orc.hilbert_taps())
    assert h[11] == 0.0 and h[10] == np.float32(-0.62794) and h[12] == np.float32(0.62794)  # demodulate.rs:58-60


def test_apsk_ring_verification(pkg):
    # apsk.rs:85-97
    L = pkg.lib()
    from rust_modem_b200.capi import Ring
    out = np.zeros((16, 2), np.float32)
    p = out.ctypes.data_as(C.POINTER(C.c_float))
    gap = (Ring * 2)(Ring(0, 4, 0.5, 0.0), Ring(5, 16, 1.0, 0.0))
    assert L.modem_const_apsk(1.0, 4, gap, 2, p) < 0
    bad_radius = (Ring * 1)(Ring(0, 16, 1.5, 0.0))
    assert L.modem_const_apsk(1.0, 4, bad_radius, 1, p) < 0


def test_taps_equal_oracle(pkg, orc):
    assert np.array_equal(pkg.lowpass_taps(), orc.lowpass_taps())
    for span, sps, beta in [(16, 8, 0.35), (8, 4, 0.25), (6, 10, 0.5)]:
        assert np.array_equal(pkg.rrc_taps(span, sps, beta), orc.rrc_taps(span, sps, beta))


def test_sigma_equals_oracle(pkg, orc):
    rrc = orc.rrc_taps(16, 8, 0.35)
    o = orc.OraclePath("qpsk", 1250, 10000, 2500, tx_taps=rrc, rx_taps=rrc, decision_delay=128, slicer_gain=1.0)
    L = pkg.lib()
    cfg = pkg.ModemCfg()
    table, bps, _ = pkg.host_constellation("qpsk")
    cfg.struct_size = C.sizeof(cfg)
    cfg.bits_per_symbol = bps
    cfg.const_iq = table.ctypes.data_as(C.POINTER(C.c_float))
    cfg.n_rx_taps = len(rrc)
    cfg.rx_taps = rrc.ctypes.data_as(C.POINTER(C.c_float))
    cfg.slicer_gain, cfg.rx_gain = 1.0, 2.0
    for db in (0.0, 4.0, 10.0):
        assert L.modem_sigma_for_ebn0(C.byref(cfg), db) == o.sigma_for_ebn0(db)


def test_no_cpu_fallback(pkg):
    """Without a usable sm_100 device creating a context must fail with an error code."""
    n = C.c_int(0)
    rc = pkg.lib().modem_gpu_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ModemError) as ei:
        pkg.Modem()
    assert ei.value.code in (-6, -2)


def test_create_rejects_bad_cfg(pkg):
    L = pkg.lib()
    ctx = C.c_void_p()
    cfg = pkg.ModemCfg()
    assert L.modem_gpu_create(C.byref(ctx), 0, C.byref(cfg)) == -1  # struct_size mismatch
    cfg.struct_size = C.sizeof(cfg)
    cfg.bits_per_symbol = 9
    assert L.modem_gpu_create(C.byref(ctx), 0, C.byref(cfg)) == -1
    assert b"bits_per_symbol" in L.modem_gpu_last_error(None)


def test_entries_reject_null_arguments_without_a_device(pkg):
    """Every compute entry returns an error code for a null context (no abort, no CPU path behind it)."""
    L = pkg.lib()
    cnt = (C.c_uint64 * 2)(0, 0)
    buf = (C.c_uint8 * 16)()
    assert L.modem_gpu_loopback(None, buf, 1, 16, C.c_float(0.0), 0, 0, None, None, None, cnt) == -1
    assert L.modem_gpu_loopback_packed(None, buf, 1, 16, C.c_float(0.0), 0, 0, None, cnt) == -1
    assert L.modem_gpu_loopback_device(None, buf, 1, 16, C.c_float(0.0), 0, 0, None, None, None, None) == -1
    assert L.modem_gpu_modulate(None, buf, 1, 16, None, None) == -1
    assert tuple(cnt) == (0, 0)


def test_hot_kernels_do_not_spill(pkg):
    """The tuned RX kernel sits at its register cap (128 at 8 CTAs/SM): an innocent change to the argument structs
    once cost 8 bytes of spill and 7 % of its speed.  The ptxas log of the in-tree build is the guard."""
    logdir = os.path.join(ROOT, "rust-modem_b200", "lib")
    hot = [("rx_fast_64.ptxas.log", "_ZN2mg14rx_fast_kernelILi64ELi0ELb0ELb0ELi64ELi8ELi4ELi5ELi64ELb0EEE"),
           ("loop_fused_64.ptxas.log", "_ZN2mg14rx_fast_kernelILi64ELi0ELb0ELb0ELi128ELi4ELi4ELi3ELi64ELb1EEE"),  # the fused loopback
           ("rx_dec.ptxas.log", "_ZN2mg13rx_dec_kernelILi64ELb0ELb1ELi1EEE"),  # the fused loopback at the reference's default rates
           ("rx_dec.ptxas.log", "_ZN2mg13rx_dec_kernelILi64ELb0ELb1ELi0EEE"),
           ("rx_dec.ptxas.log", "_ZN2mg13rx_dec_kernelILi64ELb0ELb1ELi3EEE"),  # the demodulate binary's i16 wire at the reference's rates
           ("tx_fast.ptxas.log", "_ZN2mg19tx_rect_fast_kernelILi2ELb0ELb0EEE"),
           ("tx_fast.ptxas.log", "_ZN2mg21tx_shaped_fast_kernelILi8ELi129ELb1ELi2ELi2EEE")]  # C3 TX, sign-product form
    for fn, sym in hot:
        path = os.path.join(logdir, fn)
        if not os.path.exists(path):
            pytest.skip("library was not built here")
        lines = open(path).read().splitlines()
        idx = [i for i, l in enumerate(lines) if "Compiling entry function" in l and sym in l]
        assert idx, (fn, sym)
        block = " ".join(lines[idx[0]: idx[0] + 4])
        assert "0 bytes spill stores, 0 bytes spill loads" in block, block
