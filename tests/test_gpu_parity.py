"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.

Bar (BASELINE.json north_star): demodulated bits and symbol indices bit-exact; TX and
filtered sample buffers within max error <= 1e-5 (relative to the buffer's peak -- the
per-sample relative error is unbounded at carrier zero crossings).  The design goes further
(device libm == glibc, unfused binary32 ops in reference order), so in the default mode the
buffers are also expected to be BIT-identical; `assert_buffers` checks both and reports them
separately.
"""
import os

import numpy as np
import pytest

from conftest import MEMORYLESS, path_kwargs

pytestmark = pytest.mark.gpu

TOL = 1e-5


def assert_buffers(got, ref, what, exact=True):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    peak = float(np.abs(ref).max()) if ref.size else 1.0
    err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()) / (peak or 1.0) if ref.size else 0.0
    assert err <= TOL, f"{what}: max error {err:.3g} of peak exceeds {TOL}"
    if exact:
        nbad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
        assert nbad == 0, f"{what}: within tolerance (max err {err:.3g}) but {nbad}/{ref.size} words not bit-identical"


def make(pkg, orc, **kw):
    return pkg.Modem(**kw), orc.OraclePath(**kw)


def rand_bits(seed, F, nbits):
    return np.random.default_rng(seed).integers(0, 2, (F, nbits), dtype=np.uint8)


# ----------------------------------------------------------------------------- TX
@pytest.mark.parametrize("scheme", MEMORYLESS)
def test_tx_rect_all_schemes(pkg, orc, scheme):
    """Rectangular-hold TX (exact reference semantics) for every memoryless -m scheme."""
    kw = path_kwargs(scheme, sps=8)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(11, 3, 64 * o.bps + (o.bps - 1))  # ragged tail: the partial symbol is dropped
    tx, iq = m.modulate(bits, want_iq=True)
    tx_ref, iq_ref = o.modulate(bits, want_iq=True)
    assert_buffers(iq, iq_ref, f"{scheme} baseband iq")
    assert_buffers(tx, tx_ref, f"{scheme} tx")


@pytest.mark.parametrize("sps,nsym", [(45, 101), (45, 100), (8, 513), (4, 77), (3, 333)])
def test_tx_rect_shapes(pkg, orc, sps, nsym):
    """Reference default rates (sps 45), odd frame lengths (scalar-store path), tile tails."""
    kw = path_kwargs("qpsk", sps=sps)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(5, 4, nsym * 2)
    assert_buffers(m.modulate(bits), o.modulate(bits), f"tx sps={sps} nsym={nsym}")


def test_tx_sample0_and_many_frames(pkg, orc):
    """Carrier counter carried over from a preamble (modulator.rs:9,66) and > frames_per_block frames."""
    kw = path_kwargs("qpsk", sps=8, sample0=9 * 10 - 1)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(6, 70, 128)
    assert_buffers(m.modulate(bits), o.modulate(bits), "tx sample0")


def test_tx_large_sample_index(pkg, orc):
    """n > 2^24: `s as f32` rounds (carrier.rs:18); the GPU must round the same way."""
    kw = path_kwargs("qpsk", sps=8, sample0=(1 << 24) + 12345)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(8, 2, 256)
    assert_buffers(m.modulate(bits), o.modulate(bits), "tx large n")


@pytest.mark.parametrize("force_generic", [False, True])
def test_tx_shaped_rrc(pkg, orc, force_generic, monkeypatch):
    """129-tap RRC pulse shaping: fast (sps 8) and generic polyphase kernels."""
    if force_generic:
        monkeypatch.setenv("MODEM_GPU_FORCE_GENERIC", "1")
    kw = path_kwargs("qpsk", sps=8, shaped=True)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(21, 5, 2 * 700)
    if force_generic:
        tx, iq = m.modulate(bits, want_iq=True)
        tx_ref, iq_ref = o.modulate(bits, want_iq=True)
        assert_buffers(iq, iq_ref, "rrc baseband")
    else:
        tx, tx_ref = m.modulate(bits), o.modulate(bits)
    assert_buffers(tx, tx_ref, "rrc tx")


@pytest.mark.parametrize("sign_form", [True, False])
@pytest.mark.parametrize("scheme", ["qpsk", "bpsk", "bask", "qam16", "16psk"])
def test_tx_shaped_fast_schemes(pkg, orc, scheme, sign_form, monkeypatch):
    """Tuned 129-tap / sps 8 shaped TX for constellations with one magnitude per rail (the sign-product
    form: one exact fma(+-1, round(h*v), acc) per tap instead of multiply + add) and without (plain form);
    both must be bit-identical to the oracle's multiply-round-add-round fold (fir.rs:21-24)."""
    if scheme not in MEMORYLESS:
        pytest.skip(f"{scheme} not a -m scheme name here")
    if not sign_form:
        monkeypatch.setenv("MODEM_GPU_NO_SIGN_FORM", "1")
    kw = path_kwargs(scheme, sps=8, shaped=True)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(23, 5, o.bps * 613)  # ragged against the 256-symbol tile, frames odd against the frame pairs
    assert_buffers(m.modulate(bits), o.modulate(bits), f"{scheme} shaped tx (sign form {sign_form})")


@pytest.mark.parametrize("scheme,sps", [("qam16", 4), ("oqpsk", 10), ("dcqpsk", 5)])
def test_tx_shaped_generic_schemes(pkg, orc, scheme, sps):
    kw = path_kwargs(scheme, sps=sps, shaped=True)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(22, 3, o.bps * 150)
    tx, iq = m.modulate(bits, want_iq=True)
    tx_ref, iq_ref = o.modulate(bits, want_iq=True)
    assert_buffers(iq, iq_ref, f"{scheme} shaped baseband")
    assert_buffers(tx, tx_ref, f"{scheme} shaped tx")


# ----------------------------------------------------------------------------- RX
def _rx_check(m, o, rx, what, exact=True):
    filt_ref, sym_ref, bits_ref = o.demodulate(rx)
    out = m.demodulate(rx, want_filt=True, want_soft=True)
    assert_buffers(out["filt"], filt_ref, what + " full-rate filtered (I,Q)", exact)
    K = sym_ref.shape[1]
    if K:
        d, q = o.p.decision_delay, (o.sps // 2 if o.p.scheme == b"oqpsk" else 0)
        soft_ref = np.stack([filt_ref[:, d::o.sps, 0][:, :K], filt_ref[:, d + q::o.sps, 1][:, :K]], axis=-1)
        assert_buffers(out["soft"], soft_ref, what + " decision-instant (I,Q)", exact)
    if exact:
        assert np.array_equal(out["sym"], sym_ref), what + " symbol indices"
        assert np.array_equal(out["bits"], bits_ref), what + " bits"
    return out, sym_ref, bits_ref


@pytest.mark.parametrize("scheme", MEMORYLESS)
def test_loopback_all_schemes(pkg, orc, scheme):
    """modulate -> demodulate for every memoryless scheme: buffers, symbols, bits."""
    kw = path_kwargs(scheme, sps=8)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(31, 3, o.bps * 300)
    tx = o.modulate(bits)
    out, sym_ref, bits_ref = _rx_check(m, o, tx, scheme)
    K = sym_ref.shape[1]
    if scheme not in ("qam256",):  # 64-tap low-pass ISI closes the 256-QAM eye; parity still holds
        assert np.array_equal(bits_ref, bits[:, : K * o.bps]), f"{scheme}: oracle round trip has bit errors"


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("shaped", [False, True])
def test_rx_fast_and_generic(pkg, orc, shaped, force_generic, monkeypatch):
    """The fast sps-8 kernels (64-tap low-pass, 129-tap RRC) and the generic kernel agree with the oracle."""
    if force_generic:
        monkeypatch.setenv("MODEM_GPU_FORCE_GENERIC", "1")
    kw = path_kwargs("qpsk", sps=8, shaped=shaped)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(41, 37, 2 * 1100)  # > frames_per_block frames, > 2 tiles of 512 symbols
    tx = o.modulate(bits)
    out, sym_ref, bits_ref = _rx_check(m, o, tx, f"shaped={shaped} generic={force_generic}")
    K = sym_ref.shape[1]
    assert np.array_equal(bits_ref, bits[:, : 2 * K])


@pytest.mark.parametrize("sps,nsym", [(45, 120), (4, 301), (3, 100), (8, 9), (8, 4)])
def test_rx_shapes(pkg, orc, sps, nsym):
    """Reference default rates, odd lengths, frames shorter than the filter (K small or 0)."""
    kw = path_kwargs("qpsk", sps=sps)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(43, 3, 2 * nsym)
    _rx_check(m, o, o.modulate(bits), f"rx sps={sps} nsym={nsym}")


def test_rx_phase_offset_and_sample0(pkg, orc):
    """PLL::phase_offset != 0 (demodulator.rs:50) and a carried-over sample counter."""
    kw = path_kwargs("qpsk", sps=8, phase_offset=0.3217, sample0=977)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(45, 4, 2 * 600)
    _rx_check(m, o, o.modulate(bits), "phase_offset")


def test_parity_across_2p24(pkg, orc):
    """Reference quirk: the NCO evaluates `s as f32` (carrier.rs:18), so past 2^24 samples the counter is
    quantised, the double-frequency term stops being a clean tone and the reference's own round trip makes
    bit errors.  The GPU path must reproduce that exactly, not "fix" it."""
    kw = path_kwargs("qpsk", sps=8, sample0=(1 << 24) - 3000)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(46, 3, 2 * 4096)
    tx_ref = o.modulate(bits)
    assert_buffers(m.modulate(bits), tx_ref, "tx across 2^24")
    out, sym_ref, bits_ref = _rx_check(m, o, tx_ref, "rx across 2^24")
    lb = m.loopback(bits)
    K = sym_ref.shape[1]
    assert lb["errors"] == int((bits_ref != bits[:, : 2 * K]).sum())


def test_rx_arbitrary_input(pkg, orc):
    """The demodulator on an arbitrary complex stream (not produced by our TX): only .re is read."""
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    rng = np.random.default_rng(47)
    rx = rng.standard_normal((3, 4096, 2)).astype(np.float32) * 3.0
    _rx_check(m, o, rx, "arbitrary rx")
    rx2 = rx.copy()
    rx2[..., 1] = 0.0
    a, b = m.demodulate(rx), m.demodulate(rx2)
    assert np.array_equal(a["sym"], b["sym"])  # demodulator.rs:46 uses x.re only


def test_fused_mac_flag_within_tolerance(pkg, orc):
    """MODEM_FLAG_FUSED_MAC: FMA accumulation; not bit-identical, still inside the 1e-5 bar, same decisions."""
    kw = path_kwargs("qpsk", sps=8, shaped=True)
    o = orc.OraclePath(**kw)
    m = pkg.Modem(flags=pkg.FLAG_FUSED_MAC, **kw)
    bits = rand_bits(49, 6, 2 * 900)
    tx_ref = o.modulate(bits)
    assert_buffers(m.modulate(bits), tx_ref, "fma tx", exact=False)
    out, sym_ref, bits_ref = _rx_check(m, o, tx_ref, "fma rx", exact=False)
    assert np.array_equal(out["sym"], sym_ref) and np.array_equal(out["bits"], bits_ref)


# ----------------------------------------------------------------------------- AWGN
def test_awgn_matches_oracle(pkg, orc):
    """Philox4x32-10 + Box-Muller stream: same counters, same noise (odd L, frame0 offset)."""
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    buf = np.zeros((3, 1001, 2), np.float32)
    got = m.awgn(buf, 0.75, seed=0xA5A5, frame0=(1 << 33) + 5)
    ref = o.awgn(buf, 0.75, seed=0xA5A5, frame0=(1 << 33) + 5)
    assert_buffers(got, ref, "awgn")
    assert abs(float(got.std()) - 0.75) < 0.02


def test_random_bits_match_oracle(pkg, orc):
    """Philox payload bits generated on the device (BASELINE config 4: "bits from Philox too"): vector and scalar store
    paths, a frame offset past 2^32, host and device destinations."""
    import torch

    m = pkg.Modem(**path_kwargs("qpsk", sps=8))
    for F, nbits in ((5, 16384), (3, 300), (2, 128), (4, 1000)):
        seed, f0 = 0xA5A5 + (3 << 32), (1 << 33) + 9
        ref = orc.random_bits(F, nbits, seed, f0)
        assert np.array_equal(m.random_bits(F, nbits, seed, f0), ref), (F, nbits)
        d = torch.full((F, nbits), 7, dtype=torch.uint8, device="cuda")
        m.random_bits_into(d, F, nbits, seed, f0)
        m.synchronize()
        assert np.array_equal(d.cpu().numpy(), ref), (F, nbits)


@pytest.mark.parametrize("shaped", [False, True])
def test_noisy_loopback_matches_oracle(pkg, orc, shaped):
    """Loopback with the AWGN stage fused into the RX load: decisions and error counts equal the oracle's."""
    kw = path_kwargs("qpsk", sps=8, shaped=shaped)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(51, 24, 2 * 1024)
    sigma = o.sigma_for_ebn0(3.0)
    assert m.sigma_for_ebn0(3.0) == sigma
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, sigma=sigma, seed=77, frame0=1000, threads=4)
    out = m.loopback(bits, sigma=sigma, seed=77, frame0=1000, want_tx=True)
    assert cnt_ref[0] > 0  # the test is only meaningful if noise flips some bits
    assert np.array_equal(out["sym"], sym_ref)
    assert np.array_equal(out["bits"], bits_ref)
    assert (out["errors"], out["compared"]) == cnt_ref
    # separate awgn + demodulate gives the same decisions as the fused load
    noisy = m.awgn(out["tx"], sigma, seed=77, frame0=1000)
    assert np.array_equal(m.demodulate(noisy)["sym"], sym_ref)


# ----------------------------------------------------------------------------- banks, pointers, edges
def test_multichannel_bank(pkg, orc):
    """Independent carriers: frame f uses channel f // frames_per_channel (config 5 in miniature)."""
    n_ch, fpc = 6, 3
    hz = [500 + 4 * c * 37 for c in range(n_ch)]
    kw = path_kwargs("qpsk", sps=8)
    m = pkg.Modem(**kw)
    m.set_channels([pkg.sample_freq(h, 10000) for h in hz], fpc)
    bits = rand_bits(61, n_ch * fpc, 2 * 520)
    tx = m.modulate(bits)
    out = m.demodulate(tx, want_soft=True)
    for c in range(n_ch):
        kc = dict(kw, carrier_hz=hz[c])
        o = orc.OraclePath(**kc)
        sl = slice(c * fpc, (c + 1) * fpc)
        tx_ref = o.modulate(bits[sl])
        assert_buffers(tx[sl], tx_ref, f"channel {c} tx")
        _, sym_ref, bits_ref = o.demodulate(tx_ref, want_filt=False)
        assert np.array_equal(out["sym"][sl], sym_ref) and np.array_equal(out["bits"][sl], bits_ref)


def test_device_pointers_torch(pkg, orc):
    """Device-resident buffers (torch tensors) on torch's current stream."""
    import torch

    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    bits = rand_bits(71, 16, 2 * 2048)
    F, nbits = bits.shape
    L = m.frame_samples(nbits)
    K = m.decided_symbols(L)
    d_bits = torch.from_numpy(bits).cuda()
    d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((F, K * 2), dtype=torch.uint8, device="cuda")
    err, cmp_ = m.loopback_into(d_bits, F, nbits, tx=d_tx, sym=d_sym, bits_out=d_out)
    torch.cuda.synchronize()
    tx_ref = o.modulate(bits)
    _, sym_ref, bits_ref = o.demodulate(tx_ref, want_filt=False)
    assert_buffers(d_tx.cpu().numpy(), tx_ref, "device tx")
    assert np.array_equal(d_sym.cpu().numpy(), sym_ref)
    assert np.array_equal(d_out.cpu().numpy(), bits_ref)
    assert (err, cmp_) == (0, K * 2 * F)


@pytest.mark.parametrize("chunk", [3, 4, 7])
def test_host_loopback_pipeline(pkg, orc, chunk, monkeypatch):
    """Host-buffer loopback runs as a chunked three-lane pipeline (H2D | kernels | D2H): same results,
    including a multi-carrier bank whose channels straddle chunks and noise indexed by global frame id."""
    monkeypatch.setenv("MODEM_GPU_PIPE_CHUNK", str(chunk))
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(91, 26, 2 * 700)
    sigma = o.sigma_for_ebn0(5.0)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, sigma=sigma, seed=5, frame0=40, threads=4)
    out = m.loopback(bits, sigma=sigma, seed=5, frame0=40)
    assert np.array_equal(out["sym"], sym_ref) and np.array_equal(out["bits"], bits_ref)
    assert (out["errors"], out["compared"]) == cnt_ref
    # bank: 6 channels x 4 frames, chunk sizes that do not divide the channel size
    hz = [1100 + 333 * c for c in range(6)]  # 2*f_c stays in the low-pass stop band: error-free
    mb = pkg.Modem(**kw)
    mb.set_channels([pkg.sample_freq(h, 10000) for h in hz], 4)
    bits2 = rand_bits(92, 24, 2 * 600)
    got = mb.loopback(bits2)
    for c in range(6):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c]))
        s_ref, b_ref, _ = oc.loopback(bits2[4 * c: 4 * c + 4])
        assert np.array_equal(got["sym"][4 * c: 4 * c + 4], s_ref), f"channel {c}"
        assert np.array_equal(got["bits"][4 * c: 4 * c + 4], b_ref)
    assert got["errors"] == 0 and got["compared"] == got["bits"].size


@pytest.mark.parametrize("two_kernels,ramp", [(False, True), (False, False), (True, True), (True, False)])
def test_host_loopback_pipeline_fused_and_ramped(pkg, orc, two_kernels, ramp, monkeypatch):
    """The noise-free headline shape through the host-buffer pipeline in its four forms: TX + RX kernels per chunk
    (the default) or ONE fused kernel per chunk that never materialises the samples (MODEM_GPU_PIPE_FUSED=1), equal
    chunks (default) or chunk sizes that ramp up and down at the ends of the call (16, 32, 64 ... 64, 32, 16 here):
    the same bytes from all of them."""
    monkeypatch.setenv("MODEM_GPU_PIPE_CHUNK", "64")
    if not two_kernels:
        monkeypatch.setenv("MODEM_GPU_PIPE_FUSED", "1")
    if ramp:
        monkeypatch.setenv("MODEM_GPU_PIPE_RAMP", "1")
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    bits = rand_bits(93, 301, 2 * 520)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, threads=8)
    n0 = m.launch_count
    out = m.loopback(bits)
    launches = m.launch_count - n0
    assert np.array_equal(out["sym"], sym_ref) and np.array_equal(out["bits"], bits_ref)
    assert (out["errors"], out["compared"]) == cnt_ref and cnt_ref[0] == 0
    n_chunks = (2 * 2 + 4) if ramp else 5  # 16 32 | 64 64 64 (77 -> merged with the runt) | 32 16   or   64 x 4 + 45
    if not two_kernels:
        assert launches == 1 + (7 if ramp else 5), launches  # NCO table + one fused kernel per chunk
    else:
        assert launches == 1 + 2 * (7 if ramp else 5), launches
    del n_chunks


@pytest.mark.parametrize("chunk,shaped", [(0, False), (5, False), (3, True)])
def test_device_loopback_pipeline(pkg, orc, chunk, shaped, monkeypatch):
    """modem_gpu_loopback_device: TX || RX chunk pipeline on two streams with the NCO table; device
    buffers, stream-ordered, counters accumulated on the device.  Same results as the oracle, with
    noise indexed by global frame id and a bank whose channels straddle chunks."""
    import torch

    if chunk:
        monkeypatch.setenv("MODEM_GPU_LOOP_CHUNK", str(chunk))
    kw = path_kwargs("qpsk", sps=8, shaped=shaped)
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    bits = rand_bits(95, 22, 2 * 800)
    F, nbits = bits.shape
    L = m.frame_samples(nbits)
    K = m.decided_symbols(L)
    sigma = o.sigma_for_ebn0(4.0)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, sigma=sigma, seed=11, frame0=3, threads=4)
    d_bits = torch.from_numpy(bits).cuda()
    d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(2):  # second pass: cached table, accumulating counters
        m.loopback_device_into(d_bits, F, nbits, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out, sigma=sigma, seed=11, frame0=3)
    torch.cuda.synchronize()
    assert_buffers(d_tx.cpu().numpy(), o.modulate(bits), "pipelined tx")
    assert np.array_equal(d_sym.cpu().numpy(), sym_ref) and np.array_equal(d_out.cpu().numpy(), bits_ref)
    assert tuple(d_cnt.tolist()) == (2 * cnt_ref[0], 2 * cnt_ref[1])
    # bank with non-zero per-channel phase offsets (separate RX table)
    hz = [1100 + 333 * c for c in range(4)]
    po = [0.0, 0.11, -0.2, 0.05]
    mb = pkg.Modem(**kw)
    mb.set_stream(torch.cuda.current_stream().cuda_stream)
    mb.set_channels([pkg.sample_freq(h, 10000) for h in hz], 5, phase_offsets=po)
    bits2 = rand_bits(96, 20, 2 * 640)
    d_b2 = torch.from_numpy(bits2).cuda()
    L2 = mb.frame_samples(bits2.shape[1]); K2 = mb.decided_symbols(L2)
    d_s2 = torch.empty((20, K2), dtype=torch.uint8, device="cuda")
    d_c2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    mb.loopback_device_into(d_b2, 20, bits2.shape[1], d_c2, sym=d_s2)
    torch.cuda.synchronize()
    for c in range(4):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c], phase_offset=po[c]))
        tx_c = orc.OraclePath(**dict(kw, carrier_hz=hz[c])).modulate(bits2[5 * c: 5 * c + 5])  # TX has no phase offset
        _, s_ref, _ = oc.demodulate(tx_c, want_filt=False)
        assert np.array_equal(d_s2[5 * c: 5 * c + 5].cpu().numpy(), s_ref), f"channel {c}"


def test_loop_graph_cache_survives_table_rebuilds(pkg, orc, monkeypatch):
    """The chunked device loopback replays a captured CUDA graph when the same call repeats.  The graph bakes in the NCO
    table's address and contents: a call with longer frames in between re-allocates and rebuilds that table, after which
    the old graph must not be replayed (it would read freed memory or the wrong carrier values)."""
    import torch

    monkeypatch.setenv("MODEM_GPU_LOOP_CHUNK", "5")
    kw = path_kwargs("qpsk", sps=8, shaped=True)  # two-kernel path: the chunk pipeline and its graph
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)

    def run(bits, d):
        F, nbits = bits.shape
        d["cnt"].zero_()
        m.loopback_device_into(d["bits"], F, nbits, d["cnt"], tx=d["tx"], sym=d["sym"], bits_out=d["out"])
        torch.cuda.synchronize()
        return d["tx"].cpu().numpy(), d["sym"].cpu().numpy(), tuple(d["cnt"].tolist())

    def bufs(bits):
        F, nbits = bits.shape
        L = m.frame_samples(nbits); K = m.decided_symbols(L)
        return dict(bits=torch.from_numpy(bits).cuda(), tx=torch.empty((F, L, 2), dtype=torch.float32, device="cuda"),
                    sym=torch.empty((F, K), dtype=torch.uint8, device="cuda"), out=torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda"),
                    cnt=torch.zeros(2, dtype=torch.int64, device="cuda"))

    a, b = rand_bits(97, 22, 2 * 600), rand_bits(98, 22, 2 * 2000)
    da, db = bufs(a), bufs(b)
    ref = {}
    for name, bits in (("a", a), ("b", b)):
        sym_ref, _, cnt_ref = o.loopback(bits, threads=4)
        ref[name] = (o.modulate(bits), sym_ref, cnt_ref)
    for name, bits, d in (("a", a, da), ("a", a, da), ("b", b, db), ("a", a, da), ("b", b, db), ("a", a, da)):
        tx, sym, cnt = run(bits, d)
        assert_buffers(tx, ref[name][0], f"tx of call {name}")
        assert np.array_equal(sym, ref[name][1]) and cnt == ref[name][2], name


def test_ber_sweep_matches_oracle(pkg, orc):
    """BASELINE config 4 in miniature: Eb/N0 sweep, modulate once, per-point counters equal the oracle's
    (noise key = seed + point, counter = global frame id), and two 'ranks' with disjoint frame ranges add up."""
    import torch

    kw = path_kwargs("qpsk", sps=8, shaped=True)
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    bits = rand_bits(97, 12, 2 * 1024)
    dbs = [0.0, 3.0, 6.0]
    sig = [o.sigma_for_ebn0(d) for d in dbs]
    ref = [o.loopback(bits, sigma=s, seed=0xA5A5 + p, frame0=100, threads=4, want_out=False)[2] for p, s in enumerate(sig)]
    d_bits = torch.from_numpy(bits).cuda()
    cnt = torch.zeros((3, 2), dtype=torch.int64, device="cuda")
    m.ber_sweep_into(d_bits, 12, bits.shape[1], sig, cnt, seed=0xA5A5, frame0=100)
    torch.cuda.synchronize()
    assert [tuple(r) for r in cnt.tolist()] == ref
    assert ref[0][0] > ref[1][0] > ref[2][0] > 0
    # sharded: frames [0,5) and [5,12) with their own frame0, counters summed (what the all-reduce does)
    c0 = torch.zeros((3, 2), dtype=torch.int64, device="cuda")
    c1 = torch.zeros((3, 2), dtype=torch.int64, device="cuda")
    m.ber_sweep_into(d_bits[:5].contiguous(), 5, bits.shape[1], sig, c0, seed=0xA5A5, frame0=100)
    m.ber_sweep_into(d_bits[5:].contiguous(), 7, bits.shape[1], sig, c1, seed=0xA5A5, frame0=105)
    torch.cuda.synchronize()
    assert torch.equal(c0 + c1, cnt)


def test_empty_and_degenerate_inputs(pkg, orc):
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    # no frames
    assert m.modulate(np.zeros((0, 64), np.uint8)).shape == (0, 256, 2)
    # fewer bits than one symbol: Bits yields Finished immediately (data.rs:58-62)
    assert m.modulate(np.zeros((2, 1), np.uint8)).shape == (2, 0, 2)
    # frame shorter than the decision delay: nothing decided
    bits = rand_bits(81, 2, 2 * 4)
    out = m.loopback(bits)
    assert out["sym"].shape == (2, 0) and out["compared"] == 0
    assert m.decided_symbols(35) == 0 and m.decided_symbols(36) == 1 == o.decided_symbols(36)


# ----------------------------------------------------------------------------- full size
def test_full_size_c2_properties(pkg, orc):
    """BASELINE config 2 (4096 frames x 65536 samples, QPSK, sps 8, rect + 64-tap low-pass) at full
    size: size-independent properties + sampled frames against the oracle."""
    import torch

    F, nsym = 4096, 8192
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(1234)
    d_bits = torch.randint(0, 2, (F, 2 * nsym), dtype=torch.uint8, device="cuda", generator=g)
    L = m.frame_samples(2 * nsym)
    K = m.decided_symbols(L)
    assert (L, K) == (65536, 8188)
    d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((F, K * 2), dtype=torch.uint8, device="cuda")
    err, cmp_ = m.loopback_into(d_bits, F, 2 * nsym, tx=d_tx, sym=d_sym, bits_out=d_out)
    torch.cuda.synchronize()
    # round trip: every decided bit equals the transmitted bit, and the counters say so
    assert (err, cmp_) == (0, F * K * 2)
    assert torch.equal(d_out, d_bits[:, : 2 * K])
    # symbol indices are the MSB-first packing of the bits
    assert torch.equal(d_sym, d_out[:, 0::2] * 2 + d_out[:, 1::2])
    # constant envelope: |tx| == amplitude for QPSK
    mag = torch.linalg.vector_norm(d_tx[::257], dim=-1)
    assert float((mag - 1.0).abs().max()) < 1e-6
    # sampled frames against the oracle (first, last, a few in between)
    idx = [0, 1, 31, 32, 2047, 4095]
    bits_s = d_bits[idx].cpu().numpy()
    tx_ref = o.modulate(bits_s)
    assert_buffers(d_tx[idx].cpu().numpy(), tx_ref, "C2 sampled tx")
    _, sym_ref, _ = o.demodulate(tx_ref, want_filt=False)
    assert np.array_equal(d_sym[idx].cpu().numpy(), sym_ref)


def test_full_size_c3_properties(pkg, orc):
    """BASELINE config 3 at full size: 16384 frames x 65536 samples = 2^30 samples (an 8 GiB sample buffer: every byte
    offset beyond 2^32), 129-tap RRC both sides.  Round trip without a bit error over all 2.1e9 decided bits, symbol
    packing, and sampled frames -- the first, the last, and the ones either side of the 2^32-byte boundaries of the
    sample buffer -- against the oracle, samples bit for bit."""
    import torch

    F, nsym = 16384, 8192
    kw = path_kwargs("qpsk", sps=8, shaped=True)
    m, o = make(pkg, orc, **kw)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(4321)
    d_bits = torch.randint(0, 2, (F, 2 * nsym), dtype=torch.uint8, device="cuda", generator=g)
    L = m.frame_samples(2 * nsym)
    K = m.decided_symbols(L)
    assert (L, K) == (65536, 8176)
    d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((F, K * 2), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    m.loopback_device_into(d_bits, F, 2 * nsym, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
    torch.cuda.synchronize()
    assert tuple(d_cnt.tolist()) == (0, F * K * 2)
    assert torch.equal(d_out, d_bits[:, : 2 * K])
    assert torch.equal(d_sym, d_out[:, 0::2] * 2 + d_out[:, 1::2])
    per4g = (1 << 32) // (L * 8)  # frames per 4 GiB of samples: 8192
    idx = [0, 1, per4g - 1, per4g, per4g + 1, 2 * per4g - 1, F - 2, F - 1][:7] + [F - 1]
    idx = sorted(set(idx))
    bits_s = d_bits[idx].cpu().numpy()
    tx_ref = o.modulate(bits_s)
    assert_buffers(d_tx[idx].cpu().numpy(), tx_ref, "C3 sampled tx")
    _, sym_ref, _ = o.demodulate(tx_ref, want_filt=False)
    assert np.array_equal(d_sym[idx].cpu().numpy(), sym_ref)


def test_full_size_c1_rates_fused_and_two_kernels(pkg, orc):
    """The reference's default rates (sr 10000 / baud 220 -> sps 45, carrier 1000 Hz, 64-tap low-pass) at the bench's size:
    4096 frames x 65520 samples (12 tiles per frame).  ONE fused kernel (rx_dec_kernel<..., TXF>) and TX + RX: round trip
    without a bit error, TX buffers of both forms identical in every bit (NaN pre-fill: completely written), symbols and
    bits identical, sampled frames against the oracle."""
    import torch

    F, nsym = 4096, 1456
    kw = path_kwargs("qpsk", sps=45)
    o = orc.OraclePath(**kw)
    g = torch.Generator(device="cuda").manual_seed(145)
    d_bits = torch.randint(0, 2, (F, 2 * nsym), dtype=torch.uint8, device="cuda", generator=g)
    res = {}
    for form in ("fused", "two kernels"):
        m = pkg.Modem(**kw)
        m.set_stream(torch.cuda.current_stream().cuda_stream)
        L = m.frame_samples(2 * nsym)
        K = m.decided_symbols(L)
        assert (L, K) == (65520, 1455)
        d_tx = torch.full((F, L, 2), float("nan"), dtype=torch.float32, device="cuda")
        d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
        d_out = torch.empty((F, K * 2), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        n0 = m.launch_count
        if form == "fused":
            m.loopback_device_into(d_bits, F, 2 * nsym, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
        else:
            m.modulate_into(d_bits, F, 2 * nsym, tx=d_tx)
            m.demodulate_count_into(d_tx, F, L, d_bits, 2 * nsym, d_cnt, sym=d_sym, bits=d_out)
        torch.cuda.synchronize()
        assert m.launch_count - n0 == (2 if form == "fused" else 3), form  # + the NCO table
        assert tuple(d_cnt.tolist()) == (0, F * K * 2), form
        assert torch.equal(d_out, d_bits[:, : 2 * K])
        assert torch.equal(d_sym, d_out[:, 0::2] * 2 + d_out[:, 1::2])
        assert not bool(torch.isnan(d_tx).any()), form
        res[form] = (d_tx, d_sym)
        m.close()
    assert torch.equal(res["fused"][0].view(torch.int32), res["two kernels"][0].view(torch.int32))
    assert torch.equal(res["fused"][1], res["two kernels"][1])
    idx = [0, 1, 15, 16, 2047, 4095]
    tx_ref = o.modulate(d_bits[idx].cpu().numpy())
    assert_buffers(res["fused"][0][idx].cpu().numpy(), tx_ref, "C1 sampled tx")
    _, sym_ref, _ = o.demodulate(tx_ref, want_filt=False)
    assert np.array_equal(res["fused"][1][idx].cpu().numpy(), sym_ref)


def test_full_size_c5_bank_properties(pkg, orc):
    """BASELINE config 5 as one GPU carries it: 128 carriers x 32 frames x 65536 samples (frames grouped by channel,
    frames_per_block dividing the channel size), through the fused loopback kernel and through TX + RX: round trip without
    a bit error, both forms byte-identical, and sampled frames of sampled channels (first, last, channel boundaries)
    against an oracle configured for that channel's carrier."""
    import torch

    n_ch, fpc, nsym = 128, 32, 8192
    F = n_ch * fpc
    kw = path_kwargs("qpsk", sps=8)
    hz = [1000 + (3000 * (896 + c)) // 1024 for c in range(n_ch)]  # the carriers rank 7 of 8 carries in bench.py
    res = {}
    g = torch.Generator(device="cuda").manual_seed(55)
    d_bits = torch.randint(0, 2, (F, 2 * nsym), dtype=torch.uint8, device="cuda", generator=g)
    for form in ("fused", "two kernels"):
        m = pkg.Modem(**kw)
        m.set_stream(torch.cuda.current_stream().cuda_stream)
        m.set_channels([pkg.sample_freq(h, 10000) for h in hz], fpc)
        L = m.frame_samples(2 * nsym)
        K = m.decided_symbols(L)
        d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
        d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
        d_out = torch.empty((F, K * 2), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        n0 = m.launch_count
        if form == "fused":
            m.loopback_device_into(d_bits, F, 2 * nsym, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
        else:
            m.modulate_into(d_bits, F, 2 * nsym, tx=d_tx)
            m.demodulate_count_into(d_tx, F, L, d_bits, 2 * nsym, d_cnt, sym=d_sym, bits=d_out)
        torch.cuda.synchronize()
        assert m.launch_count - n0 == (2 if form == "fused" else 3)  # + the NCO table (128 rows)
        assert tuple(d_cnt.tolist()) == (0, F * K * 2), form
        assert torch.equal(d_out, d_bits[:, : 2 * K])
        res[form] = (d_tx, d_sym)
        m.close()
    assert torch.equal(res["fused"][0].view(torch.int32), res["two kernels"][0].view(torch.int32))
    assert torch.equal(res["fused"][1], res["two kernels"][1])
    for c in (0, 1, 63, 127):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c]))
        rows = [c * fpc, c * fpc + fpc - 1]
        tx_ref = oc.modulate(d_bits[rows].cpu().numpy())
        assert_buffers(res["fused"][0][rows].cpu().numpy(), tx_ref, f"C5 channel {c} tx")
        _, sym_ref, _ = oc.demodulate(tx_ref, want_filt=False)
        assert np.array_equal(res["fused"][1][rows].cpu().numpy(), sym_ref), c


# ----------------------------------------------------------------------------- fused loopback kernel
def _device_loopback(pkg, m, bits, want_sym=True, want_bits=True, want_tx=True):
    import torch

    m.set_stream(torch.cuda.current_stream().cuda_stream)
    F, nbits = bits.shape
    L = m.frame_samples(nbits)
    K = m.decided_symbols(L)
    d_bits = torch.from_numpy(bits).cuda()
    d_tx = torch.full((F, L, 2), float("nan"), dtype=torch.float32, device="cuda") if want_tx else None  # every sample must be written
    d_sym = torch.full((F, K), 255, dtype=torch.uint8, device="cuda") if want_sym else None
    d_out = torch.full((F, 2 * K), 255, dtype=torch.uint8, device="cuda") if want_bits else None
    d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    n0 = m.launch_count
    m.loopback_device_into(d_bits, F, nbits, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
    torch.cuda.synchronize()
    return (None if d_tx is None else d_tx.cpu().numpy(), None if d_sym is None else d_sym.cpu().numpy(), None if d_out is None else d_out.cpu().numpy(),
            tuple(d_cnt.tolist()), m.launch_count - n0)


FUSED_TS = 512  # symbols per tile of the fused loopback kernel (loop_fused_64.cu: 128 threads x 4 symbols)


@pytest.mark.parametrize("F,nsym", [(1, 8), (16, 2048), (6, 1064), (19, 8 * 37), (3, 300), (21, 1112), (9, 256 * 3 + 252), (21, 1111), (40, 2048 + 5), (9, 256 * 3 + 251), (5, 256 * 2 + 4),
                                    (7, 512 + 508), (5, 512 * 2 + 4), (3, 512 * 3), (4, 512 + 8)])
def test_fused_loopback_matches_oracle_and_two_kernel_path(pkg, orc, F, nsym, monkeypatch):
    """The fused loopback kernel (TX samples made, stored and demodulated by one kernel) against the oracle and
    against the two-kernel path: TX buffer bit-identical and completely written (ragged tiles, frames that end inside
    a tile's halo, frame counts that do not fill a CTA's frame group), symbols, bits and counters equal.  Shapes whose
    tiles do not reach the frame's end fall back to two kernels (last case): same results, two launches."""
    kw = path_kwargs("qpsk", sps=8)
    bits = rand_bits(200 + F, F, 2 * nsym)
    m, o = make(pkg, orc, **kw)
    tx_ref = o.modulate(bits)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, threads=4)
    tx, sym, out, cnt, launches = _device_loopback(pkg, m, bits)
    assert_buffers(tx, tx_ref, "fused tx")
    assert np.array_equal(sym, sym_ref) and np.array_equal(out, bits_ref)
    assert cnt == (cnt_ref[0], cnt_ref[1]) and cnt[0] == 0
    monkeypatch.setenv("MODEM_GPU_NO_FUSED_LOOP", "1")
    m2 = pkg.Modem(**kw)
    tx2, sym2, out2, cnt2, launches2 = _device_loopback(pkg, m2, bits)
    assert np.array_equal(tx.view(np.uint32), tx2.view(np.uint32))
    assert np.array_equal(sym, sym2) and np.array_equal(out, out2) and cnt == cnt2
    # two kernels when K = nsym - 4 fills its tiles exactly (the last tile stops 4 samples short of the frame's end)
    # or when the rows of bits (2 bytes per symbol) do not start on 8-byte boundaries
    fallback = nsym % FUSED_TS == 4 or nsym % 4 != 0
    assert launches2 == 3 and launches == (3 if fallback else 2), (launches, launches2)  # NCO table + kernels


def test_fused_loopback_is_one_kernel_and_optional_outputs(pkg, orc):
    """Headline shape: exactly one kernel besides the NCO table; outputs may be left out."""
    kw = path_kwargs("qpsk", sps=8)
    bits = rand_bits(77, 33, 2 * 1024)
    m, o = make(pkg, orc, **kw)
    _device_loopback(pkg, m, bits)  # first call also builds the NCO table
    tx, sym, out, cnt, launches = _device_loopback(pkg, m, bits)
    assert launches == 1, launches
    tx_ref = o.modulate(bits)
    assert_buffers(tx, tx_ref, "fused tx")
    tx3, sym3, out3, cnt3, _ = _device_loopback(pkg, m, bits, want_sym=False, want_bits=False)
    assert sym3 is None and out3 is None and cnt3 == cnt
    assert np.array_equal(tx3.view(np.uint32), tx_ref.view(np.uint32))
    # no TX buffer asked for: the samples never leave the chip, the decisions are the same
    tx4, sym4, out4, cnt4, launches4 = _device_loopback(pkg, m, bits, want_tx=False)
    assert tx4 is None and launches4 == 1 and np.array_equal(sym4, sym) and np.array_equal(out4, out) and cnt4 == cnt


def test_fused_loopback_rotated_table_and_no_tmem(pkg, orc, monkeypatch):
    """The fused kernel takes ANY one 4-point table (symbol pairs are looked up by index, not derived from a sign
    symmetry): QPSK rotated by 0.3 rad has no axis-aligned slicer either, so this also exercises the nearest-point
    search inside the fused kernel.  No oracle scheme name reaches that table, so the reference here is the two-kernel
    GPU path (tx_rect_fast + rx_fast, each pinned to the oracle by the other tests).  MODEM_FLAG_NO_TMEM (no tensor
    memory allocated) must not change a bit either."""
    import ctypes as C

    tab = np.zeros((4, 2), np.float32)
    assert pkg.lib().modem_const_qpsk(C.c_float(0.3), C.c_float(1.0), tab.ctypes.data_as(C.POINTER(C.c_float))) == 2
    kw = path_kwargs("qpsk", sps=8)
    kw.pop("scheme")
    bits = rand_bits(91, 11, 2 * 1300)
    res = {}
    for name, env, flags in (("fused", None, 0), ("fused, no tmem", None, pkg.FLAG_NO_TMEM), ("two kernels", "1", 0), ("two kernels, no tmem", "1", pkg.FLAG_NO_TMEM)):
        if env:
            monkeypatch.setenv("MODEM_GPU_NO_FUSED_LOOP", env)
        m = pkg.Modem(const_iq=tab[None], bps=2, flags=flags, **kw)
        res[name] = _device_loopback(pkg, m, bits)
        m.close()
    tx, sym, out, cnt, launches = res["fused"]
    assert launches == 2 and res["two kernels"][4] == 3
    assert cnt[0] == 0 and cnt[1] == out.size and np.array_equal(out, bits[:, : out.shape[1]])
    for name, r in res.items():
        assert np.array_equal(r[0].view(np.uint32), tx.view(np.uint32)), name
        assert np.array_equal(r[1], sym) and np.array_equal(r[2], out) and r[3] == cnt, name
    # the unrotated table through the same four forms, against the oracle
    kw = path_kwargs("qpsk", sps=8)
    o = orc.OraclePath(**kw)
    tx_ref = o.modulate(bits)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, threads=4)
    monkeypatch.delenv("MODEM_GPU_NO_FUSED_LOOP")
    for env, flags in ((None, pkg.FLAG_NO_TMEM), ("1", pkg.FLAG_NO_TMEM)):
        if env:
            monkeypatch.setenv("MODEM_GPU_NO_FUSED_LOOP", env)
        m = pkg.Modem(flags=flags, **kw)
        tx, sym, out, cnt, _ = _device_loopback(pkg, m, bits)
        assert_buffers(tx, tx_ref, "no-tmem tx")
        assert np.array_equal(sym, sym_ref) and np.array_equal(out, bits_ref) and cnt == cnt_ref


def test_sign_slicer_equals_the_search(pkg, orc, monkeypatch):
    """The QPSK sign slicer of the fast RX kernel (b0 = I > 0, b1 = Q > 0 inside a window of |I|, |Q|, the nearest-point
    search outside it) must return the search's answer -- roundings and tie rule (lowest index) included -- for ANY
    input: ordinary soft values, exact zeros (all four distances tie -> symbol 0), tiny values whose distance difference
    is lost to rounding, huge values, infinities and NaNs."""
    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    rng = np.random.default_rng(5)
    F, L = 12, 8 * 700
    rx = rng.standard_normal((F, L, 2)).astype(np.float32)
    rx[1] = 0.0                                             # exact zeros: ties
    rx[2] *= np.float32(1e-9)                               # far below the window: search
    rx[3] *= np.float32(1e-3)                               # around the window's lower edge
    rx[4] *= np.float32(1e6)                                # above the window
    rx[5, ::97, 0] = np.nan                                 # NaNs poison a filter span each
    rx[6, ::131, 0] = np.inf
    rx[7, ::89, 0] = -np.inf
    rx[8] = np.float32(1e-41)                               # denormals
    rx[9, :, 0] = np.where(np.arange(L) % 16 < 8, 1.0, -1.0).astype(np.float32) * np.float32(2.0 ** -11)  # I, Q near lo
    rx[10] *= np.float32(1e30)                              # overflows to inf inside the FIR
    with np.errstate(all="ignore"):
        _, sym_ref, bits_ref = o.demodulate(rx, want_filt=False)
    got = m.demodulate(rx)
    assert np.array_equal(got["sym"], sym_ref) and np.array_equal(got["bits"], bits_ref)
    assert len(np.unique(sym_ref[0])) == 4
    monkeypatch.setenv("MODEM_GPU_NO_SIGN_SLICE", "1")
    m2 = pkg.Modem(**kw)
    got2 = m2.demodulate(rx)
    assert np.array_equal(got2["sym"], sym_ref) and np.array_equal(got2["bits"], bits_ref)


def test_fused_loopback_bank_and_fallbacks(pkg, orc):
    """A carrier bank (one NCO table row per channel) runs fused; a phase offset, another constellation or an even
    decision delay do not, and give the same answers through the two-kernel path."""
    kw = path_kwargs("qpsk", sps=8)
    hz = [700 + 411 * c for c in range(5)]
    mb = pkg.Modem(**kw)
    mb.set_channels([pkg.sample_freq(h, 10000) for h in hz], 4)
    bits = rand_bits(78, 20, 2 * 700)
    tx, sym, out, cnt, _ = _device_loopback(pkg, mb, bits)
    for c in range(5):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c]))
        rows = slice(4 * c, 4 * c + 4)
        assert_buffers(tx[rows], oc.modulate(bits[rows]), f"bank tx, channel {c}")
        s_ref, b_ref, _ = oc.loopback(bits[rows])
        assert np.array_equal(sym[rows], s_ref) and np.array_equal(out[rows], b_ref)
    assert cnt[0] == 0 and cnt[1] == out.size
    for over in (dict(phase_offset=0.3), dict(decision_delay=34), dict(scheme="bpsk")):
        kw2 = dict(path_kwargs(over.get("scheme", "qpsk"), sps=8), **{k: v for k, v in over.items() if k != "scheme"})
        m, o = make(pkg, orc, **kw2)
        b = rand_bits(79, 6, o.bps * 900)
        tx, sym, out, cnt, launches = _device_loopback(pkg, m, b) if o.bps == 2 else (None,) * 5
        if tx is None:
            got = m.loopback(b, want_tx=True)
            tx, sym, out = got["tx"], got["sym"], got["bits"]
        assert_buffers(tx, o.modulate(b), f"fallback tx {over}")
        s_ref, b_ref, _ = o.loopback(b)
        assert np.array_equal(sym, s_ref) and np.array_equal(out, b_ref), over


def _dec_fused_shape(L, K, sps, delay):
    """loop_fused_dec_supported (rx_dec.cu) restated: the tiles' staged ranges cover [0, L) without a gap."""
    NT, TS = 64, 128
    if sps == 8 or sps > NT or delay > NT - 1 or K == 0 or L % 2:
        return False
    tiles = (K + TS - 1) // TS
    nb = (tiles - 1) * TS * sps + delay - (NT - 1)
    nb0 = nb & ~1
    R = (TS - 1) * sps + NT + (nb - nb0)
    return nb0 + 2 * ((R + 1) // 2) >= L


@pytest.mark.parametrize("sps,F,nsym", [(45, 5, 300), (45, 17, 1456), (45, 3, 128), (45, 2, 129), (45, 1, 4), (45, 4, 257), (10, 6, 777), (5, 9, 1500),
                                        (4, 4, 1001), (3, 7, 2000), (3, 5, 2 * 128 + 22), (45, 3, 255), (10, 33, 128 * 3)])
def test_fused_loopback_any_sps(pkg, orc, sps, F, nsym, monkeypatch):
    """The fused loopback at the reference's default rates and other samples-per-symbol counts (rx_dec_kernel<..., TXF>):
    TX buffer bit-identical to the oracle and completely written (NaN pre-fill), symbols, bits and counters equal, one
    kernel when the tiles cover the frame; shapes that do not qualify (odd frame lengths, tiles that stop short of the
    frame's end) run the two kernels with the same results.  MODEM_FLAG_NO_TMEM and a missing TX buffer change nothing."""
    kw = path_kwargs("qpsk", sps=sps)
    bits = rand_bits(400 + sps + F, F, 2 * nsym)
    m, o = make(pkg, orc, **kw)
    tx_ref = o.modulate(bits)
    sym_ref, bits_ref, cnt_ref = o.loopback(bits, threads=4)
    tx, sym, out, cnt, launches = _device_loopback(pkg, m, bits)
    assert_buffers(tx, tx_ref, f"fused tx sps {sps}")
    assert np.array_equal(sym, sym_ref) and np.array_equal(out, bits_ref)
    assert cnt == (cnt_ref[0], cnt_ref[1])
    L = m.frame_samples(2 * nsym)
    fused = _dec_fused_shape(L, m.decided_symbols(L), sps, kw["decision_delay"])
    assert launches == (2 if fused else 3), (launches, fused)  # NCO table + kernel(s)
    if fused:
        m3 = pkg.Modem(flags=pkg.FLAG_NO_TMEM, **kw)
        tx3, sym3, out3, cnt3, l3 = _device_loopback(pkg, m3, bits)
        assert l3 == 2 and np.array_equal(tx3.view(np.uint32), tx_ref.view(np.uint32)) and np.array_equal(out3, bits_ref) and cnt3 == cnt
        tx4, sym4, out4, cnt4, l4 = _device_loopback(pkg, m, bits, want_tx=False)
        assert tx4 is None and l4 == 1 and np.array_equal(sym4, sym_ref) and cnt4 == cnt
    monkeypatch.setenv("MODEM_GPU_NO_FUSED_LOOP", "1")
    m2 = pkg.Modem(**kw)
    tx2, sym2, out2, cnt2, launches2 = _device_loopback(pkg, m2, bits)
    assert launches2 == 3 and np.array_equal(tx.view(np.uint32), tx2.view(np.uint32))
    assert np.array_equal(sym, sym2) and np.array_equal(out, out2) and cnt == cnt2


def test_fused_loopback_any_sps_bank(pkg, orc):
    """A carrier bank at the reference's default rates through the fused kernel: one NCO table row per channel."""
    kw = path_kwargs("qpsk", sps=45)
    hz = [700 + 411 * c for c in range(5)]
    mb = pkg.Modem(**kw)
    mb.set_channels([pkg.sample_freq(h, 10000) for h in hz], 4)
    bits = rand_bits(81, 20, 2 * 300)
    tx, sym, out, cnt, launches = _device_loopback(pkg, mb, bits)
    assert launches == 2
    for c in range(5):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c]))
        rows = slice(4 * c, 4 * c + 4)
        assert_buffers(tx[rows], oc.modulate(bits[rows]), f"bank tx, channel {c}")
        s_ref, b_ref, _ = oc.loopback(bits[rows])
        assert np.array_equal(sym[rows], s_ref) and np.array_equal(out[rows], b_ref)
    assert cnt[1] == out.size


# ----------------------------------------------------------------------------- packed payloads (extension)
@pytest.mark.parametrize("case", ["fused", "noisy", "shaped", "ragged", "bpsk_odd"])
@pytest.mark.parametrize("chunk", [0, 3])
def test_loopback_packed(pkg, orc, case, chunk, monkeypatch):
    """modem_gpu_loopback_packed: packed rows in, packed rows out; decisions and counters are those of the oracle's
    loopback on the unpacked bits (fused kernel, noisy two-kernel path with global frame ids across chunks, the 129-tap
    pair, row lengths that are not multiples of 8 on either side)."""
    if chunk:
        monkeypatch.setenv("MODEM_GPU_PACKED_CHUNK", str(chunk))
    scheme, nbits, F, sigma_db, kw = "qpsk", 2 * 704, 10, None, {}
    if case == "noisy":
        sigma_db = 4.0
    elif case == "shaped":
        kw = dict(shaped=True)
    elif case == "ragged":
        nbits = 2 * 701 + 1  # 1403 bits: the last packed byte carries 3 bits, the trailing half symbol is dropped
    elif case == "bpsk_odd":
        scheme, nbits = "bpsk", 517
    pk = path_kwargs(scheme, sps=8, **kw)
    m, o = make(pkg, orc, **pk)
    bits = rand_bits(301, F, nbits)
    sigma = o.sigma_for_ebn0(sigma_db) if sigma_db is not None else 0.0
    _, bits_ref, cnt_ref = o.loopback(bits, sigma=sigma, seed=9, frame0=17, threads=4)
    packed = orc.pack_bits(bits)
    n0 = m.launch_count
    got = m.loopback_packed(packed, nbits, sigma=sigma, seed=9, frame0=17)
    assert np.array_equal(got["packed"], orc.pack_bits(bits_ref)), case
    assert np.array_equal(orc.unpack_bits(got["packed"], bits_ref.shape[1]), bits_ref)
    assert (got["errors"], got["compared"]) == cnt_ref
    if case == "fused" and not chunk:
        assert m.launch_count - n0 <= 4, "unpack + fused loopback kernel + pack (+ the NCO table)"
    # pad bits of the input rows are ignored
    if nbits % 8:
        dirty = packed.copy()
        dirty[:, -1] |= (1 << (8 - nbits % 8)) - 1
        again = m.loopback_packed(dirty, nbits, sigma=sigma, seed=9, frame0=17)
        assert np.array_equal(again["packed"], got["packed"])


def test_loopback_packed_device_pointers_and_bank(pkg, orc):
    """Device-resident packed rows run in place (no copies); a multi-carrier bank through the chunked host form."""
    import torch

    kw = path_kwargs("qpsk", sps=8)
    m, o = make(pkg, orc, **kw)
    F, nbits = 6, 2 * 640
    bits = rand_bits(302, F, nbits)
    _, bits_ref, cnt_ref = o.loopback(bits)
    d_in = torch.from_numpy(orc.pack_bits(bits)).cuda()
    K = m.decided_symbols(m.frame_samples(nbits))
    d_out = torch.full((F, (2 * K + 7) // 8), 0xEE, dtype=torch.uint8, device="cuda")
    cnt = m.loopback_packed_into(d_in, F, nbits, d_out)
    assert np.array_equal(d_out.cpu().numpy(), orc.pack_bits(bits_ref))
    assert cnt == cnt_ref
    assert m.loopback_packed_into(d_in, F, nbits, None) == cnt_ref  # counters only
    hz = [1100 + 333 * c for c in range(4)]
    mb = pkg.Modem(**kw)
    mb.set_channels([pkg.sample_freq(h, 10000) for h in hz], 3)
    bits2 = rand_bits(303, 12, 2 * 600)
    import os
    os.environ["MODEM_GPU_PACKED_CHUNK"] = "5"  # not a divisor of the channel size: the chunk is cut to one
    try:
        mb2 = pkg.Modem(**kw)
        mb2.set_channels([pkg.sample_freq(h, 10000) for h in hz], 3)
        got = mb2.loopback_packed(orc.pack_bits(bits2), bits2.shape[1])
    finally:
        del os.environ["MODEM_GPU_PACKED_CHUNK"]
    for c in range(4):
        oc = orc.OraclePath(**dict(kw, carrier_hz=hz[c]))
        _, b_ref, _ = oc.loopback(bits2[3 * c: 3 * c + 3])
        assert np.array_equal(got["packed"][3 * c: 3 * c + 3], orc.pack_bits(b_ref)), f"channel {c}"
    assert got["errors"] == 0
