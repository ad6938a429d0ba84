#!/usr/bin/env python
"""Secondary measurements of the rows either side of the hot path on one GPU, device-resident
buffers, CUDA events: stateful mappers (TX), the modulate binary's real output, the demodulate binary's i16 -> lock ->
decide path.  One JSON object per line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
import ctypes as C  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
L_ = pkg.lib()
F, LSAMP, SPS = 4096, 65536, 8
STEPS = 5


def timeit(fn, st):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for i in range(3 + STEPS):
        if i >= 3:
            ev[i - 3][0].record(st)
        fn()
        if i >= 3:
            ev[i - 3][1].record(st)
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def stateful(scheme):
    m = pkg.Modem(scheme=scheme, baud_rate=1250, sample_rate=10000, carrier_hz=2500)
    st = torch.cuda.current_stream()
    m.set_stream(st.cuda_stream)
    nbits = LSAMP // SPS * m.bps
    bits = torch.randint(0, 2, (F, nbits), dtype=torch.uint8, device="cuda")
    tx = torch.empty((F, LSAMP, 2), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: m.modulate_into(bits, F, nbits, tx=tx), st)
    n = F * LSAMP
    bps_bytes = 8 + m.bps / SPS + (8 / SPS if scheme in ("bfsk", "mfsk", "dqpsk", "dbpsk") else 0)  # + state write/read
    print(json.dumps({"row": f"stateful TX {scheme}", "ms": round(ms, 4), "Msamples_s": round(n / ms / 1e3),
                      "bytes_per_sample": bps_bytes, "GBs": round(n * bps_bytes / ms / 1e6), "frac": round(n * bps_bytes / ms / 1e6 / PEAK, 3)}), flush=True)
    m.close()


def modulate_real():
    m = pkg.Modem(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500)
    st = torch.cuda.current_stream()
    m.set_stream(st.cuda_stream)
    nbits = LSAMP // SPS * 2
    bits = torch.randint(0, 2, (F, nbits), dtype=torch.uint8, device="cuda")
    P = 39
    out = torch.empty((F, P + LSAMP), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: m._ck(L_.modem_gpu_modulate_real(m._ctx, bits.data_ptr(), F, nbits, P, 1.0, out.data_ptr())), st)
    n = F * LSAMP
    print(json.dumps({"row": "modulate binary: sync tone + real f32 output (qpsk)", "ms": round(ms, 4), "Msamples_s": round(n / ms / 1e3),
                      "bytes_per_sample": 4.25, "GBs": round(n * 4.25 / ms / 1e6), "frac": round(n * 4.25 / ms / 1e6 / PEAK, 3)}), flush=True)
    m.close()
    return out


def demodulate_real(wire_f32, fmt):
    lp = pkg.lowpass_taps()
    P = 39
    m = pkg.Modem(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, rx_taps=lp,
                  decision_delay=35, slicer_gain=float(np.float32(lp.sum()) * np.float32(8000.0 if fmt == "i16" else 1.0)))
    st = torch.cuda.current_stream()
    m.set_stream(st.cuda_stream)
    x = (wire_f32 * 8000.0).round().to(torch.int16) if fmt == "i16" else wire_f32
    Fx, Lx = x.shape
    lock = 0 if fmt == "f32-nolock" else 64
    Lr = Lx - lock
    K = m.decided_symbols(Lr)
    sym = torch.empty((Fx, K), dtype=torch.uint8, device="cuda")
    out = torch.empty((Fx, K * 2), dtype=torch.uint8, device="cuda")
    po = torch.empty(Fx, dtype=torch.float32, device="cuda")
    code = pkg.capi.SAMPLES_I16 if fmt == "i16" else pkg.capi.SAMPLES_F32
    ms = timeit(lambda: m._ck(L_.modem_gpu_demodulate_real(m._ctx, x.data_ptr(), code, Fx, Lx, lock, None, 0, po.data_ptr(),
                                                           sym.data_ptr(), out.data_ptr(), None, None)), st)
    n = Fx * Lr
    b = (2 if fmt == "i16" else 4) + 0.375
    print(json.dumps({"row": f"demodulate binary: {fmt} wire -> Hilbert/PLL lock({lock}) -> low-pass -> decide", "ms": round(ms, 4),
                      "Msamples_s": round(n / ms / 1e3), "bytes_per_sample": b, "GBs": round(n * b / ms / 1e6),
                      "frac": round(n * b / ms / 1e6 / PEAK, 3), "po_mean": float(po.mean())}), flush=True)
    m.close()


def demodulate_real_rates(fmt):
    """The same path at the reference's own rates (sr 10000 / baud 220 -> sps 45, carrier 1000 Hz, preamble of
    sr / cf * 20 - 1 = 199 samples: modulate.rs:44-58,118-126): 4096 frames x (199 + 65520) samples."""
    lp = pkg.lowpass_taps()
    sps, P, nsym = 45, 199, 1456
    tx = pkg.Modem(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=1000)
    st = torch.cuda.current_stream()
    tx.set_stream(st.cuda_stream)
    bits = torch.randint(0, 2, (F, 2 * nsym), dtype=torch.uint8, device="cuda")
    wire = torch.empty((F, P + nsym * sps), dtype=torch.float32, device="cuda")
    tx._ck(L_.modem_gpu_modulate_real(tx._ctx, bits.data_ptr(), F, 2 * nsym, P, 1.0, wire.data_ptr()))
    torch.cuda.synchronize()
    tx.close()
    m = pkg.Modem(scheme="qpsk", baud_rate=220, sample_rate=10000, carrier_hz=1000, rx_taps=lp,
                  decision_delay=(P - 64) + 31 + sps // 2, slicer_gain=float(np.float32(lp.sum()) * np.float32(8000.0 if fmt == "i16" else 1.0)))
    m.set_stream(st.cuda_stream)
    x = (wire * 8000.0).round().to(torch.int16) if fmt == "i16" else wire
    Fx, Lx = x.shape
    Lr = Lx - 64
    K = m.decided_symbols(Lr)
    sym = torch.empty((Fx, K), dtype=torch.uint8, device="cuda")
    out = torch.empty((Fx, K * 2), dtype=torch.uint8, device="cuda")
    po = torch.empty(Fx, dtype=torch.float32, device="cuda")
    code = pkg.capi.SAMPLES_I16 if fmt == "i16" else pkg.capi.SAMPLES_F32
    ms = timeit(lambda: m._ck(L_.modem_gpu_demodulate_real(m._ctx, x.data_ptr(), code, Fx, Lx, 64, None, 0, po.data_ptr(),
                                                           sym.data_ptr(), out.data_ptr(), None, None)), st)
    errors = int((out != bits[:, : 2 * K]).sum())
    n = Fx * Lr
    print(json.dumps({"row": f"demodulate binary at the reference's own rates (sps 45): {fmt} wire -> Hilbert/PLL lock(64) -> low-pass -> decide",
                      "ms": round(ms, 4), "Msamples_s": round(n / ms / 1e3), "bit_errors": errors, "po_mean": float(po.mean())}), flush=True)
    m.close()


def fullrate(fmt, Fx=1024):
    """What iterating the reference's Demodulator yields: the filtered (I,Q) for EVERY input sample (demodulator.rs:44-55),
    2 x 64 MACs per sample -- FP32-lane bound by construction (256 lane-ops per sample: <= 145 GS/s exact)."""
    lp = pkg.lowpass_taps()
    m = pkg.Modem(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, rx_taps=lp, decision_delay=35,
                  slicer_gain=float(np.float32(lp.sum())))
    st = torch.cuda.current_stream()
    m.set_stream(st.cuda_stream)
    n = Fx * LSAMP
    if fmt == "c32":
        x = torch.randn((Fx, LSAMP, 2), dtype=torch.float32, device="cuda")
        filt = torch.empty((Fx, LSAMP, 2), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: m.demodulate_into(x, Fx, LSAMP, filt=filt), st)
        b = 16
    else:
        x = (torch.randn((Fx, LSAMP + 64), dtype=torch.float32, device="cuda") * 3000).to(torch.int16)
        filt = torch.empty((Fx, LSAMP, 2), dtype=torch.float32, device="cuda")
        po = torch.empty(Fx, dtype=torch.float32, device="cuda")
        ms = timeit(lambda: m._ck(L_.modem_gpu_demodulate_real(m._ctx, x.data_ptr(), pkg.capi.SAMPLES_I16, Fx, LSAMP + 64, 64, None, 0,
                                                               po.data_ptr(), None, None, None, filt.data_ptr())), st)
        b = 10
    print(json.dumps({"row": f"Demodulator iterator, full-rate (I,Q) out, {fmt} in ({Fx} frames)", "ms": round(ms, 4),
                      "Msamples_s": round(n / ms / 1e3), "bytes_per_sample": b, "GBs": round(n * b / ms / 1e6),
                      "frac_hbm": round(n * b / ms / 1e6 / PEAK, 3), "frac_fp32_lane_bound": round(n / ms / 1e3 / 145000, 3)}), flush=True)
    m.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["stateful", "real", "fullrate"]
    if "stateful" in which:
        for s in ("16cpfsk", "msk", "bfsk", "mfsk", "dqpsk"):
            stateful(s)
    if "real" in which:
        w = modulate_real()
        demodulate_real(w, "i16")
        demodulate_real(w, "f32")
        demodulate_real(w, "f32-nolock")
        del w
        torch.cuda.empty_cache()
        demodulate_real_rates("i16")
        demodulate_real_rates("f32")
    if "fullrate" in which:
        fullrate("c32")
        fullrate("i16")
