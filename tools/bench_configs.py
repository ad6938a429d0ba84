#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract line): the other BASELINE.json configs on one GPU.
Prints one JSON object per config: kernel times (CUDA events), Msamples/s, achieved GB/s."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
SPS, BPS, NSYM = 8, 2, 8192
NBITS, L = NSYM * BPS, NSYM * SPS


def run(name, F, shaped, flags=0, sigma_db=None, channels=0, steps=5, sps=8, loop=False):
    global SPS, NSYM, NBITS, L
    SPS = sps
    NSYM = 65536 // sps
    NBITS, L = NSYM * BPS, NSYM * SPS
    lp = pkg.lowpass_taps()
    if shaped:
        rrc = pkg.rrc_taps(16, 8, 0.35)
        kw = dict(tx_taps=rrc, rx_taps=rrc, decision_delay=128, slicer_gain=1.0)
    else:
        kw = dict(rx_taps=lp, decision_delay=31 + sps // 2, slicer_gain=float(np.float32(lp.sum())))
    br = {8: 1250, 45: 220}[sps]
    m = pkg.Modem(scheme="qpsk", baud_rate=br, sample_rate=10000, carrier_hz=2500 if sps == 8 else 1000, flags=flags, **kw)
    st = torch.cuda.current_stream()
    m.set_stream(st.cuda_stream)
    if channels:
        m.set_channels([pkg.sample_freq(1000 + (3000 * c) // 1024, 10000) for c in range(channels)], F // channels)
    K = m.decided_symbols(L)
    gen = torch.Generator(device="cuda").manual_seed(1)
    bits = torch.randint(0, 2, (F, NBITS), dtype=torch.uint8, device="cuda", generator=gen)
    tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    out = torch.empty((F, K * BPS), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    sigma = m.sigma_for_ebn0(sigma_db) if sigma_db is not None else 0.0
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    if loop:  # the loopback entry: ONE fused kernel where the shape allows
        for i in range(3 + steps):
            cnt.zero_()
            e = ev[i - 3] if i >= 3 else None
            if e: e[0].record(st)
            m.loopback_device_into(bits, F, NBITS, cnt, tx=tx, sym=sym, bits_out=out)
            if e: e[1].record(st)
        torch.cuda.synchronize()
        ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
        b = 8 + BPS / SPS + (1 + BPS) / SPS
        print(json.dumps({"config": name, "frames": F, "loop_ms": round(ms, 4), "loopback_Msamples_s": round(F * L / ms / 1e3, 0),
                          "GBs": round(F * L * b / ms / 1e6, 0), "frac": round(F * L * b / ms / 1e6 / PEAK, 3), "bytes_per_sample": b,
                          "errors": int(cnt[0]), "bits": int(cnt[1])}), flush=True)
        m.close()
        del tx
        torch.cuda.empty_cache()
        return
    for i in range(3 + steps):
        cnt.zero_()
        e = ev[i - 3] if i >= 3 else None
        if e: e[0].record(st)
        m.modulate_into(bits, F, NBITS, tx=tx)
        if e: e[1].record(st)
        m.demodulate_count_into(tx, F, L, bits, NBITS, cnt, sym=sym, bits=out, sigma=sigma, seed=0xA5A5)
        if e: e[2].record(st)
    torch.cuda.synchronize()
    tx_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    rx_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    n = F * L
    res = {"config": name, "frames": F, "tx_ms": round(tx_ms, 4), "rx_ms": round(rx_ms, 4),
           "loopback_Msamples_s": round(n / (tx_ms + rx_ms) / 1e3, 0),
           "tx_GBs": round(n * (8 + BPS / SPS) / tx_ms / 1e6, 0), "tx_frac": round(n * (8 + BPS / SPS) / tx_ms / 1e6 / PEAK, 3),
           "rx_GBs": round(n * (8 + (1 + 2 * BPS) / SPS) / rx_ms / 1e6, 0), "rx_frac": round(n * (8 + (1 + 2 * BPS) / SPS) / rx_ms / 1e6 / PEAK, 3),
           "errors": int(cnt[0]), "bits": int(cnt[1]), "ber": float(cnt[0]) / max(int(cnt[1]), 1)}
    print(json.dumps(res), flush=True)
    m.close()
    del tx
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c2f", "c3", "c3f", "c2n", "c3n", "c5"]
    FUSED = pkg.FLAG_FUSED_MAC
    for w in which:
        if w == "c1": run("C1 rates (sr 10000 / baud 220 -> sps 45, 1000 Hz), 4096 frames x 65520 samples (tx_rect_fast + rx_dec kernels)", 4096, False, sps=45)
        if w == "c1l": run("C1 rates, the loopback entry (ONE fused kernel: rx_dec_kernel<..., TXF>)", 4096, False, sps=45, loop=True)
        if w == "c2l": run("C2, the loopback entry (ONE fused kernel: rx_fast_kernel<..., TXF>)", 4096, False, loop=True)
        if w == "c2": run("C2 rect+lp64 exact", 4096, False)
        if w == "c2f": run("C2 rect+lp64 fused-MAC", 4096, False, flags=FUSED)
        if w == "c3": run("C3 rrc129 exact (16384 frames = 2^30 samples)", 16384, True, steps=3)
        if w == "c3f": run("C3 rrc129 fused-MAC", 16384, True, flags=FUSED, steps=3)
        if w == "c2n": run("C2 + AWGN 6 dB fused into RX load", 4096, False, sigma_db=6.0, steps=3)
        if w == "c3n": run("C4-style: rrc129 + AWGN 4 dB (Monte-Carlo rate)", 4096, True, sigma_db=4.0, steps=3)
        if w == "c5": run("C5-style bank: 128 carriers x 32 frames on one GPU", 4096, False, channels=128)
