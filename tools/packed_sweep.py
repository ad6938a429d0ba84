#!/usr/bin/env python
"""Chunk-size sweep of modem_gpu_loopback_packed at C2 (host pinned buffers, CUDA events): MODEM_GPU_PACKED_CHUNK is read
when the context is created, so every point makes its own context.  usage: packed_sweep.py [chunk ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
F, NBITS = 4096, 16384
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
bits = np.random.default_rng(1).integers(0, 2, (F, NBITS), dtype=np.uint8)
h_pk = torch.from_numpy(np.packbits(bits, axis=1)).pin_memory()
for chunk in [int(c) for c in sys.argv[1:]] or [256, 512, 1024, 2048, 4096]:
    os.environ["MODEM_GPU_PACKED_CHUNK"] = str(chunk)
    m = pkg.Modem(**kw)
    stream = torch.cuda.current_stream()
    m.set_stream(stream.cuda_stream)
    K = m.decided_symbols(m.frame_samples(NBITS))
    h_out = torch.zeros((F, (2 * K + 7) // 8), dtype=torch.uint8).pin_memory()
    ms = []
    for i in range(8):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        cnt = m.loopback_packed_into(h_pk, F, NBITS, h_out)
        b.record(stream)
        torch.cuda.synchronize()
        assert cnt == (0, F * K * 2)
        if i >= 3:
            ms.append(a.elapsed_time(b))
    print(f"chunk {chunk:5d} frames: {np.mean(ms):.4f} ms per step ({F * 65536 / np.mean(ms) / 1e3:.0f} Msamples/s), min {min(ms):.4f}", flush=True)
    m.close()
