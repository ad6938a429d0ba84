#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU oracle.

The reference (Rust, 2016 nightly) cannot be built in this image, so these vectors are the
oracle's own outputs, frozen: they pin the oracle (and glibc's sinf/cosf/logf) against drift
and give the GPU tests a fixed target that does not depend on running the oracle at all.
The reference's own known-answer vectors are restated separately in tests/test_oracle_kat.py.

    python tests/golden/make_golden.py        # rewrites the fixtures (run in the dev container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import path_kwargs  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (scheme, sps, shaped, extra kwargs, frames, symbols per frame, noise Eb/N0 dB or None)
    "qpsk_rect_lp64": ("qpsk", 8, False, {}, 3, 200, None),
    "qpsk_rrc129": ("qpsk", 8, True, {}, 3, 200, None),
    "qpsk_default_rates": ("qpsk", 45, False, {}, 2, 40, None),       # sr 10000 / baud 220 / cf 1000 (modulate.rs:44-58)
    "qam16_rect_lp64": ("qam16", 8, False, {}, 2, 160, None),
    "16apsk_rrc": ("16apsk", 8, True, {}, 2, 160, None),
    "oqpsk_rect": ("oqpsk", 8, False, {}, 2, 160, None),
    "dcqpsk_rect": ("dcqpsk", 8, False, {}, 2, 160, None),
    "bpsk_phase_offset": ("bpsk", 8, False, {"phase_offset": 0.25, "sample0": 89}, 2, 160, None),
    "qpsk_rrc129_awgn4dB": ("qpsk", 8, True, {}, 4, 256, 4.0),
}


def build(name):
    scheme, sps, shaped, extra, F, nsym, ebn0 = CASES[name]
    kw = path_kwargs(scheme, sps=sps, shaped=shaped, **extra)
    o = O.OraclePath(**kw)
    seed = sum(map(ord, name))
    bits = np.random.default_rng(seed).integers(0, 2, (F, nsym * o.bps), dtype=np.uint8)
    tx, iq = o.modulate(bits, want_iq=True)
    out = {"bits": bits, "tx": tx, "iq": iq}
    if ebn0 is None:
        filt, sym, dec = o.demodulate(tx)
        out.update(filt=filt, sym=sym, dec_bits=dec)
    else:
        sigma = o.sigma_for_ebn0(ebn0)
        noisy = o.awgn(tx, sigma, seed=0xA5A5, frame0=7)
        filt, sym, dec = o.demodulate(noisy)
        s2, d2, cnt = o.loopback(bits, sigma=sigma, seed=0xA5A5, frame0=7)
        assert np.array_equal(sym, s2) and np.array_equal(dec, d2)
        out.update(noisy=noisy, filt=filt, sym=sym, dec_bits=dec, sigma=np.float32(sigma),
                   counters=np.array(cnt, np.uint64))
    return kw, out


def main():
    for name in CASES:
        _, out = build(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
