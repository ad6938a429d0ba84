"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces every frozen vector bit for bit (guards the oracle and the
host libm against drift).  GPU: the CUDA path reproduces the same vectors through the C ABI
without the oracle being involved at all."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

NAMES = sorted(G.CASES)


def load(name):
    return dict(np.load(os.path.join(G.HERE, name + ".npz")))


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(orc, name):
    ref = load(name)
    _, out = G.build(name)
    assert sorted(out) == sorted(ref)
    for k in ref:
        assert same_bits(np.asarray(out[k]), ref[k]), f"{name}:{k} drifted from the committed fixture"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(pkg, name):
    scheme, sps, shaped, extra, F, nsym, ebn0 = G.CASES[name]
    from conftest import path_kwargs
    ref = load(name)
    m = pkg.Modem(**path_kwargs(scheme, sps=sps, shaped=shaped, **extra))
    tx, iq = m.modulate(ref["bits"], want_iq=True)
    assert same_bits(iq, ref["iq"]) and same_bits(tx, ref["tx"]), name
    if ebn0 is None:
        out = m.demodulate(ref["tx"], want_filt=True)
    else:
        sigma = float(ref["sigma"])
        assert m.sigma_for_ebn0(ebn0) == sigma
        assert same_bits(m.awgn(ref["tx"], sigma, seed=0xA5A5, frame0=7), ref["noisy"])
        out = m.demodulate(ref["tx"], want_filt=True, sigma=sigma, seed=0xA5A5, frame0=7)
        lb = m.loopback(ref["bits"], sigma=sigma, seed=0xA5A5, frame0=7)
        assert (lb["errors"], lb["compared"]) == tuple(int(x) for x in ref["counters"])
        assert int(ref["counters"][0]) > 0
    assert same_bits(out["filt"], ref["filt"]), name
    assert np.array_equal(out["sym"], ref["sym"]) and np.array_equal(out["bits"], ref["dec_bits"]), name
