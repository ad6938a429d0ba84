import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (ctypes binding of oracle/modem_oracle.c)."""
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def pkg():
    """The product binding (rust-modem_b200/, loaded as rust_modem_b200)."""
    import __graft_entry__ as g

    p = g.load_package()
    p.build_library()
    p.lib()
    return p


MEMORYLESS = ["bask", "bpsk", "qpsk", "qam16", "qam256", "16psk", "oqpsk", "dcqpsk", "16apsk"]


def path_kwargs(scheme="qpsk", sps=8, n_rx=64, shaped=False, orc=None, **over):
    """Keyword set accepted by both oracle.OraclePath and Modem for one configuration."""
    import numpy as np
    from oracle import oracle as O

    sr = 10000 if sps != 8 else 10000
    br = {8: 1250, 45: 220, 4: 2500, 10: 1000, 5: 2000, 3: 3333}[sps]
    kw = dict(scheme=scheme, baud_rate=br, sample_rate=sr, carrier_hz=2500 if sps == 8 else 1000)
    q_off = sps // 2 if scheme == "oqpsk" else 0
    if shaped:
        rrc = O.rrc_taps(16, sps, 0.35)
        kw.update(tx_taps=rrc, rx_taps=rrc, decision_delay=len(rrc) - 1, slicer_gain=1.0)
    else:
        lp = O.lowpass_taps()
        kw.update(tx_taps=None, rx_taps=lp, decision_delay=31 + sps // 2, slicer_gain=float(np.float32(lp.sum())))
    kw.update(over)
    return kw
