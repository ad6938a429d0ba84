"""host-buffer loopback variants: with / without outputs"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB, F = 16384, 4096
m = pkg.Modem(**kw)
st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
L = m.frame_samples(NB); K = m.decided_symbols(L)
h_bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8).pin_memory()
h_out = torch.empty((F, 2 * K), dtype=torch.uint8).pin_memory()
h_sym = torch.empty((F, K), dtype=torch.uint8).pin_memory()
for name, kwargs in (("bits_out", dict(bits_out=h_out)), ("none", dict()), ("sym", dict(sym=h_sym))):
    ts = []
    for i in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        err, cmp_ = m.loopback_into(h_bits, F, NB, **kwargs)
        b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, " ".join(f"{t:.3f}" for t in ts[1:]), flush=True)
