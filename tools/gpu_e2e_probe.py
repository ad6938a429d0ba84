"""host-buffer loopback (pinned) through the C ABI: time per call for a few pipeline chunk sizes"""
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
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); n0 = m.launch_count
    err, cmp_ = m.loopback_into(h_bits, F, NB, bits_out=h_out)
    b.record(st); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"call {i}: events {a.elapsed_time(b):.3f} ms, wall {1e3*(t1-t0):.3f} ms, launches {m.launch_count-n0}, errors {err}", flush=True)
