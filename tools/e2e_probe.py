"""Host-buffer loopback (the e2e number of bench.py) at C2 under different pipeline settings, one subprocess each
(the library reads its MODEM_GPU_* knobs at context creation):  python tools/e2e_probe.py [name=ENV=VAL,ENV=VAL ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--one":
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition("=")
        env = dict(os.environ, E2E_NAME=name)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        subprocess.run([sys.executable, __file__, "--one"], env=env)
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB, F = 16384, int(os.environ.get("E2E_FRAMES", "4096"))
m = pkg.Modem(**kw)
st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
L = m.frame_samples(NB); K = m.decided_symbols(L)
h_bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8).pin_memory()
h_out = torch.zeros((F, 2 * K), dtype=torch.uint8).pin_memory()
ts = []
for i in range(7):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    err, cmp_ = m.loopback_into(h_bits, F, NB, bits_out=h_out)
    b.record(st); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ok = bool((h_out == h_bits[:, : 2 * K]).all()) and err == 0 and cmp_ == F * 2 * K
print(f"{os.environ.get('E2E_NAME', '?'):28s} median {np.median(ts[2:]):.3f} ms  min {min(ts[2:]):.3f}  ({F*L/np.median(ts[2:])/1e6:.0f} GS/s)  ok {ok}", flush=True)
m.close()
