"""Times the fused loopback kernel at C2 (4096 frames x 65536 samples, device buffers, CUDA events) for each tuning
variant named on the command line (MODEM_GPU_FUSED_VARIANT; one subprocess each, the library reads it once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--one":
    for v in sys.argv[1:]:
        env = dict(os.environ, MODEM_GPU_FUSED_VARIANT=v)
        subprocess.run([sys.executable, __file__, "--one"], env=env)
    sys.exit(0)
import numpy as np, torch
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB, F = 16384, int(os.environ.get("TF_FRAMES", "4096"))
m = pkg.Modem(**kw)
st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
L = m.frame_samples(NB); K = m.decided_symbols(L)
bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8, device="cuda")
tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
sym = torch.empty((F, K), dtype=torch.uint8, device="cuda"); out = torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
res = {}
for name, txbuf in (("store", tx), ("nostore", None)):
    for _ in range(3): m.loopback_device_into(bits, F, NB, cnt, tx=txbuf, sym=sym, bits_out=out)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(10)]
    for e in ev:
        e[0].record(st); m.loopback_device_into(bits, F, NB, cnt, tx=txbuf, sym=sym, bits_out=out); e[1].record(st)
    torch.cuda.synchronize()
    res[name] = float(np.median([e[0].elapsed_time(e[1]) for e in ev]))
ok = bool((out == bits[:, : 2 * K]).all()) and int(cnt[0]) == 0
print(f"variant {os.environ.get('MODEM_GPU_FUSED_VARIANT', '0')}: store {res['store']*1e3:.1f} us ({F*L*8.625/res['store']/1e6/6552.3:.3f} of HBM), "
      f"no TX store {res['nostore']*1e3:.1f} us, bits ok {ok}, fpb env {os.environ.get('MODEM_GPU_RX_FPB')}", flush=True)
m.close()
