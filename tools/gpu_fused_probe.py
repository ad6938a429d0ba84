"""fused loopback kernel at chunk-sized frame counts, device buffers (CUDA events), then the host pipeline"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB = 16384
for F in ((4096,) if os.environ.get("MODEM_GPU_RX_FPB") else (64, 128, 256, 512, 1024, 4096)):
    m = pkg.Modem(**kw)
    st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
    L = m.frame_samples(NB); K = m.decided_symbols(L)
    bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8, device="cuda")
    tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    sym = torch.empty((F, K), dtype=torch.uint8, device="cuda"); out = torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    for _ in range(3): m.loopback_device_into(bits, F, NB, cnt, tx=tx, sym=sym, bits_out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n0 = m.launch_count
    ev[0].record(st)
    for _ in range(10): m.loopback_device_into(bits, F, NB, cnt, tx=tx, sym=sym, bits_out=out)
    ev[1].record(st); torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1]) / 10
    print(f"F={F}: {t*1e3:.1f} us per call, {F*L/t/1e6:.0f} GS/s, launches/call {(m.launch_count-n0)/10}, errors {int(cnt[0])}", flush=True)
    m.close()
