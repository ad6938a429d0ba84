"""time the two hot kernels at C2 (device buffers, CUDA events); prints rx/tx ms"""
import sys, torch, numpy as np
sys.path.insert(0, "."); import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
m = pkg.Modem(rx_taps=lp, decision_delay=35, slicer_gain=float(np.float32(lp.sum())))
st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
F, NB, L = 4096, 16384, 65536; K = m.decided_symbols(L)
bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8, device="cuda"); tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
sym = torch.empty((F, K), dtype=torch.uint8, device="cuda"); out = torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(23)]
for e in ev:
    e[0].record(st); m.modulate_into(bits, F, NB, tx=tx); e[1].record(st)
    m.demodulate_count_into(tx, F, L, bits, NB, cnt, sym=sym, bits=out); e[2].record(st)
torch.cuda.synchronize()
ok = bool((out == bits[:, :2 * K]).all())
print("%s tx %.4f ms  rx %.4f ms  (roundtrip ok=%s)" % (sys.argv[1] if len(sys.argv) > 1 else "", np.mean([e[0].elapsed_time(e[1]) for e in ev[3:]]), np.mean([e[1].elapsed_time(e[2]) for e in ev[3:]]), ok))
