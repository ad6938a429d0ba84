"""Do concurrent PCIe copies and the chunk kernels of the host-buffer pipeline slow each other down?
Two copy streams move 4 MiB blocks in and out continuously while a third stream runs the kernels of a 256-frame chunk
back to back; reports the per-chunk kernel time and the per-block copy time, with and without the other party."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB, F = 16384, 256
blk = 4 << 20
h_in = torch.empty(blk, dtype=torch.uint8).pin_memory(); h_out = torch.empty(blk, dtype=torch.uint8).pin_memory()
d_in = torch.empty(blk, dtype=torch.uint8, device="cuda"); d_o = torch.empty(blk, dtype=torch.uint8, device="cuda")
s_in, s_out, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
# warm the link and the pinned buffers first: the first few hundred blocks of a process are slower
for i in range(int(os.environ.get("WARM_BLOCKS", "400"))):
    with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s_out): h_out.copy_(d_o, non_blocking=True)
torch.cuda.synchronize()
MODES = [("TX + RX kernels", 0), ("fused kernel", 0), ("fused kernel, no TMEM", pkg.FLAG_NO_TMEM), ("TX + RX kernels", 0), ("fused kernel", 0)]
for mode, flags in MODES:
    m = pkg.Modem(flags=flags, **kw)
    m.set_stream(s_k.cuda_stream)
    L = m.frame_samples(NB); K = m.decided_symbols(L)
    bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8, device="cuda")
    tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    out = torch.empty((F, 2 * K), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    def kern():
        if mode.startswith("fused"):
            m.loopback_device_into(bits, F, NB, cnt, bits_out=out)
        else:
            m.modulate_into(bits, F, NB, tx=tx)
            m.demodulate_count_into(tx, F, L, bits, NB, cnt, bits=out)
    for _ in range(3): kern()
    torch.cuda.synchronize()
    res = {}
    for copies in (False, True):
        for kernels in (False, True):
            if not copies and not kernels: continue
            torch.cuda.synchronize()
            ce_in, ce_out, ke = [], [], []
            n_blocks = 40
            if copies:
                for i in range(n_blocks):
                    with torch.cuda.stream(s_in):
                        a = ev(); a.record(); d_in.copy_(h_in, non_blocking=True); b = ev(); b.record(); ce_in.append((a, b))
                    with torch.cuda.stream(s_out):
                        a = ev(); a.record(); h_out.copy_(d_o, non_blocking=True); b = ev(); b.record(); ce_out.append((a, b))
            if kernels:
                with torch.cuda.stream(s_k):
                    for i in range(40):
                        a = ev(); a.record(s_k); kern(); b = ev(); b.record(s_k); ke.append((a, b))
            torch.cuda.synchronize()
            med = lambda xs: float(np.median([a.elapsed_time(b) for a, b in xs[5:-5]])) * 1e3 if xs else float("nan")
            res[(copies, kernels)] = (med(ke), med(ce_in), med(ce_out))
    print(f"{mode:28s} chunk kernels alone {res[(False, True)][0]:6.1f} us, beside copies {res[(True, True)][0]:6.1f} us | 4 MiB in/out alone "
          f"{res[(True, False)][1]:6.1f}/{res[(True, False)][2]:6.1f} us, beside kernels {res[(True, True)][1]:6.1f}/{res[(True, True)][2]:6.1f} us", flush=True)
    m.close()
