"""Why is the host-buffer pipeline slower with the (cheaper) fused kernel?  Runs the e2e call in a loop for ~2 s per
mode while sampling nvidia-smi clocks, optionally with a background kernel that keeps one SM per ... busy."""
import os, subprocess, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
lp = pkg.lowpass_taps()
kw = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, decision_delay=35, slicer_gain=float(lp.sum()), rx_taps=lp)
NB, F = 16384, 4096
h_bits = torch.randint(0, 2, (F, NB), dtype=torch.uint8).pin_memory()
st = torch.cuda.current_stream()
side = torch.cuda.Stream()
big_a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda"); big_b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
def smi():
    q = "clocks.sm,clocks.mem,clocks.gr,clocks.video,pstate,power.draw,utilization.gpu,utilization.memory,pcie.link.gen.current"
    return subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for mode, env, bg in (("two kernels", "1", None), ("fused", "0", None), ("fused + background HBM copy stream", "0", "copy"), ("fused + background spin", "0", "spin"), ("two kernels", "1", None), ("fused", "0", None)):
    os.environ["MODEM_GPU_PIPE_TWO_KERNELS"] = env
    os.environ["MODEM_GPU_PIPE_RAMP"] = "0"
    m = pkg.Modem(**kw)
    m.set_stream(st.cuda_stream)
    L = m.frame_samples(NB); K = m.decided_symbols(L)
    h_out = torch.zeros((F, 2 * K), dtype=torch.uint8).pin_memory()
    ts, samples = [], []
    t_end = time.time() + 1.5
    while time.time() < t_end:
        if bg == "copy":
            with torch.cuda.stream(side):
                for _ in range(8): big_b.copy_(big_a, non_blocking=True)  # ~0.7 ms of HBM traffic beside the call
        if bg == "spin":
            with torch.cuda.stream(side):
                torch.cuda._sleep(6_000_000)  # one CTA spinning ~3 ms
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        m.loopback_into(h_bits, F, NB, bits_out=h_out)
        b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        if len(ts) % 100 == 0: samples.append(smi())
    print(f"{mode:36s} calls {len(ts):4d}  median {np.median(ts):.3f} ms  first10 {np.median(ts[:10]):.3f}  last100 {np.median(ts[-100:]):.3f}  | smi {samples[-1] if samples else smi()}", flush=True)
    m.close()
