#!/usr/bin/env python
"""bench.py -- loopback throughput of the batched modulate -> demodulate path on B200.

Workload (BASELINE.json configs[1], "C2"): 4096 frames x 65536 complex f32 samples per GPU,
QPSK, 8 samples/symbol, rectangular hold (the reference's TX pulse) + the reference's 64-tap
low-pass (src/bin/demodulate.rs:82-147), no channel noise.  One step = one pass of the hot path over
every frame: the fused loopback kernel makes the TX samples from the bits, stores them to the 2 GiB TX
buffer in HBM, demodulates them (decimate, slice, demap) and counts bit errors against the input bits.
With N > 1 ranks each rank owns its own 4096 frames ("weak" scaling, frames are independent) and ONE
tiny NCCL all-reduce sums the error counters at the end of the K timed steps (one exchange per sweep,
SURVEY.md 8e); the variant that reduces after every step is timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                          the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  Besides the contract keys it carries `configs`: the other named shapes of
BASELINE.json (C1 rates, C3 129-tap RRC, C4-style AWGN sweep, C5 carrier bank) measured in the same run at the
same N, each with its kernel times, Msamples/s and roofline fraction.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, NSYM, SPS, BPS = 4096, 8192, 8, 2
NBITS = NSYM * BPS
L = NSYM * SPS
PATH = dict(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500)
WORKLOAD = "C2: 4096 frames x 65536 complex f32 samples per GPU, QPSK, sps 8, rect hold + 64-tap low-pass, no noise"
# sources whose change invalidates the committed dram__bytes capture (profiles/traffic.json)
TRAFFIC_SOURCES = ["rust-modem_b200/csrc/rx_fast.cuh", "rust-modem_b200/csrc/common.cuh", "rust-modem_b200/csrc/loop_fused_64.cu",
                   "rust-modem_b200/csrc/tx_fast.cu", "rust-modem_b200/csrc/rx_fast_64.cu", "rust-modem_b200/csrc/rx_dec.cu"]


def path_kwargs(lowpass):
    kw = dict(PATH)
    kw.update(rx_taps=lowpass, decision_delay=31 + SPS // 2, slicer_gain=float(np.float32(lowpass.sum())))
    return kw


def config_block(world, frames, extra=None):
    """The `config` object: the same keys in both arms (--impl ours / reference)."""
    c = {"workload": WORKLOAD, "frames_per_gpu": frames, "samples_per_frame": L,
         "l2": ("per step 2 GiB of TX samples are written per GPU (16x the 126 MB L2), no flush between steps: every step's "
                "buffers are far larger than L2; a step is ONE launch of the fused loopback kernel (rx_fast_kernel<..., TXF>: "
                "makes the TX samples from the bits, stores them, demodulates them from registers; DESIGN.md 4)"),
         "parallelism": f"frames sharded over {world} GPU(s), one NCCL all-reduce of 2 u64 counters per {'sweep' if world > 1 else 'sweep (none at 1 GPU)'}"}
    if extra:
        c.update(extra)
    return c


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def source_sha():
    h = hashlib.sha256()
    for rel in TRAFFIC_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def committed_traffic():
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu capture -- only while the kernel sources are
    the ones that capture was taken from (profiles/traffic.json carries their hash); otherwise null."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return {}, "no profiles/traffic.json"
    t = json.load(open(tp))
    if t.get("source_sha") != source_sha():
        return {}, f"stale: kernel sources changed since the capture ({t.get('source_sha')} != {source_sha()})"
    return t, t.get("_source", "")


def bind_near_gpu(local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned host buffer is allocated (first touch places
    the pages): with unbound ranks the buffers of 8 ranks land wherever the scheduler ran them."""
    info = {"bound": False}
    try:
        import torch
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000") and len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip()) if os.path.exists(base + "/numa_node") else -1
        cpus = open(base + "/local_cpulist").read().strip() if os.path.exists(base + "/local_cpulist") else ""
        info.update(pci=bus, numa_node=node, local_cpulist=cpus, n_cpus=os.cpu_count())
        ids = set()
        for part in filter(None, cpus.split(",")):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        ids &= allowed
        if ids and ids != allowed:
            os.sched_setaffinity(0, ids)
            info["bound"] = True
        del torch
    except Exception as e:  # binding is best effort: a VM often exposes one node and no topology
        info["error"] = str(e)[:120]
    return info


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(frames, threads):
    """The oracle (C port of the reference's scalar CPU path) on a bounded sample of the workload."""
    from oracle import oracle as O

    lp = O.lowpass_taps()
    o = O.OraclePath(**path_kwargs(lp))
    bits = np.random.default_rng(99).integers(0, 2, (frames, NBITS), dtype=np.uint8)
    o.loopback(bits[: max(1, threads)], threads=threads, want_out=False)  # warm
    t = time.perf_counter()
    _, _, cnt = o.loopback(bits, threads=threads, want_out=False)
    dt = time.perf_counter() - t
    assert cnt[0] == 0
    return frames * L / dt / 1e6, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the Rust crate cannot be built
    here -- no rustc), all host threads.  Each step is the WHOLE C2 workload of one GPU (4096 frames) when the run then
    still ends within ~2.5 minutes, else the largest sample of it that does."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, _ = cpu_baseline(4 * threads, threads)  # Msamples/s, probe
    budget_s = 150.0
    frames = int(min(FRAMES, max(4 * threads, budget_s * rate * 1e6 / L / (args.warmup + args.steps))))
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_baseline(frames, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    sample = (f"{frames} of {FRAMES} frames x {L} samples per step" + (" (the whole workload of one GPU)" if frames == FRAMES else "") +
              f", {threads} threads (frames sharded)")
    emit({
        "impl": "reference", "metric": "loopback Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.gpus, FRAMES),
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract goes to the process's original stdout; everything else any
    library prints (e.g. NCCL's version banner at N > 1) has been routed to stderr."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


# ------------------------------------------------------------------------------------------------- other named shapes
def other_configs(pkg, torch, dist, world, rank, local_rank, peak, which, steps=5):
    """BASELINE.json configs 1, 3, 4, 5 on this rank's GPU (every rank runs the same shape on its own frames; the time
    of a row is the max over ranks, its Msamples/s the whole job's).  Device-resident buffers, CUDA events."""
    stream = torch.cuda.current_stream()
    rows = []

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def events(n):
        return [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]

    def run_two_kernels(name, note, F, sps, shaped, channels=0, fused_ok=False):
        nsym = 65536 // sps
        nbits, Lc = nsym * BPS, nsym * sps
        lp = pkg.lowpass_taps()
        if shaped:
            rrc = pkg.rrc_taps(16, 8, 0.35)
            kw = dict(tx_taps=rrc, rx_taps=rrc, decision_delay=128, slicer_gain=1.0)
        else:
            kw = dict(rx_taps=lp, decision_delay=31 + sps // 2, slicer_gain=float(np.float32(lp.sum())))
        m = pkg.Modem(scheme="qpsk", baud_rate={8: 1250, 45: 220}[sps], sample_rate=10000, carrier_hz=2500 if sps == 8 else 1000,
                      device=local_rank, **kw)
        m.set_stream(stream.cuda_stream)
        if channels:
            # 1024 carriers over 8 GPUs, all below sr/2 (modulate.rs:68): this rank's 128.  hz_c = 1000 + 3000 c / 1024 keeps
            # every carrier's double-frequency image in the low-pass stop band, so the loopback is error-free (SURVEY.md
            # 8d.5's 500 + 4 c puts the images of the lowest and highest carriers inside the pass band: genuine bit errors)
            c0 = channels * rank
            m.set_channels([pkg.sample_freq(1000 + (3000 * (c0 + c)) // 1024, 10000) for c in range(channels)], F // channels)
        K = m.decided_symbols(Lc)
        gen = torch.Generator(device="cuda").manual_seed(17 + rank)
        bits = torch.randint(0, 2, (F, nbits), dtype=torch.uint8, device="cuda", generator=gen)
        tx = torch.empty((F, Lc, 2), dtype=torch.float32, device="cuda")
        sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
        out = torch.empty((F, K * BPS), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        ev = events(steps)
        for i in range(3 + steps):
            e = ev[i - 3] if i >= 3 else None
            if e: e[0].record(stream)
            m.modulate_into(bits, F, nbits, tx=tx)
            if e: e[1].record(stream)
            m.demodulate_count_into(tx, F, Lc, bits, nbits, cnt, sym=sym, bits=out)
            if e: e[2].record(stream)
        torch.cuda.synchronize()
        tx_ms = tmax(float(np.mean([e[0].elapsed_time(e[1]) for e in ev])))
        rx_ms = tmax(float(np.mean([e[1].elapsed_time(e[2]) for e in ev])))
        n = F * Lc
        tx_b, rx_b = 8 + BPS / sps, 8 + (1 + 2 * BPS) / sps
        row = {"config": name, "note": note, "frames_per_gpu": F, "samples_per_frame": Lc,
               "tx_ms": round(tx_ms, 4), "rx_ms": round(rx_ms, 4), "ms": round(tx_ms + rx_ms, 4),
               "Msamples_s": round(world * n / (tx_ms + rx_ms) / 1e3, 0),
               "roofline": {"tx_frac": round(n * tx_b / tx_ms / 1e6 / peak, 3), "rx_frac": round(n * rx_b / rx_ms / 1e6 / peak, 3),
                            "frac": round(n * (tx_b + rx_b) / (tx_ms + rx_ms) / 1e6 / peak, 3), "bytes_per_sample": tx_b + rx_b,
                            "bound": "hbm (the 129-tap kernels are co-limited by the FP32 lanes: DESIGN.md 4)" if shaped else "hbm"},
               "errors": int(cnt[0]), "bits": int(cnt[1])}
        if fused_ok:  # the shape also runs as ONE fused kernel through the loopback entry
            cnt.zero_()
            evf = events(steps)
            n0 = m.launch_count
            for i in range(3 + steps):
                e = evf[i - 3] if i >= 3 else None
                if e: e[0].record(stream)
                m.loopback_device_into(bits, F, nbits, cnt, tx=tx, sym=sym, bits_out=out)
                if e: e[1].record(stream)
            torch.cuda.synchronize()
            f_ms = tmax(float(np.mean([e[0].elapsed_time(e[1]) for e in evf])))
            fb = 8 + BPS / sps + (1 + BPS) / sps
            row.update(fused_ms=round(f_ms, 4), fused_Msamples_s=round(world * n / f_ms / 1e3, 0),
                       fused_launches_per_step=(m.launch_count - n0) / (3 + steps), fused_errors=int(cnt[0]))
            row["roofline"].update(fused_frac=round(n * fb / f_ms / 1e6 / peak, 3), fused_bytes_per_sample=fb)
        m.close()
        del tx, bits, sym, out
        torch.cuda.empty_cache()
        return row

    def run_sweep(F):
        """C4-style: 129-tap RRC both sides, Eb/N0 0..10 dB; per sweep the frames are modulated once and demodulated 11
        times with Philox AWGN added on the fly; counters [11][2] stay on the device, one all-reduce per sweep."""
        from rust_modem_b200.capi import Comm  # noqa: F401
        nbits = NSYM * BPS
        rrc = pkg.rrc_taps(16, 8, 0.35)
        m = pkg.Modem(scheme="qpsk", baud_rate=1250, sample_rate=10000, carrier_hz=2500, tx_taps=rrc, rx_taps=rrc, decision_delay=128,
                      slicer_gain=1.0, device=local_rank)
        m.set_stream(stream.cuda_stream)
        dbs = list(range(11))
        sig = [m.sigma_for_ebn0(float(d)) for d in dbs]
        K = m.decided_symbols(L)
        bits = torch.empty((F, nbits), dtype=torch.uint8, device="cuda")
        tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
        cnt = torch.zeros((11, 2), dtype=torch.int64, device="cuda")
        ev = events(3)
        for i in range(1 + 3):
            cnt.zero_()
            e = ev[i - 1] if i >= 1 else None
            if e: e[0].record(stream)
            m.random_bits_into(bits, F, nbits, 0xA5A5 + i, frame0=rank * F)  # payload bits from Philox, in the library
            m.ber_sweep_into(bits, F, nbits, sig, cnt, seed=0xA5A5 + i, frame0=rank * F, tx=tx)
            if e: e[1].record(stream)
        torch.cuda.synchronize()
        ms = tmax(float(np.mean([e[0].elapsed_time(e[1]) for e in ev])))
        if world > 1:
            dist.all_reduce(cnt)
        c = cnt.cpu().numpy()
        import math
        n = F * L
        m.close()
        del tx
        torch.cuda.empty_cache()
        return {"config": "C4-style AWGN sweep", "note": "QPSK, sps 8, 129-tap RRC both sides, Eb/N0 0..10 dB: per sweep: Philox payload bits (library), modulate once, 11 noisy demodulations (Philox4x32-10 AWGN added while loading); steady state (warm sweep first, NCCL set-up outside the timer)",
                "frames_per_gpu": F, "samples_per_frame": L, "ms": round(ms, 3), "points": 11,
                "Msamples_s": round(world * 11 * n / ms / 1e3, 0), "Gbit_s": round(world * 11 * F * K * BPS / ms / 1e6, 2),
                "ber_0dB": float(c[0, 0]) / float(c[0, 1]), "ber_0dB_theory": 0.5 * math.erfc(1.0),
                "ber_6dB": float(c[6, 0]) / float(c[6, 1]), "ber_6dB_theory": 0.5 * math.erfc(math.sqrt(10 ** 0.6)),
                "roofline": {"frac": round(11 * n * (8 + (2 * BPS) / SPS) / ms / 1e6 / peak, 3), "bound": "FP32 lanes + the noise generator (DESIGN.md 4), reported against HBM",
                             "bytes_per_sample": 8 + 2 * BPS / SPS}}

    for w in which:
        if w == "c1":
            rows.append(run_two_kernels("C1 rates", "the reference's default rates (sr 10000 / baud 220 -> sps 45, carrier 1000 Hz, modulate.rs:44-58), 4096 frames x 65520 samples, rect hold + 64-tap low-pass; fused_* = the same loopback as ONE kernel (rx_dec_kernel<..., TXF>)", 4096, 45, False, fused_ok=True))
        if w == "c3":
            rows.append(run_two_kernels("C3", "129-tap RRC both sides, sps 8, 16384 frames x 65536 = 2^30 samples per GPU, exact MACs", 16384, 8, True))
        if w == "c4":
            rows.append(run_sweep(2048))
        if w == "c5":
            rows.append(run_two_kernels("C5 bank", "1024 independent carriers (1000 + 3000c/1024 Hz) over 8 GPUs = 128 carriers x 32 frames per GPU (weak scaling: every rank carries 128 carriers at any N), rect hold + 64-tap low-pass",
                                        4096, 8, False, channels=128, fused_ok=True))
    return rows


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU (debug; the bench config is 4096)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="c1,c3,c4,c5", help="other named shapes to measure in the same run ('' = none)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    binding = bind_near_gpu(local_rank)  # before torch allocates any pinned memory
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    pkg = g.load_package()
    from rust_modem_b200.capi import Comm

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    F = args.frames
    lp = pkg.lowpass_taps()
    m = pkg.Modem(device=local_rank, **path_kwargs(lp))
    stream = torch.cuda.current_stream()
    m.set_stream(stream.cuda_stream)
    K = m.decided_symbols(L)

    comm = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = Comm(m, world, rank, bytes(uid.cpu().numpy().tobytes()))

    gen = torch.Generator(device="cuda").manual_seed(0x5EED0001 + rank)
    d_bits = torch.randint(0, 2, (F, NBITS), dtype=torch.uint8, device="cuda", generator=gen)
    d_tx = torch.empty((F, L, 2), dtype=torch.float32, device="cuda")
    d_sym = torch.empty((F, K), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((F, K * BPS), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")

    def step():
        """one pass of the hot path: the device-resident loopback (for this workload ONE fused kernel that makes the TX
        samples, stores them and demodulates them); the error counters accumulate on the device"""
        m.loopback_device_into(d_bits, F, NBITS, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)

    def reduce_counters():
        if comm is not None:
            pkg.lib().modem_gpu_allreduce_counters(comm._c, d_cnt.data_ptr(), 2)

    def kernels_serial(evs):
        """the two kernels of the unfused path launched back to back over the whole 2 GiB buffer, for comparison"""
        evs[0].record(stream)
        m.modulate_into(d_bits, F, NBITS, tx=d_tx)
        evs[1].record(stream)
        m.demodulate_count_into(d_tx, F, L, d_bits, NBITS, d_cnt, sym=d_sym, bits=d_out)
        evs[2].record(stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    reduce_counters()
    sync_all()
    assert int(d_cnt[0]) == 0 and int(d_cnt[1]) == world * W * F * K * BPS, d_cnt.tolist()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- the timed region: K steps, then the sweep's single all-reduce
    d_cnt.zero_()
    launches0 = m.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0.record(stream)
    for i in range(args.steps):
        step()
    reduce_counters()
    t1.record(stream)
    sync_all()
    launches = m.launch_count - launches0
    total_ms = t0.elapsed_time(t1)
    assert int(d_cnt[0]) == 0 and int(d_cnt[1]) == world * args.steps * F * K * BPS, d_cnt.tolist()
    # ---- the same with an all-reduce after EVERY step (round 1's form: the collective's latency is exposed each step)
    d_cnt.zero_()
    sync_all()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for i in range(args.steps):
        d_cnt.zero_()
        step()
        reduce_counters()
    r1.record(stream)
    sync_all()
    each_ms = r0.elapsed_time(r1)
    # ---- per-kernel times: the fused launch alone, and the two kernels of the unfused path
    fused = launches == args.steps
    lev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3 + args.steps)]
    for e in lev:
        e[0].record(stream)
        m.loopback_device_into(d_bits, F, NBITS, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
        e[1].record(stream)
    sync_all()
    loop_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in lev[3:]]))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3 + args.steps)]
    for e in evs:
        kernels_serial(e)
    sync_all()
    tx_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs[3:]]))
    rx_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs[3:]]))
    clocks = sampler.stop() if rank == 0 else None

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_per_step = max_over_ranks(total_ms) / args.steps
    ms_each = max_over_ranks(each_ms) / args.steps
    value = world * F * L / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: the same loopback through the C ABI with HOST (pinned) buffers
    h_bits = torch.empty((F, NBITS), dtype=torch.uint8).pin_memory()
    h_bits.copy_(d_bits.cpu())
    h_out = torch.empty((F, K * BPS), dtype=torch.uint8).pin_memory()
    h_out.zero_()
    e2e_ms = []
    for i in range(2 + args.e2e_steps):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        err, cmp_ = m.loopback_into(h_bits, F, NBITS, bits_out=h_out)
        b.record(stream)
        torch.cuda.synchronize()
        assert (err, cmp_) == (0, F * K * BPS)
        if i >= 2:
            e2e_ms.append(a.elapsed_time(b))
    assert bool((h_out == h_bits[:, : K * BPS]).all())
    e2e_step = max_over_ranks(float(np.mean(e2e_ms))) if e2e_ms else float("nan")
    e2e_value = world * F * L / (e2e_step * 1e-3) / 1e6
    # ---- the same step on PACKED payloads (extension, modem_gpu_loopback_packed: 8 bits per byte in both directions,
    # unpacked / packed on the device either side of the same kernels) -- reported BESIDE the reference-format e2e
    PB, OB = (NBITS + 7) // 8, (K * BPS + 7) // 8
    h_pk = torch.from_numpy(np.packbits(h_bits.numpy(), axis=1)).pin_memory()
    h_pk_out = torch.zeros((F, OB), dtype=torch.uint8).pin_memory()
    pk_ms = []
    for i in range(2 + args.e2e_steps):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        err, cmp_ = m.loopback_packed_into(h_pk, F, NBITS, h_pk_out)
        b.record(stream)
        torch.cuda.synchronize()
        assert (err, cmp_) == (0, F * K * BPS)
        if i >= 2:
            pk_ms.append(a.elapsed_time(b))
    assert np.array_equal(h_pk_out.numpy(), np.packbits(h_out.numpy(), axis=1)), "packed loopback differs from the byte-per-bit one"
    pk_step = max_over_ranks(float(np.mean(pk_ms))) if pk_ms else float("nan")
    pk_value = world * F * L / (pk_step * 1e-3) / 1e6
    del h_bits, h_out, h_pk, h_pk_out

    peak, peak_src = measured_peaks()
    del d_tx, d_sym, d_out
    torch.cuda.empty_cache()
    which = [w for w in args.configs.split(",") if w] if F == FRAMES else []
    rows = other_configs(pkg, torch, dist, world, rank, local_rank, peak, which) if which else []

    if rank == 0:
        # algorithmic bytes per complex sample (DESIGN.md "roofline"): TX writes 8 B and reads bps/sps B of
        # bits; RX reads 8 B, writes (1 + bps)/sps B of symbols + bits and reads bps/sps B of reference bits.
        tx_bps = 8 + BPS / SPS
        rx_bps = 8 + (1 + BPS) / SPS + BPS / SPS
        kern = {"tx_rect_kernel": (tx_ms, tx_bps), "rx_fast_kernel": (rx_ms, rx_bps)}
        dom = max(kern, key=lambda k: kern[k][0])
        if fused:
            # one kernel: writes the 8 B sample, reads bps/sps B of bits (once: they are also the reference bits of the
            # error count) and writes (1 + bps)/sps B of symbols + bits -- SURVEY.md 8(d)'s fused figure
            kern["loop_fused_kernel"] = (loop_ms, 8 + BPS / SPS + (1 + BPS) / SPS)
            dom = "loop_fused_kernel"
        traffic, traffic_src = committed_traffic()
        ach = {k: F * L * b / (ms * 1e-3) / 1e9 for k, (ms, b) in kern.items()}
        out = {
            "metric": "loopback Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(world, F),  # exactly the keys and values of the --impl reference line
            "host_binding": binding,
            "gpu_launches": int(launches),
            "allreduce": {"per_sweep_ms_per_step": ms_per_step, "every_step_ms_per_step": ms_each,
                          "note": "value uses ONE all-reduce of the counters after the K timed steps (inside the timed region); every_step = an all-reduce after each step"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach[dom], "peak": peak, "unit": "GB/s",
                         "frac": ach[dom] / peak, "traffic": traffic.get(dom), "traffic_source": traffic_src, "peak_source": peak_src,
                         "serial_ms": tx_ms + rx_ms,
                         "kernels": {k: {"ms": kern[k][0], "bytes_per_sample": kern[k][1], "achieved_gbs": ach[k],
                                         "frac": ach[k] / peak} for k in kern}},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(F * NBITS),
                    "d2h_bytes_per_step": int(F * K * BPS + 16), "ms_per_step": e2e_step},
            "e2e_packed": {"value": pk_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(F * PB),
                           "d2h_bytes_per_step": int(F * OB + 16), "ms_per_step": pk_step,
                           "note": "EXTENSION, not the reference's payload format: modem_gpu_loopback_packed, 8 bits per byte in both "
                                   "directions (the reference carries one byte per bit, data.rs:35-40); same kernels, same decisions"},
            "configs": rows,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            frames = 64 * threads
            v, dt = cpu_baseline(frames, threads)
            v1, _ = cpu_baseline(8, 1)
            out["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port",
                                   "sample": f"{frames} of {FRAMES} frames x {L} samples, {dt:.1f} s wall",
                                   "single_thread_value": v1}
        emit(out)
    if comm is not None:
        comm.close()
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
