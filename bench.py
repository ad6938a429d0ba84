#!/usr/bin/env python
"""bench.py -- loopback throughput of the batched modulate -> demodulate path on B200.

Workload (BASELINE.json configs[1], "C2"): 4096 frames x 65536 complex f32 samples per GPU,
QPSK, 8 samples/symbol, rectangular hold (the reference's TX pulse) + the reference's 64-tap
low-pass (src/bin/demodulate.rs:82-147), no channel noise.  One step = modulate every frame
into the TX buffer in HBM, demodulate it back (decimate, slice, demap) and count bit errors
against the input bits; with N > 1 ranks each rank owns its own 4096 frames ("weak" scaling,
frames are independent) and one tiny NCCL all-reduce sums the error counters.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                          the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).
"""
import argparse
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


def path_kwargs(lowpass):
    kw = dict(PATH)
    kw.update(rx_taps=lowpass, decision_delay=31 + SPS // 2, slicer_gain=float(np.float32(lowpass.sum())))
    return kw


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    Rust crate cannot be built here -- no rustc), all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = 8 * threads
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_baseline(frames, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals])) * 1e3
    sample = f"{frames} of {FRAMES} frames x {L} samples per step, {threads} threads (frames sharded)"
    emit({
        "impl": "reference", "metric": "loopback Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

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
        """one pass of the hot path: the device-resident loopback (for this workload ONE fused kernel that makes the
        TX samples, stores them and demodulates them; two kernels when MODEM_GPU_NO_FUSED_LOOP=1) + the tiny
        all-reduce of the counters"""
        d_cnt.zero_()
        m.loopback_device_into(d_bits, F, NBITS, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
        if comm is not None:
            pkg.lib().modem_gpu_allreduce_counters(comm._c, d_cnt.data_ptr(), 2)

    def kernels_serial(evs):
        """the two hot kernels launched back to back over the whole 2 GiB buffer (no chunk pipeline):
        the per-kernel times the roofline is computed from"""
        d_cnt.zero_()
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

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    assert int(d_cnt[0]) == 0 and int(d_cnt[1]) == world * F * K * BPS, d_cnt.tolist()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = m.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0.record(stream)
    for i in range(args.steps):
        step()
    t1.record(stream)
    sync_all()
    launches = m.launch_count - launches0
    total_ms = t0.elapsed_time(t1)
    assert int(d_cnt[0]) == 0 and int(d_cnt[1]) == world * F * K * BPS, d_cnt.tolist()
    # per-kernel times for the roofline: whole-buffer launches, serial
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(3 + args.steps)]
    for e in evs:
        kernels_serial(e)
    sync_all()
    tx_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs[3:]]))
    rx_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs[3:]]))
    # the loopback entry alone (no counter reset, no all-reduce): the fused kernel's launch duration when it is one launch
    fused = launches == args.steps
    lev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(3 + args.steps)]
    for e in lev:
        e[0].record(stream)
        m.loopback_device_into(d_bits, F, NBITS, d_cnt, tx=d_tx, sym=d_sym, bits_out=d_out)
        e[1].record(stream)
    sync_all()
    loop_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in lev[3:]]))
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    value = world * F * L / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: the same loopback through the C ABI with HOST (pinned) buffers
    h_bits = torch.empty((F, NBITS), dtype=torch.uint8).pin_memory()
    h_bits.copy_(d_bits.cpu())
    h_out = torch.empty((F, K * BPS), dtype=torch.uint8).pin_memory()
    e2e_ms = []
    for i in range(1 + args.e2e_steps):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        err, cmp_ = m.loopback_into(h_bits, F, NBITS, bits_out=h_out)
        b.record(stream)
        torch.cuda.synchronize()
        assert (err, cmp_) == (0, F * K * BPS)
        if i:
            e2e_ms.append(a.elapsed_time(b))
    assert bool((h_out == h_bits[:, : K * BPS]).all())
    et = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = world * F * L / (float(et.item()) * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peaks()
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
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram__bytes_read+write per launch, from the committed ncu capture
        if os.path.exists(tp):
            traffic = json.load(open(tp))
        ach = {k: F * L * b / (ms * 1e-3) / 1e9 for k, (ms, b) in kern.items()}
        out = {
            "metric": "loopback Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": F, "samples_per_frame": L,
                       "l2": ("per step 2 GiB of TX samples are written per GPU (16x the 126 MB L2), no flush between steps: "
                              "every step's buffers are far larger than L2; a step is ONE launch of the fused loopback kernel "
                              "(rx_fast_kernel<..., TXF>: makes the TX samples from the bits, stores them, demodulates them from "
                              "registers; DESIGN.md 4); roofline.kernel is that launch timed alone; roofline.kernels also lists the "
                              "two kernels of the unfused path (modulate, then demodulate from memory) for comparison") if fused else
                             ("per step 2 GiB of TX samples are written and read per GPU (16x the 126 MB L2), no flush "
                              "between steps: every step's inputs are far larger than L2; a step is one whole-buffer TX launch and one "
                              "whole-buffer RX launch; roofline.kernels are the same two launches timed separately"),
                       "parallelism": f"frames sharded over {world} GPU(s), one NCCL all-reduce of 2 u64 counters"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach[dom], "peak": peak, "unit": "GB/s",
                         "frac": ach[dom] / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                         "serial_ms": tx_ms + rx_ms,
                         "kernels": {k: {"ms": kern[k][0], "bytes_per_sample": kern[k][1], "achieved_gbs": ach[k],
                                         "frac": ach[k] / peak} for k in kern}},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(F * NBITS),
                    "d2h_bytes_per_step": int(F * K * BPS + 16), "ms_per_step": float(et.item())},
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
