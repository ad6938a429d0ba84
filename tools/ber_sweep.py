#!/usr/bin/env python
"""BASELINE config 4: AWGN BER-vs-Eb/N0 Monte-Carlo sweep, frames sharded over the GPUs of one box,
one NCCL all-reduce of the [points][2] counters.  QPSK, sps 8, 129-tap RRC both sides, Eb/N0 0..10 dB.

  python tools/ber_sweep.py [--bits 1e10] [--frames-per-batch 4096]          (1 GPU)
  torchrun --nproc-per-node N tools/ber_sweep.py --bits 1e11                 (N GPUs)
"""
import argparse, json, math, os, sys, time
import numpy as np, torch, torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--bits", type=float, default=1e10)
ap.add_argument("--frames-per-batch", type=int, default=4096)
args = ap.parse_args()
rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = g.load_package()
from rust_modem_b200.capi import Comm

NSYM, BPS = 8192, 2
rrc = pkg.rrc_taps(16, 8, 0.35)
m = pkg.Modem("qpsk", 1250, 10000, 2500, tx_taps=rrc, rx_taps=rrc, decision_delay=128, slicer_gain=1.0, device=lr)
st = torch.cuda.current_stream(); m.set_stream(st.cuda_stream)
comm = None
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0: uid.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = Comm(m, world, rank, bytes(uid.cpu().numpy().tobytes()))
dbs = list(range(0, 11)); sig = [m.sigma_for_ebn0(float(d)) for d in dbs]
L = m.frame_samples(NSYM * BPS); K = m.decided_symbols(L)
frames_total = int(math.ceil(args.bits / len(dbs) / (K * BPS)))           # frames per Eb/N0 point, whole job
f0, nf = pkg.shard_range(frames_total, rank, world)                          # this rank's contiguous frame range
FB = args.frames_per_batch
cnt = torch.zeros((len(dbs), 2), dtype=torch.int64, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(0xA5A5 + rank)
tx = torch.empty((FB, L, 2), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
done = 0
while done < nf:
    n = min(FB, nf - done)
    bits = torch.randint(0, 2, (n, NSYM * BPS), dtype=torch.uint8, device="cuda", generator=gen)
    m.ber_sweep_into(bits, n, NSYM * BPS, sig, cnt, seed=0xA5A5, frame0=f0 + done, tx=tx)
    done += n
if comm is not None:
    pkg.lib().modem_gpu_allreduce_counters(comm._c, cnt.data_ptr(), cnt.numel())   # the single collective
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    c = cnt.cpu().numpy()
    rows = [{"ebn0_db": d, "errors": int(e), "bits": int(b), "ber": e / b, "theory": 0.5 * math.erfc(math.sqrt(10 ** (d / 10)))}
            for d, (e, b) in zip(dbs, c)]
    print(json.dumps({"config": "C4 BER sweep", "n_gpus": world, "total_bits": int(c[:, 1].sum()), "seconds": dt,
                      "Gbit_per_s": c[:, 1].sum() / dt / 1e9, "points": rows}))
if world > 1:
    dist.destroy_process_group()
