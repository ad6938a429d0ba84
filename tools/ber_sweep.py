#!/usr/bin/env python
"""BASELINE config 4: AWGN BER-vs-Eb/N0 Monte-Carlo sweep, frames sharded over the GPUs of one box,
one NCCL all-reduce of the [points][2] counters.  QPSK, sps 8, 129-tap RRC both sides, Eb/N0 0..10 dB.

Everything of the sweep runs in the library on the device: the payload bits come from Philox
(modem_gpu_random_bits), each batch is modulated once and demodulated once per Eb/N0 point with Philox AWGN added
while loading (modem_gpu_ber_sweep), the counters stay on the device until the single all-reduce.
Timing is steady state: one warm-up batch first (context set-up, NCO table, NCCL communicator), then CUDA events
around the sweep proper, max over ranks.

  python tools/ber_sweep.py [--bits 1e10] [--frames-per-batch 2048]          (1 GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ber_sweep.py --bits 1e11
"""
import argparse, json, math, os, sys
import numpy as np, torch, torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--bits", type=float, default=1e10)
ap.add_argument("--frames-per-batch", type=int, default=2048)
args = ap.parse_args()
rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = g.load_package()
from rust_modem_b200.capi import Comm

NSYM, BPS = 8192, 2
NBITS = NSYM * BPS
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
L = m.frame_samples(NBITS); K = m.decided_symbols(L)
frames_total = int(math.ceil(args.bits / len(dbs) / (K * BPS)))           # frames per Eb/N0 point, whole job
f0, nf = pkg.shard_range(frames_total, rank, world)                          # this rank's contiguous frame range
FB = args.frames_per_batch
cnt = torch.zeros((len(dbs), 2), dtype=torch.int64, device="cuda")
bits = torch.empty((FB, NBITS), dtype=torch.uint8, device="cuda")
tx = torch.empty((FB, L, 2), dtype=torch.float32, device="cuda")
SEED = 0xA5A5


def sweep():
    done = 0
    while done < nf:
        n = min(FB, nf - done)
        m.random_bits_into(bits, n, NBITS, SEED, frame0=f0 + done)
        m.ber_sweep_into(bits, n, NBITS, sig, cnt, seed=SEED, frame0=f0 + done, tx=tx)
        done += n
    if comm is not None:
        pkg.lib().modem_gpu_allreduce_counters(comm._c, cnt.data_ptr(), cnt.numel())   # the single collective


# warm-up: one batch through every stage, the all-reduce included
m.random_bits_into(bits, min(FB, nf), NBITS, SEED, frame0=f0)
m.ber_sweep_into(bits, min(FB, nf), NBITS, sig, cnt, seed=SEED, frame0=f0, tx=tx)
if comm is not None:
    pkg.lib().modem_gpu_allreduce_counters(comm._c, cnt.data_ptr(), cnt.numel())
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
cnt.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(st)
sweep()
e1.record(st)
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
if rank == 0:
    c = cnt.cpu().numpy()
    rows = [{"ebn0_db": d, "errors": int(e), "bits": int(b), "ber": e / b, "theory": 0.5 * math.erfc(math.sqrt(10 ** (d / 10)))}
            for d, (e, b) in zip(dbs, c)]
    print(json.dumps({"config": "C4 BER sweep (steady state: warm-up batch first, CUDA events, max over ranks; bits from Philox in the library)",
                      "n_gpus": world, "total_bits": int(c[:, 1].sum()), "seconds": dt, "frames_per_batch": FB,
                      "Gbit_per_s": c[:, 1].sum() / dt / 1e9,
                      "Msamples_per_s_per_gpu": len(dbs) * nf * L / dt / 1e6, "points": rows}))
if comm is not None:
    comm.close()
m.close()
if world > 1:
    dist.destroy_process_group()
