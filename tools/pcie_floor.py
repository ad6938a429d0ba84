"""PCIe floor of the host-buffer loopback step: the step's two transfers alone (64 MiB of bits in, 64 MiB of results out,
pinned memory, copy engines only), per rank and concurrently on every rank when launched under torchrun.
  python tools/pcie_floor.py                       one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_floor.py
Prints, per pattern, the max over ranks of the CUDA-event time and the per-direction GB/s per GPU."""
import os, sys, json
import torch
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both, chunks=1, stagger=0):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    c = n // chunks
    if stagger:
        with torch.cuda.stream(s2):
            torch.cuda._sleep(stagger)  # offsets the copy-out stream by `stagger` GPU cycles
    for i in range(chunks):
        with torch.cuda.stream(s1):
            d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
res = {}
for both in (False, True):
    for chunks, stagger in ((1, 0), (16, 0)) + (((16, 90000),) if both else ()):
        run(both, chunks, stagger)
        t = min(run(both, chunks, stagger) for _ in range(5))
        key = f"{'in+out' if both else 'in only'}, {chunks} transfer(s) per direction" + (f", copy-out offset by {stagger} cycles" if stagger else "")
        res[key] = round(t, 4)
        if rank == 0:
            print(f"ranks={world} {key}: {t:.3f} ms (max over ranks) -> {n/t/1e6:.1f} GB/s per direction per GPU", flush=True)
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_per_direction": n, "ms": res}), flush=True)
if world > 1:
    dist.destroy_process_group()
