import torch, time
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both, chunks=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    c = n // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1):
            d_in[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for both in (False, True):
    for chunks in (1, 16):
        run(both, chunks)
        t = min(run(both, chunks) for _ in range(5))
        print(f"both={both} chunks={chunks}: {t:.3f} ms  -> {n/t/1e6:.1f} GB/s per direction")
