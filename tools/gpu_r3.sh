#!/bin/bash
# fused loopback: parity tests, then the bench with and without it
TAG=${1:-r26}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest.log
for nf in 0 1; do
MODEM_GPU_NO_FUSED_LOOP=$nf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_nf$nf.json 2> gpurun_out/${TAG}_bench_nf$nf.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_nf$nf.json").read().strip().splitlines()[-1])
    print("no_fused=$nf value %.0f Ms/s  ms/step %.3f  e2e %.0f Ms/s (%.3f ms)  launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["gpu_launches"]))
    for k,v in d["roofline"]["kernels"].items(): print("  %s: %.3f ms  %.0f GB/s  frac %.3f"%(k,v["ms"],v["achieved_gbs"],v["frac"]))
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/${TAG}_bench_nf$nf.err").read()[-1500:])
PY
done
