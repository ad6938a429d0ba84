#!/bin/bash
TAG=${1:-r27}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
python tools/gpu_fused_probe.py
python tools/gpu_e2e_probe.py
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.0f Ms/s  ms/step %.3f  e2e %.0f Ms/s (%.3f ms)  launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["gpu_launches"]))
    print("  dominant", d["roofline"]["kernel"], d["roofline"]["frac"])
    for k,v in d["roofline"]["kernels"].items(): print("  %s: %.3f ms  %.0f GB/s  frac %.3f"%(k,v["ms"],v["achieved_gbs"],v["frac"]))
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
