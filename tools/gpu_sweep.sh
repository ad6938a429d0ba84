#!/bin/bash
# tests + RX tuning-variant sweep.  $1 = tag
TAG=${1:-sw}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
for v in ${VARIANTS:-0 1 2 3 4}; do
  MODEM_GPU_RX_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_v$v.json 2> gpurun_out/${TAG}_v$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_v$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]["kernels"]
    print("variant $v: value %.0f Ms/s  ms/step %.3f e2e %.2f ms |"%(d["value"],d["ms_per_step"],d["e2e"]["ms_per_step"]), "  ".join("%s %.3f ms %.3f"%(k.split('_kernel')[0],v["ms"],v["frac"]) for k,v in r.items()))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/${TAG}_v$v.err").read()[-1500:])
PY
done
