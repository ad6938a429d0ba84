#!/bin/bash
# iteration pass: parity tests, bench, optional ncu.  $1 = tag, $2 = "ncu" to profile
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]["kernels"]
    print("value %.0f Ms/s  ms/step %.3f  e2e %.0f Ms/s (%.2f ms)  launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"],d["gpu_launches"]))
    for k,v in r.items(): print("  %s: %.3f ms  %.0f GB/s  frac %.3f"%(k,v["ms"],v["achieved_gbs"],v["frac"]))
    print("  clocks",d["clocks"],"cpu",d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
if [ "$2" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
  ncu --set full --clock-control none --import-source on -k regex:'tx_rect|rx_fast|tx_shaped' -s 6 -c 2 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu exit $?"
fi
