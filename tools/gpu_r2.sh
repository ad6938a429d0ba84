#!/bin/bash
# parity (all GPU tests), C3-family configs, e2e pipeline with/without the ramped schedule
mkdir -p gpurun_out
TAG=${1:-r24}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log
python tools/bench_configs.py c3 c3f c3n > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs exit $?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    if l.strip():
        d=json.loads(l); print(d["config"], "tx %.3f rx %.3f loop %.0f errors %s"%(d["tx_ms"],d["rx_ms"],d["loopback_Msamples_s"],d["errors"]))
PY
for nr in 0 1; do
MODEM_GPU_PIPE_NO_RAMP=$nr timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 8 --no-cpu-baseline 2> gpurun_out/e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('no_ramp=$nr: e2e %.0f Ms/s  %.3f ms/step'%(d['e2e']['value'],d['e2e']['ms_per_step']))" || tail -3 gpurun_out/e2e.err
done
