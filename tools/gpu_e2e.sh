#!/bin/bash
mkdir -p gpurun_out
for ch in ${CHUNKS:-64 128 256 384 512 768 1024}; do
  MODEM_GPU_PIPE_CHUNK=$ch timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline 2> gpurun_out/e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pipe chunk=$ch: e2e %.0f Ms/s  %.3f ms/step'%(d['e2e']['value'],d['e2e']['ms_per_step']))" || tail -3 gpurun_out/e2e.err
done
