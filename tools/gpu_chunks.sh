#!/bin/bash
# sweep the device-loopback chunk size with and without the captured graph
mkdir -p gpurun_out
for ng in 0 1; do for ch in ${CHUNKS:-48 64 96 128 192 256 512}; do
  MODEM_GPU_NO_GRAPH=$ng MODEM_GPU_LOOP_CHUNK=$ch timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2> gpurun_out/chunk.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nograph=$ng chunk=$ch: value %.0f Ms/s  ms/step %.3f (serial kernels %.3f)  launches %d'%(d['value'],d['ms_per_step'],d['roofline']['serial_ms'],d['gpu_launches']))" || tail -3 gpurun_out/chunk.err
done; done
