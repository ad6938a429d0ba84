#!/bin/bash
mkdir -p gpurun_out
for tm in 0 1; do for fpb in ${FPBS:-2 4 8 16 32}; do
  MODEM_GPU_RX_TILEMAJOR=$tm MODEM_GPU_RX_FPB=$fpb timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline 2> gpurun_out/fpb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['kernels']['rx_fast_kernel']
print('tile_major=$tm fpb=$fpb: rx %.3f ms frac %.3f   value %.0f'%(r['ms'],r['frac'],d['value']))" || tail -3 gpurun_out/fpb.err
done; done
