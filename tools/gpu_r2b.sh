#!/bin/bash
# round 2, call B: pipe rates, CTA-shape variants of the fused and the plain fast RX kernel, ncu of the best fused variant
mkdir -p gpurun_out
tools/bin/pipe_rate 2>&1 | tee gpurun_out/r2b_pipe_rate.txt
for fpb in 8 16 32; do MODEM_GPU_RX_FPB=$fpb python tools/time_fused.py 4 2>&1 | tee -a gpurun_out/r2b_variants.txt; done
MODEM_GPU_RX_TILEMAJOR=0 python tools/time_fused.py 4 0 2>&1 | tee -a gpurun_out/r2b_variants.txt
for v in 0 1 2; do MODEM_GPU_RX64_VARIANT=$v python tools/bench_configs.py c2 c2n c5 2>/dev/null | cut -c1-200 | sed "s/^/rx64 variant $v: /" | tee -a gpurun_out/r2b_variants.txt; done
export MODEM_GPU_FUSED_VARIANT=4
CMD="python tools/time_fused.py --one"
ncu --set full --clock-control none --import-source on -k regex:'rx_fast' -s 3 -c 1 -o gpurun_out/r2b_fused_v4 $CMD > gpurun_out/r2b_ncu.log 2>&1; echo "ncu exit $?"
