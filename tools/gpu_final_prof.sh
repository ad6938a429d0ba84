#!/bin/bash
# final-state evidence for profiles/: launch list + --set full of both hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/final_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/final_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'tx_rect_fast|rx_fast' -s 6 -c 2 -o gpurun_out/final_prof $CMD > gpurun_out/final_ncu_full.log 2>&1
echo "full exit $?"
