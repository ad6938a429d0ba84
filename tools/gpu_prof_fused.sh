#!/bin/bash
# launch list + ncu --set full of the bench command with the fused loopback kernel.  $1 = tag
TAG=${1:-fz}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'rx_fast|tx_rect' -s 4 -c 3 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu exit $?"
