#!/bin/bash
# ncu passes on the bench command (1 GPU).  $1 = tag
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'tx_rect|rx_fast' -s 6 -c 4 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full exit $?"; tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/
