#!/bin/bash
# ncu --set full of the C3 kernels (129-tap RRC both sides) under tools/bench_configs.py c3.  $1 = tag
TAG=${1:-c3}
mkdir -p gpurun_out
CMD="python tools/bench_configs.py ${CFG:-c3}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-tx_shaped|rx_fast}" -s ${SKIP:-4} -c ${COUNT:-2} -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_ncu_full.log
