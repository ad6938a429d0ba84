#!/bin/bash
# one ncu --set full capture of the hot kernels under the bench command.  $1 = tag, env passes through
TAG=${1:-p}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-rx_fast}" -s ${SKIP:-3} -c ${COUNT:-1} -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_ncu_full.log
