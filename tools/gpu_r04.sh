#!/bin/bash
# C3 (129-tap) variant sweep + ncu captures of the C3 kernels and the current C2 kernels
mkdir -p gpurun_out
for v in 0 21 22 23 24 25 26 27; do
  MODEM_GPU_RX_VARIANT=$v python tools/bench_configs.py c3 > gpurun_out/r04_c3_v$v.json 2> gpurun_out/r04_c3_v$v.err
  echo "v$v: $(cut -c1-200 gpurun_out/r04_c3_v$v.json)"
done
CMD="python tools/bench_configs.py c3"
ncu --set full --clock-control none --import-source on -k regex:'tx_shaped_fast|rx_fast' -s 4 -c 2 -o gpurun_out/r04_c3_prof $CMD > gpurun_out/r04_ncu_c3.log 2>&1
echo "ncu c3 exit $?"
CMD2="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD2 > gpurun_out/r04_plain.log 2>&1 || { echo "plain failed"; tail gpurun_out/r04_plain.log; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r04_launches.csv $CMD2 > gpurun_out/r04_ncu_launch.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'tx_rect_fast|rx_fast' -s 6 -c 2 -o gpurun_out/r04_c2_prof $CMD2 > gpurun_out/r04_ncu_c2.log 2>&1
echo "ncu c2 exit $?"
ls -la gpurun_out/*.ncu-rep
