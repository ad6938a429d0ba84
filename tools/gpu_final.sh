#!/bin/bash
# single-GPU evidence of the round (gpurun, one B200): tests, bench line, reference arm, other configs, front-end rows,
# launch list and ncu --set full of the same bench command (each ncu pass only after the command exited 0 without ncu)
mkdir -p gpurun_out
[ -x tools/bin/check_sqrt ] && [ -x tools/bin/check_sincos_dev ] || bash tools/build_tools.sh > gpurun_out/fin_build_tools.log 2>&1
O=gpurun_out/fin
python -m pytest tests -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 ${O}_pytest.log
python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2>> ${O}_bench.err; echo "ref exit $?"
python tools/bench_configs.py c1 c1l c2 c2l c2f c3 c3f c2n c3n c5 > ${O}_configs.jsonl 2> ${O}_configs.err; echo "configs exit $?"
python tools/bench_frontend.py > ${O}_frontend.jsonl 2> ${O}_frontend.err; echo "frontend exit $?"
python tools/ber_sweep.py --bits 2e10 > ${O}_ber_sweep_1gpu.json 2> ${O}_ber.err; echo "ber sweep exit $?"
python tools/pcie_floor.py > ${O}_pcie_floor_1gpu.txt 2>&1; echo "pcie floor exit $?"
python tools/packed_sweep.py 256 512 1024 2048 4096 > ${O}_packed_sweep.txt 2>&1; echo "packed sweep exit $?"
tools/bin/check_sqrt > ${O}_check_sqrt.txt 2>&1; echo "check_sqrt exit $?"
tools/bin/check_sincos_dev > ${O}_check_sincos_dev.txt 2>&1; echo "check_sincos_dev exit $?"
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline --configs c1"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches.csv $CMD > ${O}_ncu_launch.log 2>&1; echo "launch list exit $?"
# the 12th matching launch is the last fused one of the per-kernel timing loop; the TX / RX pairs of the unfused path follow
ncu --set full --clock-control none --import-source on -k regex:'tx_rect_fast|rx_fast' -s 11 -c 4 -o ${O}_c2_prof $CMD > ${O}_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
C1="python tools/bench_configs.py c1"
$C1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rx_dec' -s 3 -c 1 -o ${O}_c1_prof $C1 > ${O}_ncu_c1.log 2>&1; echo "ncu c1 exit $?"
C1L="python tools/bench_configs.py c1l"
$C1L > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rx_dec' -s 3 -c 1 -o ${O}_c1l_prof $C1L > ${O}_ncu_c1l.log 2>&1; echo "ncu c1l exit $?"
C2N="python tools/bench_configs.py c2n"
$C2N > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rx_fast' -s 3 -c 1 -o ${O}_c2n_prof $C2N > ${O}_ncu_c2n.log 2>&1; echo "ncu c2n exit $?"
