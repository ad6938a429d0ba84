#!/bin/bash
# final single-GPU evidence of the round: tests, bench line, other configs, front-end rows, launch list
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/fin_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/fin_pytest.log
python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/fin_bench_ref.json 2>> gpurun_out/fin_bench.err; echo "ref exit $?"
python tools/bench_configs.py c1 c2 c2f c3 c3f c2n c3n c5 > gpurun_out/fin_configs.jsonl 2> gpurun_out/fin_configs.err; echo "configs exit $?"
python tools/bench_frontend.py > gpurun_out/fin_frontend.jsonl 2> gpurun_out/fin_frontend.err; echo "frontend exit $?"
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/fin_launches.csv $CMD > gpurun_out/fin_ncu_launch.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'tx_rect_fast|rx_fast' -s 4 -c 3 -o gpurun_out/fin_c2_prof $CMD > gpurun_out/fin_ncu_c2.log 2>&1; echo "ncu exit $?"
