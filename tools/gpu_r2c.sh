#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu -k "fused or loopback or rx_ or full_size or noisy or bank" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2c_pytest.log
python tools/time_fused.py 0 1 3 4 5 6 9 10 11 2>&1 | tee gpurun_out/r2c_variants.txt
for v in 4 9 10; do MODEM_GPU_FUSED_VARIANT=$v python -m pytest tests -q -x -m gpu -k "fused" 2>&1 | tail -1 | sed "s/^/variant $v pytest: /" | tee -a gpurun_out/r2c_variants.txt; done
python tools/bench_configs.py c2 c3 > gpurun_out/r2c_configs.jsonl 2>/dev/null; cut -c1-260 gpurun_out/r2c_configs.jsonl
