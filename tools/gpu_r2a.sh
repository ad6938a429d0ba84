#!/bin/bash
# round 2, call A: parity of the reworked fast RX / fused kernels, then the fused tuning variants
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2a_pytest.log
python tools/time_fused.py 0 1 2 3 4 5 6 7 8 2>&1 | tee gpurun_out/r2a_variants.txt
for fpb in 8 32; do MODEM_GPU_RX_FPB=$fpb python tools/time_fused.py 0 1 3 2>&1 | tee -a gpurun_out/r2a_variants.txt; done
python tools/bench_configs.py c2 c3 c2n c5 > gpurun_out/r2a_configs.jsonl 2> gpurun_out/r2a_configs.err; cat gpurun_out/r2a_configs.jsonl
