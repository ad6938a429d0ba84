#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu -k "pipeline or fused or bank" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2f_pytest.log
python tools/time_fused.py 0 2>&1 | tee gpurun_out/r2f_e2e.txt
python tools/e2e_probe.py fused_ramp= fused_noramp=MODEM_GPU_PIPE_RAMP=0 two_kernels=MODEM_GPU_PIPE_TWO_KERNELS=1 two_noramp=MODEM_GPU_PIPE_TWO_KERNELS=1,MODEM_GPU_PIPE_RAMP=0 \
   fused_c128=MODEM_GPU_PIPE_CHUNK=128 fused_c512=MODEM_GPU_PIPE_CHUNK=512 fused_c1024=MODEM_GPU_PIPE_CHUNK=1024 fused_c512_noramp=MODEM_GPU_PIPE_CHUNK=512,MODEM_GPU_PIPE_RAMP=0 2>&1 | tee -a gpurun_out/r2f_e2e.txt
MODEM_GPU_PIPE_TRACE=1 python tools/e2e_probe.py trace_fused= 2>&1 | tail -26 | tee -a gpurun_out/r2f_e2e.txt
