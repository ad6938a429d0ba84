#!/bin/bash
mkdir -p gpurun_out
MODEM_GPU_PIPE_TRACE=1 python tools/e2e_probe.py trace_fused=MODEM_GPU_PIPE_RAMP=0 2>&1 | tail -18 | tee gpurun_out/r2h_e2e.txt
MODEM_GPU_PIPE_TRACE=1 python tools/e2e_probe.py trace_two=MODEM_GPU_PIPE_TWO_KERNELS=1,MODEM_GPU_PIPE_RAMP=0 2>&1 | tail -18 | tee -a gpurun_out/r2h_e2e.txt
