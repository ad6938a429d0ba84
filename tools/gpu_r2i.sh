#!/bin/bash
mkdir -p gpurun_out
for f in 64 128 256 512 1024; do TF_FRAMES=$f python tools/time_fused.py 0; done 2>&1 | tee gpurun_out/r2i_e2e.txt
python tools/pcie_floor.py 2>&1 | tee -a gpurun_out/r2i_e2e.txt
python tools/e2e_probe.py fused_fpb2=MODEM_GPU_RX_FPB=2,MODEM_GPU_PIPE_RAMP=0 fused_fpb4=MODEM_GPU_RX_FPB=4,MODEM_GPU_PIPE_RAMP=0 fused_fpb16=MODEM_GPU_RX_FPB=16,MODEM_GPU_PIPE_RAMP=0 2>&1 | tee -a gpurun_out/r2i_e2e.txt
