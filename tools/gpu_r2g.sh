#!/bin/bash
mkdir -p gpurun_out
python tools/time_fused.py 0 2>&1 | tee gpurun_out/r2g_e2e.txt
(nvidia-smi --query-gpu=clocks.sm,clocks.mem,pstate,pcie.link.gen.current,pcie.link.width.current,power.draw --format=csv,noheader -lms 50 > gpurun_out/r2g_smi.txt &) 
sleep 0.5
E2E_REPS=1 python tools/e2e_probe.py fused= fused_store=MODEM_GPU_PIPE_STORE_TX=1 two_kernels=MODEM_GPU_PIPE_TWO_KERNELS=1,MODEM_GPU_PIPE_RAMP=0 fused_c512_store=MODEM_GPU_PIPE_CHUNK=512,MODEM_GPU_PIPE_STORE_TX=1 2>&1 | tee -a gpurun_out/r2g_e2e.txt
python tools/pcie_floor.py 2>&1 | tee -a gpurun_out/r2g_e2e.txt
pkill -x nvidia-smi
sort gpurun_out/r2g_smi.txt | uniq -c | sort -rn | head -12
