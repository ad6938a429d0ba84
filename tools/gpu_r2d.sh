#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu -k "fused or sign_slicer or pipeline" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2d_pytest.log
python tools/pcie_floor.py 2>&1 | tee gpurun_out/r2d_e2e.txt
python tools/e2e_probe.py two_kernels= fused=MODEM_GPU_PIPE_FUSED=1 \
  fused_c128=MODEM_GPU_PIPE_FUSED=1,MODEM_GPU_PIPE_CHUNK=128 fused_c64=MODEM_GPU_PIPE_FUSED=1,MODEM_GPU_PIPE_CHUNK=64 fused_c512=MODEM_GPU_PIPE_FUSED=1,MODEM_GPU_PIPE_CHUNK=512 \
  zc1=MODEM_GPU_PIPE_ZEROCOPY=1 zc4=MODEM_GPU_PIPE_ZEROCOPY=4 zc16=MODEM_GPU_PIPE_ZEROCOPY=16 2>&1 | tee -a gpurun_out/r2d_e2e.txt
MODEM_GPU_PIPE_TRACE=1 python tools/e2e_probe.py trace_two= 2>&1 | tail -22 | tee -a gpurun_out/r2d_e2e.txt
MODEM_GPU_PIPE_TRACE=1 python tools/e2e_probe.py trace_fused=MODEM_GPU_PIPE_FUSED=1 2>&1 | tail -22 | tee -a gpurun_out/r2d_e2e.txt
