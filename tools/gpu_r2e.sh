#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu -k "fused or sign_slicer or pipeline or full_size" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2e_pytest.log
python tools/time_fused.py 0 2>&1 | tee gpurun_out/r2e_e2e.txt
python tools/e2e_probe.py two_kernels= zc1=MODEM_GPU_PIPE_ZEROCOPY=1 zc2=MODEM_GPU_PIPE_ZEROCOPY=2 zc8=MODEM_GPU_PIPE_ZEROCOPY=8 \
   zc1_fpb4=MODEM_GPU_PIPE_ZEROCOPY=1,MODEM_GPU_RX_FPB=4 zc1_fpb32=MODEM_GPU_PIPE_ZEROCOPY=1,MODEM_GPU_RX_FPB=32 2>&1 | tee -a gpurun_out/r2e_e2e.txt
