#!/bin/bash
# e2e (host-buffer) pipeline: PCIe floor of the box, chunk sweep, host-loopback parity tests
mkdir -p gpurun_out
python tools/pcie_floor.py > gpurun_out/e2e2_pcie.txt 2>&1; cat gpurun_out/e2e2_pcie.txt
python -m pytest tests -q -m gpu -k "loopback or host or pipe" > gpurun_out/e2e2_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/e2e2_pytest.log
CHUNKS="${CHUNKS:-128 192 256 384 512}" bash tools/gpu_e2e.sh | tee gpurun_out/e2e2_sweep.txt
