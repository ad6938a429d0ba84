#!/bin/bash
# builds the stand-alone measurement / checking tools under tools/bin/ (git-ignored; they travel with gpurun snapshots)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
A="-gencode arch=compute_100a,code=sm_100a"
nvcc $A -O2 -fmad=false -o tools/bin/check_sqrt tools/check_sqrt.cu
nvcc $A -O2 -fmad=false -Xcompiler -fopenmp -I rust-modem_b200/csrc -o tools/bin/check_sincos_dev tools/check_sincos_dev.cu
nvcc $A -O3 -std=c++17 -fmad=false -o tools/bin/pipe_rate tools/pipe_rate.cu
nvcc $A -O2 -o tools/bin/wc_probe tools/wc_probe.cu
[ -f tools/fp_rate.cu ] && nvcc $A -O3 -o tools/bin/fp_rate tools/fp_rate.cu
echo "tools built"
