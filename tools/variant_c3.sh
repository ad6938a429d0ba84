#!/bin/bash
# timings of tuning builds (tools/bin/variants/libmodem_gpu_<n>.so) against the in-tree library
W="${@:-c3}"
echo "in-tree:"; python tools/bench_configs.py $W | cut -c1-230
for f in tools/bin/variants/libmodem_gpu_*.so; do echo "$f:"; MODEM_GPU_LIB=$PWD/$f python tools/bench_configs.py $W | cut -c1-230; done
