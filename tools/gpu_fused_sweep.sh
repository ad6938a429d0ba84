#!/bin/bash
# launch-shape sweep of the fused loopback kernel: frames per CTA, grid order
for fpb in 4 8 12 16 24 32; do
  for tm in 1 0; do
    echo -n "fpb=$fpb tile_major=$tm: "; MODEM_GPU_RX_FPB=$fpb MODEM_GPU_RX_TILEMAJOR=$tm python tools/gpu_fused_probe.py 2>&1 | grep "F=4096"
  done
done
python __graft_entry__.py smoke 2>&1 | tail -2
