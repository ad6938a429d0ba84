#!/bin/bash
# run tools/gpu_rxonly.py against each prebuilt library variant in rust-modem_b200/lib/variants/
cp rust-modem_b200/lib/libmodem_gpu.so /tmp/orig.so
for v in rust-modem_b200/lib/variants/*.so; do
  cp $v rust-modem_b200/lib/libmodem_gpu.so
  python tools/gpu_rxonly.py $(basename $v .so) 2>&1 | tail -1
  python tools/gpu_rxonly.py $(basename $v .so) 2>&1 | tail -1
done
cp /tmp/orig.so rust-modem_b200/lib/libmodem_gpu.so
