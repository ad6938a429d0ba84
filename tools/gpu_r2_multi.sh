#!/bin/bash
# the 8-GPU call (gpurun --gpus 8): topology, PCIe floors at 2/4/8 ranks (FLOORS=1), the bench line with every named shape at
# 8 ranks, BASELINE config 4 (1e11 bits) on 8 ranks
mkdir -p gpurun_out
O=gpurun_out/r2m
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "${FLOORS:-0}" = 1 ]; then
  { nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread"; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; free -g | head -2; } > ${O}_topology.txt 2>&1
  for n in 2 4 8; do $TR --nproc-per-node $n --master-port $((29500+n)) tools/pcie_floor.py 2>/dev/null | grep "^ranks\|^{" ; done | tee ${O}_pcie_floor.txt
  python tools/pcie_floor.py 2>/dev/null | grep "^ranks\|^{" | tee -a ${O}_pcie_floor.txt
fi
$TR --nproc-per-node 8 --master-port 29610 bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_final_bench_8gpu.json 2> ${O}_final_bench_8gpu.err; echo "bench8 rc=$?"
$TR --nproc-per-node 8 --master-port 29613 tools/ber_sweep.py --bits 1e11 > ${O}_ber_sweep_8gpu.json 2> ${O}_ber_sweep_8gpu.err; echo "ber8 rc=$?"
tail -n 2 ${O}_final_bench_8gpu.err; tail -n 2 ${O}_ber_sweep_8gpu.err
