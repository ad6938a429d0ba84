#!/bin/bash
# round 2, the 8-GPU call: topology, PCIe floors at 2/4/8 ranks, the bench line (with every named shape) at 8 and 2 ranks
mkdir -p gpurun_out
O=gpurun_out/r2m
{ nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread"; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; free -g | head -2; } > ${O}_topology.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do $TR --nproc-per-node $n --master-port $((29500+n)) tools/pcie_floor.py 2>/dev/null | grep "^ranks\|^{" ; done | tee ${O}_pcie_floor.txt
python tools/pcie_floor.py 2>/dev/null | grep "^ranks\|^{" | tee -a ${O}_pcie_floor.txt
$TR --nproc-per-node 8 --master-port 29610 bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_bench_8gpu.json 2> ${O}_bench_8gpu.err; echo "bench8 rc=$?"
$TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --configs c5 > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err; echo "bench2 rc=$?"
$TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 10 --warmup 3 --configs "" > ${O}_bench_4gpu.json 2> ${O}_bench_4gpu.err; echo "bench4 rc=$?"
tail -3 ${O}_bench_8gpu.err
