#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n exit $?"
  tail -1 gpurun_out/scale_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  N=%d value %.0f Ms/s  ms/step %.3f  e2e %.0f Ms/s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value']))" || tail -5 gpurun_out/scale_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 tools/ber_sweep.py --bits 1e11 > gpurun_out/r01_ber_n$N.json 2> gpurun_out/ber_n$N.err; echo "ber sweep exit $?"; cut -c1-330 gpurun_out/r01_ber_n$N.json
