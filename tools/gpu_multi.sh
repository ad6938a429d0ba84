#!/bin/bash
# multi-GPU bench: $1 = number of GPUs, $2 = tag
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "bench N=$N exit $?"
tail -1 gpurun_out/${TAG}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d value %.0f Ms/s ms/step %.3f e2e %.0f Ms/s launches %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))" || tail -20 gpurun_out/${TAG}_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2>> gpurun_out/${TAG}_n$N.err
echo "reference arm exit $?"; cut -c1-260 gpurun_out/${TAG}_ref_n$N.json
