#!/bin/bash
# tools/ab_dynamic_shard.sh N: the 256-file corpus on N GPUs with units claimed on demand (default) and with static shares
N=$1
mkdir -p gpurun_out
for mode in 1 0 1 0; do
  SEGMA_DYNAMIC_SHARD=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550 + mode)) bench.py --gpus $N --steps 2 --warmup 1 --workload corpus 2> gpurun_out/dyn_${mode}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dynamic=$mode', d['n_gpus'], 'GPUs', round(d['value'],3), 'audio-h/s device-timed,', round(d['e2e']['value'],3), 'end to end,', round(d['ms_per_step'],1), 'ms per pass')"
done | tee gpurun_out/r02z_dynamic_shard_ab_${N}gpu.txt
