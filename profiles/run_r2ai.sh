#!/bin/bash
# r2ai: end-to-end arm at 8 and 4 ranks of one box, change list against mapped mirror (DRONECHASE_B200_LIDAR_TRANSFER)
set -x
mkdir -p gpurun_out
for n in 8 4; do for m in pairs mapped; do
  DRONECHASE_B200_LIDAR_TRANSFER=$m timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 60 --warmup 5 --no-cpu --no-also --no-rollout > gpurun_out/r2ai_${m}_n$n.json 2> gpurun_out/r2ai_${m}_n$n.err
  python -c "import json;d=json.loads(open('gpurun_out/r2ai_${m}_n$n.json').read().strip().splitlines()[-1]);print('$m N=$n value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
done; done | tee gpurun_out/r2ai_table.txt
