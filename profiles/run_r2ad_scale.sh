#!/bin/bash
# r2ad: scaling on one 8-GPU box.  weak = 65,536 envs per GPU (default bench line); strong = BASELINE config 3 literally:
# 65,536 envs in total sharded over the GPUs (--total-envs).  One JSON line per run -> gpurun_out/r2ad_{weak,strong}_nN.json
set -x
mkdir -p gpurun_out
run() { # $1 = N, rest = bench args
  local n=$1; shift
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 "$@"
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; fi
}
for n in 1 2 4 8; do
  run $n --steps 100 --warmup 5 --no-cpu --no-also > gpurun_out/r2ad_weak_n$n.json 2> gpurun_out/r2ad_weak_n$n.err
  run $n --steps 100 --warmup 5 --no-cpu --no-also --total-envs 65536 > gpurun_out/r2ad_strong_n$n.json 2> gpurun_out/r2ad_strong_n$n.err
done
for f in gpurun_out/r2ad_*_n*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e, r = d.get("e2e") or {}, d.get("e2e_device_rollout") or {}
    print("%-34s N=%d envs/GPU %6d  value %.4g  %.4f ms  e2e %.4g  rollout %.4g" % (sys.argv[1].split("/")[-1], d["n_gpus"], d["config"]["envs_per_gpu"], d["value"], d["ms_per_step"], e.get("value", 0), r.get("value", 0)))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done | tee gpurun_out/r2ad_scale_table.txt
