#!/bin/bash
# r2at: the final build at 8 GPUs of one box, weak (65,536 envs per GPU) and strong (65,536 in total): device step, numpy VecEnv end to end
# (change-list transfer, terminal observations on), device rollout with the fused policy
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu --no-also > gpurun_out/r2at_weak_n8.json 2> gpurun_out/r2at_weak_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu --no-also --total-envs 65536 > gpurun_out/r2at_strong_n8.json 2> gpurun_out/r2at_strong_n8.err
for f in gpurun_out/r2at_*_n8.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e, r = d.get("e2e") or {}, d.get("e2e_device_rollout") or {}
    print("%-28s N=%d envs/GPU %6d  value %.4g  %.4f ms  e2e %.4g  rollout %.4g (tf32 %.4g, torch %.4g)" % (sys.argv[1].split("/")[-1], d["n_gpus"], d["config"]["envs_per_gpu"], d["value"], d["ms_per_step"], e.get("value", 0), r.get("value", 0), r["policy_tf32"]["value"], r["policy_torch_module"]["value"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done | tee gpurun_out/r2at_scale_table.txt
nproc; tail -2 gpurun_out/r2at_weak_n8.err
