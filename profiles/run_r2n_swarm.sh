#!/bin/bash
# r2n: swarm preset with 8 lanes per env in env_kernel's game-logic pass (EnvCtx GS = 8): parity tests, then bench lines
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_stage03.py -m gpu -x -q -k "swarm" > gpurun_out/r2n_pytest_swarm.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2n_pytest_swarm.log
for k in 1 2; do
  timeout 300 python bench.py --preset swarm --envs 8192 --steps 100 --warmup 5 --no-cpu --no-also --no-e2e --no-rollout --sub-batches $k > gpurun_out/r2n_swarm_k$k.json 2> gpurun_out/r2n_swarm_k$k.err
  python -c "import json;d=json.loads(open('gpurun_out/r2n_swarm_k$k.json').read().strip().splitlines()[-1]);print('K=$k',d['value'],d['ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 24 --csv --log-file gpurun_out/r2n_swarm_launches.csv python bench.py --preset swarm --envs 8192 --steps 8 --warmup 3 --spinup 40 --no-cpu --no-also --no-e2e --no-rollout --sub-batches 1 > gpurun_out/r2n_ncu.log 2>&1
grep -c env_kernel gpurun_out/r2n_swarm_launches.csv
