#!/bin/bash
# r2ac: the build with the rewritten float32 dynamics / PDL launches / one-pass LiDAR -- whole GPU suite, ncu launch list and
# --set full capture of the step kernels (inputs of profiles/traffic.json), bench lines of every preset
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_pytest_gpu.log
tail -4 gpurun_out/r2ac_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-also --no-e2e --no-rollout > gpurun_out/r2ac_bench_short.json 2> gpurun_out/r2ac_bench.err || exit 1
# ncu only after the plain run exited 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ac_launches.csv python bench.py --steps 20 --warmup 3 --spinup 20 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2ac_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dyn_kernel|env_kernel" -s 640 -c 4 -o gpurun_out/r2ac_full -f python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2ac_ncu2.log 2>&1
for p in "exp02_vFinal 65536" "exp03_vFinal 65536" "exp02_v2_full 8192" "swarm 8192" "level5_c1 65536" "level5_fusion 16384" "level5_dumb_multiobs 8192" "level5_eval_2bt 65536" "stage02 65536" "stage02_10lm 4096" "stage01 65536"; do set -- $p
  timeout 200 python bench.py --preset $1 --envs $2 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2ac_bench_$1_$2.json 2>> gpurun_out/r2ac_bench.err; done
timeout 300 python bench.py --workload lidar --steps 50 > gpurun_out/r2ac_bench_lidar.json 2>> gpurun_out/r2ac_bench.err
for f in gpurun_out/r2ac_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
e=d.get('e2e') or {}
print('$f'.split('r2ac_bench_')[1], '%.4g %s  %.4f ms  frac %.3f' % (d['value'], d['unit'], d['ms_per_step'], (d.get('roofline') or {}).get('frac', 0)))"; done > gpurun_out/r2ac_bench_all.txt 2>&1
cat gpurun_out/r2ac_bench_all.txt
