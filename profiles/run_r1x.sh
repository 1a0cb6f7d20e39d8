#!/bin/bash
# r1x: final round-1 build -- GPU tests, bench lines of every preset, ncu launch list + --set full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1x_pytest_gpu.log
tail -5 gpurun_out/r1x_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r1x_bench_exp02_vFinal.json 2> gpurun_out/r1x_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1x_bench_reference.json 2>> gpurun_out/r1x_bench.err
timeout 200 python bench.py --preset level5_c1 --no-cpu > gpurun_out/r1x_bench_level5_c1.json 2>> gpurun_out/r1x_bench.err
for p in "exp02_v2_full 65536" "exp03_vFinal 65536" "swarm 8192" "level5_fusion 16384" "level5_dumb_multiobs 8192" "level5_eval_2bt 65536" "stage02 65536" "stage01 65536"; do set -- $p
  timeout 200 python bench.py --preset $1 --envs $2 --no-e2e --no-cpu > gpurun_out/r1x_bench_$1_$2.json 2>> gpurun_out/r1x_bench.err; done
timeout 200 python bench.py --preset level5_fusion --envs 16384 --student --no-e2e --no-cpu > gpurun_out/r1x_bench_level5_fusion_student_16384.json 2>> gpurun_out/r1x_bench.err
for f in gpurun_out/r1x_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
e=d.get('e2e') or {}
print('$f'.split('r1x_bench_')[1], '%.4g %s  %.4f ms  frac %.3f  e2e %s' % (d['value'], d['unit'], d['ms_per_step'], (d.get('roofline') or {}).get('frac', 0), e.get('value')))"; done > gpurun_out/r1x_bench_all.txt 2>&1
cat gpurun_out/r1x_bench_all.txt
# ncu only after the plain runs exited 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1x_launches.csv python bench.py --steps 20 --warmup 3 --spinup 20 --no-e2e --no-cpu > gpurun_out/r1x_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dyn_kernel|env_kernel" -s 640 -c 4 -o gpurun_out/r1x_full -f python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1x_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"stack_kernel" -s 60 -c 2 -o gpurun_out/r1x_stack_multi -f python bench.py --preset level5_dumb_multiobs --envs 8192 --steps 20 --warmup 3 --spinup 60 --no-e2e --no-cpu > gpurun_out/r1x_ncu3.log 2>&1
ls -la gpurun_out | tail -30
