#!/bin/bash
# every preset of BASELINE.md section 4, one JSON line each (device-resident value only; exp02_vFinal and level5_c1 with e2e)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1t_exp02_vFinal.json 2> gpurun_out/bench_r1t.err
python bench.py --preset level5_c1 --no-cpu > gpurun_out/bench_r1t_level5_c1.json 2>> gpurun_out/bench_r1t.err
for p in "exp02_v2_full 65536" "exp03_vFinal 65536" "swarm 8192" "level5_fusion 16384" "stage02 65536" "stage02_10lm 4096" "stage01 65536" "exp02_vFinal 8192"; do set -- $p
  python bench.py --preset $1 --envs $2 --no-e2e --no-cpu > gpurun_out/bench_r1t_$1_$2.json 2>> gpurun_out/bench_r1t.err; done
python bench.py --workload lidar > gpurun_out/bench_r1t_lidar.json 2>> gpurun_out/bench_r1t.err
for f in gpurun_out/bench_r1t_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
e=d.get('e2e') or {}
print('$f'.split('r1t_')[1], '%.4g %s  %.4f ms  frac %.3f  e2e %s' % (d['value'], d['unit'], d['ms_per_step'], (d.get('roofline') or {}).get('frac', 0), e.get('value')))"; done
