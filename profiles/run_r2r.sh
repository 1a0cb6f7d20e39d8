#!/bin/bash
# r2r: ncu --set full of the two step kernels on the current build (headline workload)
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu --no-also --no-e2e --no-rollout > gpurun_out/r2r_bench_short.json 2> gpurun_out/r2r_bench.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dyn_kernel|env_kernel" -s 640 -c 4 -o gpurun_out/r2r_full -f python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2r_ncu.log 2>&1
ls -la gpurun_out/r2r_full.ncu-rep
