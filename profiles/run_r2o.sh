#!/bin/bash
# r2o: float32 dynamics rewritten for instruction count (quad_substep_f32, folded cf2x model): whole GPU suite, bench line,
# per-launch instruction counts
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2o_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-also --no-e2e --no-rollout > gpurun_out/r2o_bench_short.json 2> gpurun_out/r2o_bench.err; tail -c 600 gpurun_out/r2o_bench_short.json
for k in 1 2 3 4; do timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 65536 $k; done 2>&1 | grep ms/step | tee gpurun_out/r2o_variants.txt
timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 8192 1 2 2>&1 | grep ms/step | tee -a gpurun_out/r2o_variants.txt
timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so swarm 8192 1 2 2>&1 | grep ms/step | tee -a gpurun_out/r2o_variants.txt
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 24 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-also --no-e2e --no-rollout > gpurun_out/r2o_ncu.log 2>&1
