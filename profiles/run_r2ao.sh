#!/bin/bash
set -x
mkdir -p gpurun_out
./profiles/micro/mma_tf32_rate > gpurun_out/r2ao_mma_rate.txt 2>&1; cat gpurun_out/r2ao_mma_rate.txt
timeout 300 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 12 -c 2 -o gpurun_out/r2ao_policy -f python profiles/r2_policy_bench.py 65536 > gpurun_out/r2ao_ncu.log 2>&1; tail -3 gpurun_out/r2ao_ncu.log
ls -la gpurun_out/
