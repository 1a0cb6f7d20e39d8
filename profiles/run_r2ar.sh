#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_policy.py -x -q -s > gpurun_out/r2ar_pytest_policy.log 2>&1; echo "rc=$?"; grep -E "fused vs|passed|failed|Error" gpurun_out/r2ar_pytest_policy.log | tail
timeout 200 python profiles/r2_policy_bench.py 65536 > gpurun_out/r2ar_policy_bench.json 2> gpurun_out/r2ar_policy_bench.err; cat gpurun_out/r2ar_policy_bench.json; tail -3 gpurun_out/r2ar_policy_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 24 -c 4 -o gpurun_out/r2ar_policy -f python profiles/r2_policy_bench.py 65536 > gpurun_out/r2ar_ncu.log 2>&1; tail -2 gpurun_out/r2ar_ncu.log
