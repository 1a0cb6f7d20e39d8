#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_policy.py -x -q > gpurun_out/r2an_pytest_policy.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2an_pytest_policy.log
timeout 200 python profiles/r2_policy_bench.py 65536 > gpurun_out/r2an_policy_bench.json 2> gpurun_out/r2an_policy_bench.err; cat gpurun_out/r2an_policy_bench.json; tail -3 gpurun_out/r2an_policy_bench.err
timeout 200 python -m pytest tests/test_gpu_io_data.py -x -q 2>&1 | tail -3
