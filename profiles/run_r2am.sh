#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2am_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2am_pytest_gpu.log
timeout 600 python profiles/r2_e2e_ab.py exp02_v2_full 65536 > gpurun_out/r2am_e2e_ab.txt 2>&1; cat gpurun_out/r2am_e2e_ab.txt
