#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapters.py tests/test_gpu_level5.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python profiles/r2_e2e_ab.py exp02_v2_full 65536 > gpurun_out/r2al_e2e_ab.txt 2>&1; cat gpurun_out/r2al_e2e_ab.txt
