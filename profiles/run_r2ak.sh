#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python profiles/r2_e2e_pieces.py exp02_v2_full 65536 > gpurun_out/r2ak_e2e_pieces.txt 2>&1; cat gpurun_out/r2ak_e2e_pieces.txt
timeout 600 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q 2>&1 | tail -3
