#!/bin/bash
set -x
mkdir -p gpurun_out
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2x_phase_clocks.txt 2>&1; tail -9 gpurun_out/r2x_phase_clocks.txt
timeout 600 python -m pytest tests/test_gpu_stage03.py -m gpu -x -q -k "custom_model or teacher" 2>&1 | tail -5
