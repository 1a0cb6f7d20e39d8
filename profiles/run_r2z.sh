#!/bin/bash
# r2z: what happens to the phases of env_kernel's warp chain with 16 / 8 envs per warp (profiling build knobs)
set -x
mkdir -p gpurun_out
for epw in 32 16 8; do echo "== DC_EPW=$epw"; DC_EPW=$epw DC_EPB=$((epw*4)) DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full 2>&1 | tail -8; done > gpurun_out/r2z_phase_clocks_epw.txt
cat gpurun_out/r2z_phase_clocks_epw.txt
timeout 900 python -m pytest tests/test_gpu_stage03.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py tests/test_gpu_stage02.py tests/test_gpu_stage01.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2y_pytest.log
