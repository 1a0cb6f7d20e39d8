#!/bin/bash
# r2y: float32 LiDAR projection in the float32 build of env_kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage03.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py tests/test_gpu_stage02.py tests/test_gpu_stage01.py tests/test_gpu_vec_env.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2y_pytest.log
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2y_phase_clocks.txt 2>&1; tail -9 gpurun_out/r2y_phase_clocks.txt
V=dronechase_b200/csrc/libdronechase_b200.so
{ timeout 200 python profiles/r2_variants.py $V exp02_v2_full 65536 1 2
  timeout 200 python profiles/r2_variants.py $V exp02_v2_full 8192 1
  timeout 200 python profiles/r2_variants.py $V swarm 8192 1
  timeout 200 python profiles/r2_variants.py $V exp03_vFinal 65536 2; } > gpurun_out/r2y_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2y_variants.txt
