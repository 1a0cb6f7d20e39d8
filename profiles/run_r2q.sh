#!/bin/bash
# r2q: env_kernel P4 as one slot pass; phase clocks; envs-per-warp sweep on the profiling build (DC_EPW / DC_EPB knobs)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage03.py tests/test_gpu_full_size.py tests/test_gpu_sub_batches.py tests/test_gpu_driven.py tests/test_gpu_stage02.py tests/test_gpu_stage01.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest.log
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2q_phase_clocks.txt 2>&1; tail -8 gpurun_out/r2q_phase_clocks.txt
V=dronechase_b200/csrc/libdronechase_b200.so
{ timeout 200 python profiles/r2_variants.py $V exp02_v2_full 65536 1 2 3 4
  for epw in 8 16 32; do for epb in 32 64 128; do
    [ $epb -ge $epw ] && { echo "EPW=$epw EPB=$epb"; DC_EPW=$epw DC_EPB=$epb timeout 200 python profiles/r2_variants.py build/libdc_phases.so exp02_v2_full 65536 2 4; }
  done; done; } > gpurun_out/r2q_variants.txt 2>&1
grep -E "EPW|ms/step|Error" gpurun_out/r2q_variants.txt
