#!/bin/bash
# r2aa: work-list atomic issued before P5's stores, P3 inputs prefetched during P0, float32 LiDAR projection: whole suite + timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2aa_pytest_gpu.log
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2aa_phase_clocks.txt 2>&1; tail -9 gpurun_out/r2aa_phase_clocks.txt
V=dronechase_b200/csrc/libdronechase_b200.so
{ timeout 200 python profiles/r2_variants.py $V exp02_v2_full 65536 1 2
  timeout 200 python profiles/r2_variants.py $V exp02_v2_full 8192 1
  timeout 200 python profiles/r2_variants.py $V swarm 8192 1
  timeout 200 python profiles/r2_variants.py $V level5_c1 16384 2
  timeout 200 python profiles/r2_variants.py $V exp03_vFinal 65536 2; } > gpurun_out/r2aa_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2aa_variants.txt
