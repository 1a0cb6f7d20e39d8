#!/bin/bash
# r2v: nearest invader once per pursuer and step, state prefetch in dyn_kernel: whole GPU suite, timings, phase clocks
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2v_pytest_gpu.log
V=dronechase_b200/csrc/libdronechase_b200.so
{ timeout 200 python profiles/r2_variants.py $V exp02_v2_full 65536 1 2
  timeout 200 python profiles/r2_variants.py $V exp02_v2_full 8192 1
  timeout 200 python profiles/r2_variants.py $V swarm 8192 1
  timeout 200 python profiles/r2_variants.py $V level5_c1 16384 2
  timeout 200 python profiles/r2_variants.py $V exp03_vFinal 65536 2; } > gpurun_out/r2v_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2v_variants.txt
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2v_phase_clocks.txt 2>&1; tail -7 gpurun_out/r2v_phase_clocks.txt
DC_LIB=build/libdc_phases.so timeout 200 python profiles/timeline.py exp02_v2_full 65536 2 | tail -4
