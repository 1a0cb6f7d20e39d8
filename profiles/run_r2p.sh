#!/bin/bash
# r2p: where env_kernel's warp chain goes (phase clocks), sub-batch sweep on the new dynamics
set -x
mkdir -p gpurun_out
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2p_phase_clocks.txt 2>&1; tail -12 gpurun_out/r2p_phase_clocks.txt
for k in 1 2 3 4; do timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 65536 $k; done > gpurun_out/r2p_variants.txt 2>&1
timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 8192 1 2 >> gpurun_out/r2p_variants.txt 2>&1
tail -20 gpurun_out/r2p_variants.txt
