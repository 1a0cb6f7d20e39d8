#!/bin/bash
set -x
mkdir -p gpurun_out
DC_LIB=build/libdc_phases.so timeout 300 python profiles/phase_clocks.py exp02_v2_full > gpurun_out/r2ae_phase_clocks.txt 2>&1; tail -9 gpurun_out/r2ae_phase_clocks.txt
timeout 200 python profiles/r2_variants.py build/libdc_phases.so exp02_v2_full 65536 2 2>&1 | grep ms/step
timeout 200 python profiles/r2_variants.py build/libdc_phases.so exp02_v2_full 65536 2 2>&1 | grep ms/step
