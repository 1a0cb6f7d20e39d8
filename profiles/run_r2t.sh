#!/bin/bash
set -x
mkdir -p gpurun_out
for k in 1 2 3; do DC_LIB=build/libdc_phases.so timeout 200 python profiles/timeline.py exp02_v2_full 65536 $k; done > gpurun_out/r2t_timeline.txt 2>&1
DC_LIB=build/libdc_phases.so timeout 200 python profiles/timeline.py exp02_v2_full 8192 1 >> gpurun_out/r2t_timeline.txt 2>&1
tail -50 gpurun_out/r2t_timeline.txt
