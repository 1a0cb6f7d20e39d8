#!/bin/bash
# r2s: dyn_kernel launch-bound variants (5..8 resident blocks per SM = 96 / 80 / 72 / 64 registers) x sub-batch counts
set -x
mkdir -p gpurun_out
for n in 5 6 7 8; do timeout 200 python profiles/r2_variants.py build/libdc_dyn$n.so exp02_v2_full 65536 1 2 3; done > gpurun_out/r2s_variants.txt 2>&1
for n in 5 6 8; do timeout 200 python profiles/r2_variants.py build/libdc_dyn$n.so exp02_v2_full 8192 1; timeout 200 python profiles/r2_variants.py build/libdc_dyn$n.so swarm 8192 1 2;  done >> gpurun_out/r2s_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2s_variants.txt
