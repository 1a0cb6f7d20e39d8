#!/bin/bash
# r2w: can every armed drone of the batch be resident at once?  dyn_kernel with 64-thread blocks / 56-64 registers
set -x
mkdir -p gpurun_out
for v in t64_b16 t64_b18 t128_b9 t64_b20; do timeout 200 python profiles/r2_variants.py build/libdc_$v.so exp02_v2_full 65536 1 2; done > gpurun_out/r2w_variants.txt 2>&1
for v in t64_b18 t64_b20; do timeout 200 python profiles/r2_variants.py build/libdc_$v.so exp02_v2_full 8192 1; timeout 200 python profiles/r2_variants.py build/libdc_$v.so swarm 8192 1; timeout 200 python profiles/r2_variants.py build/libdc_$v.so level5_c1 16384 2; done >> gpurun_out/r2w_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2w_variants.txt
