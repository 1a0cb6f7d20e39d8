#!/bin/bash
set -x
mkdir -p gpurun_out
for v in env4 env5; do timeout 200 python profiles/r2_variants.py build/libdc_$v.so exp02_v2_full 65536 1 2; timeout 200 python profiles/r2_variants.py build/libdc_$v.so swarm 8192 1; timeout 200 python profiles/r2_variants.py build/libdc_$v.so level5_c1 16384 2; timeout 200 python profiles/r2_variants.py build/libdc_$v.so exp02_v2_full 8192 1; done > gpurun_out/r2ab_variants.txt 2>&1
grep -E "ms/step|Error" gpurun_out/r2ab_variants.txt
