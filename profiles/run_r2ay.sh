#!/bin/bash
# ncu --set full of the adopted policy_kernel build (32 x 32 warp tiles): launches 25-26 are 3xTF32, 27-28 TF32
mkdir -p gpurun_out
timeout 90 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 24 -c 4 -o gpurun_out/r2ay_policy -f python profiles/r2_policy_bench.py 65536 > gpurun_out/r2ay_ncu.log 2>&1; tail -1 gpurun_out/r2ay_ncu.log
