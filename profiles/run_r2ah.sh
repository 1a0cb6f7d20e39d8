#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python profiles/r2_e2e_ab.py exp02_v2_full 65536 > gpurun_out/r2ah_e2e_ab.txt 2>&1
cat gpurun_out/r2ah_e2e_ab.txt
