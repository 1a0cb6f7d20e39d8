#!/bin/bash
set -x
mkdir -p gpurun_out
for m in mapped pairs; do timeout 300 python profiles/r2_e2e_where.py exp02_v2_full 65536 $m; done > gpurun_out/r2ag_e2e_where.txt 2>&1
cat gpurun_out/r2ag_e2e_where.txt
