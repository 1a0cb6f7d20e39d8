#!/bin/bash
# r2u: programmatic dependent launch of the step kernels: timelines with DC_PDL = 0 / 1 / 2 (profiling build), product timings, parity subset
set -x
mkdir -p gpurun_out
for pdl in 0 1 2; do for k in 1 2; do echo "== DC_PDL=$pdl K=$k"; DC_PDL=$pdl DC_LIB=build/libdc_phases.so timeout 200 python profiles/timeline.py exp02_v2_full 65536 $k | tail -4; done; done > gpurun_out/r2u_timeline.txt 2>&1
for pdl in 0 2; do echo "== DC_PDL=$pdl 8192"; DC_PDL=$pdl DC_LIB=build/libdc_phases.so timeout 200 python profiles/timeline.py exp02_v2_full 8192 1 | tail -3; done >> gpurun_out/r2u_timeline.txt 2>&1
cat gpurun_out/r2u_timeline.txt
for pdl in 0 1 2; do echo "== DC_PDL=$pdl"; DC_PDL=$pdl timeout 200 python profiles/r2_variants.py build/libdc_phases.so exp02_v2_full 65536 1 2; done > gpurun_out/r2u_variants.txt 2>&1
timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 65536 1 2 >> gpurun_out/r2u_variants.txt 2>&1
timeout 200 python profiles/r2_variants.py dronechase_b200/csrc/libdronechase_b200.so exp02_v2_full 8192 1 >> gpurun_out/r2u_variants.txt 2>&1
grep -E "==|ms/step|Error" gpurun_out/r2u_variants.txt
timeout 900 python -m pytest tests/test_gpu_stage03.py tests/test_gpu_full_size.py tests/test_gpu_sub_batches.py tests/test_gpu_level5.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest.log
