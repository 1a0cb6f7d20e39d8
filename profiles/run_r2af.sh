#!/bin/bash
# r2af: the sphere update as a change list (dc_diff_hits + dc_host_apply_pairs) against the mapped mirror: parity, end-to-end step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q -k "sparse_lidar" > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2af_pytest.log
NO_TERMINAL_DICTS=1 timeout 600 python profiles/r2_e2e_modes.py exp02_v2_full 65536 > gpurun_out/r2af_e2e_modes.txt 2>&1; cat gpurun_out/r2af_e2e_modes.txt | tail -8
