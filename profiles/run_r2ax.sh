#!/bin/bash
# last confirmation on the final tree: smoke + the closed-loop stage03 tests, the numpy adapters and the policy kernel
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 170 python -m pytest tests/test_gpu_stage03.py tests/test_gpu_adapters.py tests/test_gpu_policy.py tests/test_gpu_driven.py -q > gpurun_out/r2ax_pytest.log 2>&1; tail -2 gpurun_out/r2ax_pytest.log
