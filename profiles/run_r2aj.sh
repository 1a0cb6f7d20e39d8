#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2aj_pytest.log
timeout 400 python bench.py --no-cpu --no-also --no-rollout > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; python -c "import json;d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e'])"
