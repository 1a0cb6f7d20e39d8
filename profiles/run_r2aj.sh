#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapters.py -m gpu -x -q > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2aj_pytest.log
timeout 600 python profiles/r2_e2e_ab.py exp02_v2_full 65536 > gpurun_out/r2aj_e2e_ab.txt 2>&1; cat gpurun_out/r2aj_e2e_ab.txt
timeout 300 python profiles/r2_e2e_where.py exp02_v2_full 65536 pairs 2>&1 | tail -8
timeout 400 python bench.py --no-cpu --no-also --no-rollout > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; python -c "import json;d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e'])"
