#!/bin/bash
# the adopted policy_kernel build (32 x 32 warp tiles, shallower prefetch): parity tests + the policy bench + the rollout arm of bench.py
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -x -q > gpurun_out/r2aw_pytest.log 2>&1; tail -2 gpurun_out/r2aw_pytest.log
timeout 100 python profiles/r2_policy_bench.py 65536 > gpurun_out/r2aw_policy_bench.json 2>/dev/null; cat gpurun_out/r2aw_policy_bench.json
timeout 200 python bench.py --no-cpu --no-also --no-e2e --steps 100 > gpurun_out/r2aw_bench_rollout.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2aw_bench_rollout.json').read().strip().splitlines()[-1]); r=d['e2e_device_rollout']
print(d['value'], d['ms_per_step'], r['value'], r['policy_tf32']['value'], r['policy_torch_module']['value'])"
