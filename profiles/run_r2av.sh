#!/bin/bash
# A/B of policy_kernel builds (profiles/micro/variants/*.so: -DDCP_WIDE / -DDCP_PF_TF32 / -DDCP_PF_X3) against the in-tree build
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -2
: > gpurun_out/r2av_variants.txt
timeout 100 python profiles/r2_policy_bench.py 65536 >> gpurun_out/r2av_variants.txt 2>/dev/null
for v in wide pf4 widepf4 wide_pf2_1; do timeout 100 python profiles/r2_policy_bench.py 65536 profiles/micro/variants/libdc_$v.so >> gpurun_out/r2av_variants.txt 2>/dev/null; done
python - <<'PY'
import json
for l in open('gpurun_out/r2av_variants.txt'):
    d = json.loads(l)
    print("%-50s 3xTF32 %.3f ms (err %.1e)   TF32 %.3f ms (err %.1e)" % (d["lib"], d["fused_3xtf32_ms"], d["fused_3xtf32_err"], d["fused_tf32_ms"], d["fused_tf32_err"]))
PY
