#!/bin/bash
# r2h: new default bench line (exp02_v2_full), reference arm, ncu launch list + --set full capture of the headline kernels,
# ncu of the stand-alone lidar / raycast / mirror kernels, rollout + pipeline tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_adapters.py -q -x > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
timeout 400 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2>> gpurun_out/r2h_bench.err
timeout 300 python bench.py --workload lidar --steps 50 > gpurun_out/r2h_bench_lidar.json 2>> gpurun_out/r2h_bench.err
# ncu only after the plain runs exited 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 20 --warmup 3 --spinup 20 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2h_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dyn_kernel|env_kernel" -s 640 -c 4 -o gpurun_out/r2h_full -f python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-also --no-rollout > gpurun_out/r2h_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lidar_kernel" -s 10 -c 2 -o gpurun_out/r2h_lidar -f python bench.py --workload lidar --steps 20 > gpurun_out/r2h_ncu3.log 2>&1
ls -la gpurun_out | tail -12
