#!/bin/bash
# final validation of the round-2 build: whole GPU suite, default bench line (+ the reference arm), ncu launch list of the same
# command, --set full of the auxiliary kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest_gpu.log
tail -4 gpurun_out/r2_final_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_final_bench.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err; cut -c1-400 gpurun_out/r2_final_bench_reference.json
# ncu only after the plain runs exited 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 20 --warmup 3 --spinup 20 --no-cpu --no-also --e2e-steps 3 > gpurun_out/r2_final_ncu1.log 2>&1; tail -2 gpurun_out/r2_final_ncu1.log
timeout 200 python profiles/r2_aux_kernels.py > gpurun_out/r2_final_aux.log 2>&1; tail -2 gpurun_out/r2_final_aux.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"lidar_kernel|raycast_kernel|scatter|diff_hits|lw_obs_kernel|stack_kernel" -c 40 -o gpurun_out/r2_final_aux -f python profiles/r2_aux_kernels.py > gpurun_out/r2_final_ncu2.log 2>&1; tail -2 gpurun_out/r2_final_ncu2.log
ls -la gpurun_out | tail -12
