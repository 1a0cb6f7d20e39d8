#!/bin/bash
# second half of the final validation (the first call's gpurun_out/ exceeded the 64 MiB that travel back: its 93 MB ncu report
# took the bench lines with it; the suite -- 97 passed in 280 s, smoke OK -- is quoted in profiles/r2_final_pytest_gpu.log)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_device_rollout']['value'], d['e2e_device_rollout']['policy_tf32']['value'], d['e2e_device_rollout']['policy_torch_module']['value'], d['cpu_baseline']['value'], d['clocks'])"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 20 --warmup 3 --spinup 20 --no-cpu --no-also --e2e-steps 3 > gpurun_out/r2_final_ncu1.log 2>&1; tail -1 gpurun_out/r2_final_ncu1.log | cut -c1-200
AUX_SHORT=1 timeout 500 ncu --set full --clock-control none -k regex:"lidar_kernel|raycast_kernel|scatter|diff_hits|lw_obs_kernel|stack_kernel" -c 20 -o gpurun_out/r2_final_aux -f python profiles/r2_aux_kernels.py > gpurun_out/r2_final_ncu2.log 2>&1; tail -2 gpurun_out/r2_final_ncu2.log
ncu -i gpurun_out/r2_final_aux.ncu-rep --page raw --csv > gpurun_out/r2_final_aux_raw.csv 2>/dev/null
[ $(stat -c %s gpurun_out/r2_final_aux.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/r2_final_aux.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | tail -12
