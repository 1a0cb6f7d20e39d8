set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_level5_multiobs.py tests/test_gpu_rollout.py tests/test_gpu_adapters.py tests/test_gpu_io_data.py -m gpu -q > gpurun_out/r1w_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1w_pytest_new.log
tail -40 gpurun_out/r1w_pytest_new.log
timeout 200 python bench.py --preset level5_dumb_multiobs --envs 8192 --no-e2e --no-cpu > gpurun_out/r1w_bench_level5_dumb_multiobs.json 2> gpurun_out/r1w_bench.err; cut -c1-300 gpurun_out/r1w_bench_level5_dumb_multiobs.json; tail -5 gpurun_out/r1w_bench.err
