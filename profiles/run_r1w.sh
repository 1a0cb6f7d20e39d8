set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_level5_multiobs.py tests/test_gpu_level5.py -m gpu -q > gpurun_out/r1w_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1w_pytest_new.log
tail -40 gpurun_out/r1w_pytest_new.log
