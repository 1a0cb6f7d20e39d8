set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1v_pytest_gpu.log
tail -15 gpurun_out/r1v_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r1v_bench_exp02_vFinal.json 2> gpurun_out/r1v_bench.err; tail -c 1500 gpurun_out/r1v_bench_exp02_vFinal.json
timeout 200 python bench.py --preset level5_fusion --envs 16384 --no-e2e --no-cpu > gpurun_out/r1v_bench_level5_fusion.json 2>> gpurun_out/r1v_bench.err
timeout 200 python bench.py --preset level5_fusion --envs 16384 --no-e2e --no-cpu --student > gpurun_out/r1v_bench_level5_fusion_student.json 2>> gpurun_out/r1v_bench.err
cut -c1-260 gpurun_out/r1v_bench_level5_fusion*.json
