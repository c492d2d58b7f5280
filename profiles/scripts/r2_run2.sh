set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2b_pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_kernels.log
tail -5 gpurun_out/r2b_pytest_kernels.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_kernels_gpu.py > gpurun_out/r2b_pytest_rest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_rest.log
tail -5 gpurun_out/r2b_pytest_rest.log
timeout 300 python bench.py --workload pm25 --steps 10 --warmup 3 --cpu-baseline skip > gpurun_out/r2b_bench_pm25.json 2> gpurun_out/r2b_bench_pm25.err; echo rc=$?
timeout 300 python bench.py --workload hcp --steps 5 --warmup 3 --cpu-baseline skip > gpurun_out/r2b_bench_hcp.json 2> gpurun_out/r2b_bench_hcp.err; echo rc=$?
tail -3 gpurun_out/r2b_bench_pm25.err gpurun_out/r2b_bench_hcp.err
