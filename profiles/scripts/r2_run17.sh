set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2v_pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest_k.log
tail -3 gpurun_out/r2v_pytest_k.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -3 gpurun_out/r2v_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench_default.json 2> gpurun_out/r2v_bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2v_bench_default.err
