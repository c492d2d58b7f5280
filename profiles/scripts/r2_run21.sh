set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_simcode_gpu.py tests/test_dsvi_gpu.py -x -q > gpurun_out/r2A_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2A_pytest.log
tail -3 gpurun_out/r2A_pytest.log
B="--cpu-baseline skip --no-e2e --others skip"
timeout 200 python bench.py --workload ecog --steps 5 --warmup 3 $B > gpurun_out/r2A_ecog.json 2> gpurun_out/r2A_ecog.err; echo "ecog rc=$?"
timeout 200 python bench.py --workload pm25 --steps 10 --warmup 3 $B > gpurun_out/r2A_pm25.json 2> gpurun_out/r2A_pm25.err; echo "pm25 rc=$?"
timeout 200 python bench.py --workload hcp --steps 10 --warmup 3 $B > gpurun_out/r2A_hcp.json 2> gpurun_out/r2A_hcp.err; echo "hcp rc=$?"
