set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dsvi_gpu.py -x -q > gpurun_out/r2C_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2C_pytest.log
tail -3 gpurun_out/r2C_pytest.log
B="--cpu-baseline skip --no-e2e --others skip"
for gb in 6 13 50; do
NMGP_SAMPLE_BUDGET_GB=$gb timeout 200 python bench.py --workload ecog --steps 5 --warmup 3 $B > gpurun_out/r2C_ecog_gb$gb.json 2> gpurun_out/r2C_ecog_gb$gb.err; echo "ecog gb=$gb rc=$?"
done
