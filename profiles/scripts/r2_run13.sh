set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dsvi_gpu.py -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r_pytest.log
for hv in 1 2; do
NMGP_GRAM_HALVES=$hv timeout 300 python bench.py --workload ecog --steps 5 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2r_ecog_hv$hv.json 2> gpurun_out/r2r_ecog_hv$hv.err; echo "ecog hv=$hv rc=$?"
done
timeout 300 python bench.py --workload pm25 --steps 10 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2r_pm25.json 2> gpurun_out/r2r_pm25.err; echo "pm25 rc=$?"
