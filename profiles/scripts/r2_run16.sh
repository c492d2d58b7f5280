set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dsvi_gpu.py -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -3 gpurun_out/r2u_pytest.log
for tm in 1 0; do
NMGP_GRAM_TMA=$tm timeout 300 python bench.py --workload ecog --steps 5 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2u_ecog_tma$tm.json 2> gpurun_out/r2u_ecog_tma$tm.err; echo "ecog tma=$tm rc=$?"
done
