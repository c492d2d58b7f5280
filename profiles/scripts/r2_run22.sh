set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fullsize_properties_gpu.py -x -q > gpurun_out/r2B_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2B_pytest.log
tail -5 gpurun_out/r2B_pytest.log
B="--cpu-baseline skip --no-e2e --no-graph --others skip"
timeout 300 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2B_plain_pm25.json 2> gpurun_out/r2B_plain_pm25.err; echo "pm25 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_lq|k_gram_mma" -s 12 -c 3 -o gpurun_out/r2B_prof_pm25 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2B_ncu_pm25.log 2>&1; echo "ncu pm25 rc=$?"
