set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --workload pm25 --steps 10 --warmup 3 --cpu-baseline skip > gpurun_out/r2c_bench_pm25.json 2> gpurun_out/r2c_bench_pm25.err; echo rc=$?
timeout 300 python bench.py --workload hcp --steps 5 --warmup 3 --cpu-baseline skip > gpurun_out/r2c_bench_hcp.json 2> gpurun_out/r2c_bench_hcp.err; echo rc=$?
timeout 300 python bench.py --workload ecog --steps 5 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2c_bench_ecog.json 2> gpurun_out/r2c_bench_ecog.err; echo rc=$?
timeout 300 python bench.py --workload sim --steps 50 --warmup 5 --cpu-baseline skip > gpurun_out/r2c_bench_sim.json 2> gpurun_out/r2c_bench_sim.err; echo rc=$?
for f in pm25 hcp ecog sim; do tail -n 3 gpurun_out/r2c_bench_$f.err; done
