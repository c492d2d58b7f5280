set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for wl in sim pm25 hcp ecog; do
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-baseline skip > gpurun_out/r2e_bench_$wl.json 2> gpurun_out/r2e_bench_$wl.err; echo "$wl rc=$?"
  tail -n 2 gpurun_out/r2e_bench_$wl.err
done
timeout 300 python bench.py --workload sim --steps 200 --warmup 5 --cpu-baseline skip > gpurun_out/r2e_bench_sim200.json 2> gpurun_out/r2e_bench_sim200.err; echo "rc=$?"
