set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
for wl in pm25 hcp sim; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --cpu-baseline skip > gpurun_out/r2j_bench_$wl.json 2> gpurun_out/r2j_bench_$wl.err; echo "$wl rc=$?"
  tail -n 2 gpurun_out/r2j_bench_$wl.err
done
